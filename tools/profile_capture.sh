#!/bin/bash
# Run on the GPU box (one GPU): the ncu launch list of three builds and one `--set full` capture of each of the
# three heavy kernels.  Reports land in gpurun_out/; summarise them afterwards with tools/ncu_summary.py.
#   gpurun -- bash tools/profile_capture.sh r2_final
tag=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/profile_run.py > gpurun_out/${tag}_launches.log 2>&1
# -s: launches to skip, so that the capture is a warm, steady-state one (first_order: 1 per build; the others: 5)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_first_order -s 1 -c 1 -f \
    -o gpurun_out/${tag}_k_first_order python tools/profile_run.py > gpurun_out/${tag}_ncu_k_first_order.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ray_scatter -s 7 -c 1 -f \
    -o gpurun_out/${tag}_k_ray_scatter python tools/profile_run.py > gpurun_out/${tag}_ncu_k_ray_scatter.log 2>&1
# k_point_scatter_prepare matches the same expression: 1 + 5 launches per build, skip into the second build's pass 3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_point_scatter -s 9 -c 1 -f \
    -o gpurun_out/${tag}_k_point_scatter python tools/profile_run.py > gpurun_out/${tag}_ncu_k_point_scatter.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cube_color -s 1 -c 1 -f \
    -o gpurun_out/${tag}_k_cube_color python tools/cubemap_run.py 4 2 > gpurun_out/${tag}_ncu_k_cube_color.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
