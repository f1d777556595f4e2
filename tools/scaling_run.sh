#!/bin/bash
# 1/2/4/8-GPU strong-scaling sweep of bench.py (shipped and stress workloads); writes gpurun_out/scale_*.json
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
MODE=${MODE:-p2p}
GPUS=${GPUS:-"1 2 4 8"}
for wl in shipped stress; do
  steps=20; [ $wl = stress ] && steps=5
  for n in $GPUS; do
    [ $n -gt $NG ] && continue
    if [ $n = 1 ]; then
      python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/scale_${MODE}_${wl}_$n.json 2> gpurun_out/scale_${MODE}_${wl}_$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --mode $MODE --workload $wl --steps $steps --warmup 3 2> gpurun_out/scale_${MODE}_${wl}_$n.err | grep '^{' > gpurun_out/scale_${MODE}_${wl}_$n.json
    fi
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${MODE}_${wl}_$n.json"))
    print("${MODE}", "${wl}", $n, "gpus: %.3f ms/build  %.3e texels/s  e2e %.3f ms  allgather %.3f ms" % (d["ms_per_step"], d["value"], d["e2e"]["build_time_s"]*1e3, sum(v for k,v in d["stage_ms"].items() if "allgather" in k)))
except Exception as e:
    print("${MODE}", "${wl}", $n, "failed", e)
PY
  done
done
