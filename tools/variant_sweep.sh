#!/bin/bash
# bench.py (kernel-only) with every library variant under build/variants/ (differently compiled kernels) and the default
mkdir -p gpurun_out
for lib in default build/variants/*.so; do
  name=$(basename $lib .so)
  if [ "$lib" = default ]; then unset SFSIM_ATMOSPHERE_LIB; else export SFSIM_ATMOSPHERE_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/var_$name.json") if l.startswith("{")][-1])
    s = d["stage_ms"]
    print("%-12s build %.3f ms first_order %.3f point_scatter %.3f ray_scatter %.3f frac %.3f" % ("$name", d["ms_per_step"], s["first_order"], s["point_scatter"], s["ray_scatter"], d["roofline"]["frac"]))
except Exception as e:
    print("$name failed", e)
PY
done
