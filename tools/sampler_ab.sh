#!/bin/bash
# A/B of the extinction sampler variants (atm_device.cuh) on one GPU
mkdir -p gpurun_out
for cfg in "0 4" "0 2" "1 2" "2 2"; do
  set -- $cfg
  ATMLUT_POW=$1 ATMLUT_DEGREE=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/sampler_$1_$2.json 2> gpurun_out/sampler_$1_$2.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/sampler_$1_$2.json") if l.startswith("{")][-1])
    print("pow $1 degree $2: build %.3f ms first_order %.3f" % (d["ms_per_step"], d["stage_ms"]["first_order"]))
except Exception as e:
    print("pow $1 degree $2 failed", e)
PY
done
