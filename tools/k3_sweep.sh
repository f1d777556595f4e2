#!/bin/bash
# A/B of the first-order kernel's CTA shape on one GPU (ATMLUT_FIRST_ORDER_WARPS x ATMLUT_FIRST_ORDER_PASSES)
mkdir -p gpurun_out
for w in 8 4 2; do for p in 16 8 4; do
  ATMLUT_FIRST_ORDER_WARPS=$w ATMLUT_FIRST_ORDER_PASSES=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/k3_sweep_${w}_${p}.json 2> gpurun_out/k3_sweep_${w}_${p}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/k3_sweep_${w}_${p}.json") if l.startswith("{")][-1])
    print("warps $w passes $p: build %.3f ms first_order %.3f frac %.3f" % (d["ms_per_step"], d["stage_ms"]["first_order"], d["roofline"]["frac"]))
except Exception as e:
    print("warps $w passes $p failed", e)
PY
done; done
