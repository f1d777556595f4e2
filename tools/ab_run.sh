#!/bin/bash
# A/B of the kernel variants on one GPU: each line = one bench.py run (no extras) with one variant switched.
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$name.json"))
    s = d["stage_ms"]
    print("%-22s build %.3f ms  first_order %.3f  point_scatter %.3f  ray_scatter %.3f  final %.3f  e2e %.3f ms  pageable %.3f ms" % (
        "$name", d["ms_per_step"], s["first_order"], s["point_scatter"], s["ray_scatter"], s["final_accumulate_and_files"],
        d["e2e"]["build_time_s"] * 1e3, d.get("e2e_pageable", {}).get("build_time_s", 0) * 1e3))
except Exception as e:
    print("$name failed:", e, open("gpurun_out/ab_$name.err").read()[-400:])
PY
}
run all_new X=1
run old_k6 ATMLUT_K6=1
run old_k4 ATMLUT_K4=1
run no_poly ATMLUT_K3_POLY=0
run all_old ATMLUT_K6=1 ATMLUT_K4=1 ATMLUT_K3_POLY=0
