"""Print the fields of a bench.py JSON line we steer by (the line may be preceded by library chatter)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception as e:
        print(path, "no JSON line:", e)
        continue
    s = d.get("stage_ms", {})
    print("%s: n_gpus %d  build %.3f ms  e2e %.3f ms  pageable %s  launches %s" % (
        path, d["n_gpus"], d["ms_per_step"], d["e2e"]["build_time_s"] * 1e3,
        "%.3f ms" % (d["e2e_pageable"]["build_time_s"] * 1e3) if "e2e_pageable" in d else "-", d.get("kernel_launches_per_build")))
    print("   stages", {k: round(v, 3) for k, v in s.items()})
    if "stage_ms_max_over_ranks" in d and d["n_gpus"] > 1:
        print("   max   ", {k: round(v, 3) for k, v in d["stage_ms_max_over_ranks"].items()})
    for k in ("sharded_identical", "sharded_identical_detail"):
        if k in d:
            print("   %s %s" % (k, d[k]))
    for k in ("nccl_mode", "p2p_mode", "stress"):
        if k in d:
            print("   %s: %.3f ms %s" % (k, d[k]["ms_per_step"], {a: b for a, b in d[k].items() if a in ("sharded_identical", "stage_ms_rank0", "all_gathers_per_build")}))
    r = d.get("roofline") or {}
    r6 = d.get("roofline_k6") or {}
    print("   roofline K3 frac %s  K6 frac %s  clocks %s" % (r.get("frac"), r6.get("frac"), d.get("clocks")))
