#!/usr/bin/env python
"""Pretty-print a bench.py JSON line: per workload ms, kernels, roofline fractions.  Usage: tools/show_bench.py FILE..."""
import json, sys
for f in sys.argv[1:]:
    txt = open(f).read().strip()
    if not txt:
        print(f, "EMPTY"); continue
    d = json.loads(txt.splitlines()[-1])
    print("==", f, "n_gpus", d.get("n_gpus"), "value %.3g" % (d.get("value") or 0), "ms %.3f" % (d.get("ms_per_step") or 0),
          "e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (d.get("e2e") or {}).items() if k in ("ms_per_step", "mode", "result_ok")})
    if d.get("e2e", {}).get("pipelined"):
        print("   e2e pipelined ms %.1f" % d["e2e"]["pipelined"]["ms_per_step"])
    if d.get("phases_ms_rank0"):
        print("   phases", {k: round(v, 3) for k, v in d["phases_ms_rank0"].items()})
    for name, w in (d.get("workloads") or {}).items():
        if "error" in w:
            print("  %-20s ERROR %s" % (name, w["error"])); continue
        r = w.get("roofline") or {}
        print("  %-20s %8.3f ms  path_frac %.3f  ok=%s  top=%s frac=%s" % (name, w["ms_per_step"], w.get("path_frac_of_measured_hbm", 0) or 0,
              w.get("parity_properties_ok"), r.get("kernel"), ("%.3f" % r["frac"]) if r.get("frac") else None))
        ks = w.get("kernels") or {}
        if ks:
            print("      " + "  ".join("%s=%.3f(x%g)" % (k, v["ms_per_step"], v["launches_per_step"]) for k, v in sorted(ks.items())))
        if w.get("phases_ms_rank0"):
            print("      phases", {k: round(v, 3) for k, v in w["phases_ms_rank0"].items()})
