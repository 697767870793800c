"""Helper: one-line summary of a bench.py JSON line read from stdin (tuning aid, not a test)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        r = d.get("roofline", {})
        print(tag, "value %.0f LP/s  %.3f ms/step  e2e %.0f  fwd %.3f bwd %.3f  frac %.3f  clocks %s  per-rank %s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("fwd_ms", 0), r.get("bwd_ms", 0),
            r.get("step", {}).get("frac", 0), d.get("clocks"), d.get("per_rank_ms_per_step")))
