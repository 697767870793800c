"""Prints the results of tests/ab_variants.sh (gpurun_out/ab_<v>.json, abcfg_<v>.jsonl): python tests/ab_show.py old main ..."""
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in sys.argv[1:]:
    try:
        d = json.loads(open(os.path.join(root, "gpurun_out", "ab_%s.json" % v)).read().strip().splitlines()[-1])
        line = "%-6s %8.0f LP/s fwd %.3f bwd %.3f ms" % (v, d["value"], d["roofline"]["fwd_ms"], d["roofline"]["bwd_ms"])
    except Exception as e:  # noqa: BLE001
        line = "%-6s ERR %s" % (v, e)
    try:
        for l in open(os.path.join(root, "gpurun_out", "abcfg_%s.jsonl" % v)):
            d = json.loads(l)
            line += " | %s %.0f" % (d["config"][:8], d["lp_per_s"])
    except Exception:  # noqa: BLE001
        line += " | (no config lines)"
    print(line)
