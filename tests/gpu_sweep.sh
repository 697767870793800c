#!/usr/bin/env bash
# Tuning helper (not a test): sweeps CTA size / residency of the solver kernels on the GPU box.
for T in ${THREADS_LIST:-256 384 512}; do for C in ${CTAS_LIST:-1 2 3}; do
  echo "T=$T C=$C"
  INTOPT_THREADS=$T INTOPT_MAX_CTAS_PER_SM=$C python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu ${EXTRA:-} 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        print("  value %.0f LP/s  e2e %.0f  fwd %.3f ms bwd %.3f ms  step-frac %.3f exec-frac %.3f ctas/sm %s smem %s" % (
            d["value"], d["e2e"]["value"], d["roofline"]["fwd_ms"], d["roofline"]["bwd_ms"], d["roofline"]["step"]["frac"],
            d["roofline"]["step"]["executed_frac"], d["problem"]["ctas_per_sm"], d["problem"]["smem_bytes"]))
    elif "Error" in l or "error" in l or "assert" in l.lower():
        print("  " + l.strip()[:300])
'
done; done
