"""Diagnostic (not a test): per-phase SM-cycle breakdown of the forward kernel on the Energy LP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from neuripsintopt_b200 import _capi
from neuripsintopt_b200.ip_model_whole import IPOfunc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
A, b, G, h, F = bench.energy_problem()
tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)
f = IPOfunc(tt(A), tt(b), tt(G), tt(h), check_finite=False, **bench.OPTS)
c, g = bench.synthetic_costs(F, B, 1)
cd = torch.tensor(c, device="cuda")
f(cd); torch.cuda.synchronize()
_capi.phase_cycles(True)
f(cd); torch.cuda.synchronize()
ph = _capi.phase_cycles(False, read=True)
tot = sum(ph.values())
print("B=%d  total CTA-cycles %.3g  per LP %.0f cycles" % (B, tot, tot / B))
for k, v in ph.items():
    print("  %-24s %10.0f cycles/LP  %5.1f%%" % (k, v / B, 100.0 * v / tot))
