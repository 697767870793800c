"""Diagnostic (not a test): SM cycles of each solver building block on the Energy LP, alone on one SM and with
every CTA slot of the device busy (2 CTAs/SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from neuripsintopt_b200 import _capi, ipofunc as _ipo
from neuripsintopt_b200.ip_model_whole import IPOfunc
A, b, G, h, F = bench.energy_problem()
tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)
f = IPOfunc(tt(A), tt(b), tt(G), tt(h), check_finite=False, **bench.OPTS)
c, g = bench.synthetic_costs(F, 4, 1)
f(torch.tensor(c, device="cuda")); torch.cuda.synchronize()
dp = [v for v in _ipo._PROBLEM_CACHE.values() if hasattr(v, "handle")][0]
which = [int(a) for a in sys.argv[1:]] or range(len(_capi.MICRO))
for w in which:
    base1 = _capi.microbench(dp.handle, 0, 50, 1)[0]
    a1 = _capi.microbench(dp.handle, w, 50, 1)
    a296 = _capi.microbench(dp.handle, w, 50, 296)
    print("%-24s alone %8.0f cyc   full device: median %8.0f max %8.0f" % (_capi.MICRO[w], a1[0], a296[0], a296[1]))
