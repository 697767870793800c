"""Diagnostic (not a test): per-phase SM-cycle breakdown of the forward kernel on the Twitter shortest-path LP
(221 x 2852, per-instance b; 512-thread CTAs, one per SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuripsintopt_b200 import _capi
from neuripsintopt_b200 import ipofunc as _ipo
from neuripsintopt_b200.ip_model_whole import IPOfunc
from tests import golden_util as GU
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
z = GU.load("twitter_sp"); case = GU.cases(z)[0]; kw = dict(case["kw"])
inp = GU.problem_inputs("twitter_sp", z, 0)
tt = lambda v: torch.as_tensor(np.asarray(v), dtype=torch.float32)
bb = np.tile(z["b_rr"], (B // len(z["b_rr"]) + 1, 1))[:B]
rng = np.random.default_rng(0)
c0 = case["c"][0].astype(np.float64)
c = torch.tensor((c0 * np.abs(1 + 0.3 * rng.standard_normal((B, c0.size)))).astype(np.float32), device="cuda")
f = IPOfunc(tt(inp["A"]), tt(bb), torch.Tensor(), torch.Tensor(), check_finite=False, **kw)
f(c); torch.cuda.synchronize()
status, iters = IPOfunc.status()
info = [v.info for v in _ipo._PROBLEM_CACHE.values() if hasattr(v, "info")][-1]
_capi.phase_cycles(True)
f(c); torch.cuda.synchronize()
ph = _capi.phase_cycles(False, read=True)
tot = sum(ph.values())
print("# Twitter LP %dx%d, envelope %d doubles, threads %d, ctas/sm %d, smem %d B, mean iterations %.2f" % (
    info["m"], info["n"], info["envelope"], info["threads"], info["ctas_per_sm"], info["smem_bytes"], float(iters.float().mean())))
print("B=%d  total CTA-cycles %.3g  per LP %.0f cycles" % (B, tot, tot / B))
for k, v in ph.items():
    print("  %-24s %10.0f cycles/LP  %5.1f%%" % (k, v / B, 100.0 * v / tot))
