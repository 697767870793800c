"""Diagnostic (not a test): one batched forward + backward of the Twitter shortest-path LP (221 x 2852, per-instance b,
SURVEY config 3) — the process the Twitter ncu captures under profiles/ run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from neuripsintopt_b200.ip_model_whole import IPOfunc
from tests import golden_util as GU

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
z = GU.load("twitter_sp")
case = GU.cases(z)[0]
kw = dict(case["kw"], thr=0.1)
inp = GU.problem_inputs("twitter_sp", z, 0)
rng = np.random.default_rng(0)
c0 = case["c"][0].astype(np.float64)
bb = np.tile(z["b_rr"], (B // len(z["b_rr"]) + 1, 1))[:B]
c = torch.tensor((c0 * np.abs(1 + 0.3 * rng.standard_normal((B, c0.size)))).astype(np.float32), device="cuda", requires_grad=True)
tt = lambda v: torch.as_tensor(np.asarray(v), dtype=torch.float32)  # noqa: E731
f = IPOfunc(tt(inp["A"]), tt(bb), torch.Tensor(), torch.Tensor(), check_finite=False, **kw)
for _ in range(2):
    cc = c.detach().requires_grad_(True)
    x = f(cc)
    x.backward(torch.ones_like(x) / c0.size)
torch.cuda.synchronize()
status, iters = IPOfunc.status()
print("twitter B=%d mean iters %.2f bad %d" % (B, float(iters.float().mean()), int((status != 0).sum())))
