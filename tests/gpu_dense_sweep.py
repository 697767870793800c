"""Diagnostic (not a test): dense LP sweep of SURVEY §8(d) config 5 (m x 2m, feasible & bounded generator): parity of a
few instances against the oracle and throughput."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuripsintopt_b200.ip_model_whole import IPOfunc
from neuripsintopt_b200 import ipofunc as _ipo
from oracle import hsd_numpy as O

def gen(m, B):
    rng = np.random.default_rng(m)
    n = 2 * m
    A = rng.standard_normal((m, n)).astype(np.float32)
    x0 = rng.uniform(0.1, 1.0, n)
    b = (A.astype(np.float64) @ x0).astype(np.float32)
    y0 = rng.standard_normal((B, m)); s0 = rng.uniform(0.1, 1.0, (B, n))
    c = (y0 @ A.astype(np.float64) + s0).astype(np.float32)
    return A, b, c

for m in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    B = 256 if m <= 128 else 64
    A, b, c = gen(m, B)
    t0 = time.time()
    f = IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=0.1,
                damping=1e-6, check_finite=False)
    ct = torch.tensor(c, device="cuda", requires_grad=True)
    g = torch.ones_like(ct)
    x = f(ct); x.backward(g); torch.cuda.synchronize()
    t_first = time.time() - t0
    status, iters = IPOfunc.status()
    info = [v.info for v in _ipo._PROBLEM_CACHE.values() if hasattr(v, "info")][-1]
    ex = eg = 0.0
    for i in range(2):
        r = O.ipofunc(c[i], np.ones(2 * m, np.float32), A=A, b=b, bounds=[(0.0, None)], thr=0.1, damping=1e-6)
        ex = max(ex, float(np.abs(x[i].detach().double().cpu().numpy() - r["x"]).max() / max(1.0, np.abs(r["x"]).max())))
        eg = max(eg, float(np.linalg.norm(ct.grad[i].double().cpu().numpy() - r["grad"]) / np.linalg.norm(r["grad"])))
        assert int(iters[i]) == r["iters"], (int(iters[i]), r["iters"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        cc = ct.detach().requires_grad_(True); f(cc).backward(g)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("m=%d n=%d B=%d: x_rel %.1e grad_rel %.1e iters %.2f bad %d | %.0f LP/s (%.2f ms/step) | smem %d B factor_in_smem %d gather_in_smem %d ctas/sm %d threads %d | setup+first %.1f s" % (
        m, 2 * m, B, ex, eg, float(iters.float().mean()), int((status != 0).sum()), B / ms * 1e3, ms,
        info["smem_bytes"], info["factor_in_smem"], info["gather_in_smem"], info["ctas_per_sm"], info["threads"], t_first), flush=True)
