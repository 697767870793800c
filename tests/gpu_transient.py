"""Diagnostic: per-step forward/backward times from process start (is there a warm-up transient?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from neuripsintopt_b200.ip_model_whole import IPOfunc
A, b, G, h, F = bench.energy_problem()
tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)
f = IPOfunc(tt(A), tt(b), tt(G), tt(h), check_finite=False, **bench.OPTS)
c, g = bench.synthetic_costs(F, 1024, 1)
cd, gd = torch.tensor(c, device="cuda"), torch.tensor(g, device="cuda")
t0 = time.time()
out = []
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    cc = cd.detach().requires_grad_(True)
    e0.record(); x = f(cc); em.record(); x.backward(gd); e1.record(); e1.synchronize()
    out.append((time.time() - t0, e0.elapsed_time(em), em.elapsed_time(e1)))
for k in list(range(0, 12)) + list(range(12, len(out), 8)):
    print("step %3d  t=%.2fs  fwd %.3f ms  bwd %.3f ms" % (k, *out[k]))
