"""Diagnostic: forward time with/without an L2 flush before it, and with concurrent pinned copies."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from neuripsintopt_b200.ip_model_whole import IPOfunc
A, b, G, h, F = bench.energy_problem()
tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)
f = IPOfunc(tt(A), tt(b), tt(G), tt(h), check_finite=False, **bench.OPTS)
c, g = bench.synthetic_costs(F, 1024, 1)
cd, gd = torch.tensor(c, device="cuda"), torch.tensor(g, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
hbuf = torch.empty(8 * 1024 * 1440, dtype=torch.float32).pin_memory(); dbuf = torch.empty_like(hbuf, device="cuda")
side = torch.cuda.Stream()
def run(mode, n=30):
    tf = tb = 0.0
    for k in range(n + 5):
        if mode == "flush": flush.zero_()
        if mode == "copy":
            with torch.cuda.stream(side): dbuf.copy_(hbuf, non_blocking=True); hbuf.copy_(dbuf, non_blocking=True)
        e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        cc = cd.detach().requires_grad_(True)
        e0.record(); x = f(cc); em.record(); x.backward(gd); e1.record(); e1.synchronize()
        if k >= 5: tf += e0.elapsed_time(em); tb += em.elapsed_time(e1)
    torch.cuda.synchronize()
    print("%-6s fwd %.3f ms  bwd %.3f ms" % (mode, tf / n, tb / n))
for mode in ("warm", "flush", "copy", "warm"):
    run(mode)
