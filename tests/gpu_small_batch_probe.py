import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from neuripsintopt_b200.ip_model_whole import IPOfunc
A, b, G, h, F = bench.energy_problem()
tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)
f = IPOfunc(tt(A), tt(b), tt(G), tt(h), check_finite=False, **bench.OPTS)
for B in (24, 128, 296, 1024):
    c, g = bench.synthetic_costs(F, B, 1)
    cd, gd = torch.tensor(c, device="cuda"), torch.tensor(g, device="cuda")
    def step():
        cc = cd.detach().requires_grad_(True)
        f(cc).backward(gd)
    for _ in range(20): step()
    torch.cuda.synchronize()
    K = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K): step()
    e1.record(); t_enq = time.perf_counter() - t0
    e1.synchronize(); t_all = time.perf_counter() - t0
    print(json.dumps(dict(B=B, ms_per_step_gpu=e0.elapsed_time(e1)/K, host_enqueue_ms_per_step=1e3*t_enq/K, wall_ms=1e3*t_all/K)))
