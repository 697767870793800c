"""Diagnostic: where does the end-to-end loop lose time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from neuripsintopt_b200.ip_model_whole import IPOfunc
A, b, G, h, F = bench.energy_problem()
tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)
f = IPOfunc(tt(A), tt(b), tt(G), tt(h), check_finite=False, **bench.OPTS)
c, g = bench.synthetic_costs(F, 1024, 1)
cd, gd = torch.tensor(c, device="cuda"), torch.tensor(g, device="cuda")
ch, gh = torch.tensor(c).pin_memory(), torch.tensor(g).pin_memory()
xh, gch = torch.empty_like(ch).pin_memory(), torch.empty_like(ch).pin_memory()
def timeit(name, body, n=40):
    for _ in range(5): body()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): body()
    e1.record(); tcpu = time.perf_counter() - t0; e1.synchronize()
    print("%-34s gpu %.3f ms/step   cpu enqueue %.3f ms/step" % (name, e0.elapsed_time(e1) / n, 1e3 * tcpu / n))
def resident():
    cc = cd.detach().requires_grad_(True); x = f(cc); x.backward(gd)
def fwd_only():
    with torch.no_grad(): f(cd)
def with_h2d():
    cc = ch.to("cuda", non_blocking=True).requires_grad_(True); gg = gh.to("cuda", non_blocking=True); x = f(cc); x.backward(gg)
def with_all():
    cc = ch.to("cuda", non_blocking=True).requires_grad_(True); gg = gh.to("cuda", non_blocking=True); x = f(cc); x.backward(gg)
    xh.copy_(x.detach(), non_blocking=True); gch.copy_(cc.grad, non_blocking=True)
timeit("resident fwd+bwd (no sync)", resident)
timeit("forward only (no_grad)", fwd_only)
timeit("+ h2d same stream", with_h2d)
timeit("+ h2d + d2h same stream", with_all)

# which call blocks the host?
def stamp(): return time.perf_counter()
for it in range(4):
    torch.cuda.synchronize(); t = [stamp()]
    cc = ch.to("cuda", non_blocking=True).requires_grad_(True); t.append(stamp())
    gg = gh.to("cuda", non_blocking=True); t.append(stamp())
    x = f(cc); t.append(stamp())
    x.backward(gg); t.append(stamp())
    xd = x.detach(); t.append(stamp())
    xh.copy_(xd, non_blocking=True); t.append(stamp())
    gr = cc.grad; t.append(stamp())
    gch.copy_(gr, non_blocking=True); t.append(stamp())
    names = ["h2d c", "h2d g", "forward", "backward", "detach", "d2h x", "grad", "d2h grad"]
    print("iter", it, "  ".join("%s %.2f" % (nm, 1e3 * (b - a)) for nm, a, b in zip(names, t, t[1:])), "| x", x.dtype, x.is_contiguous(), "grad", gr.dtype, gr.is_contiguous(), xh.is_pinned(), gch.is_pinned())

print("--- variants")
plain = torch.randn(1024, 1440, device="cuda")
ph = torch.empty(1024, 1440, pin_memory=True)
def t_ms(fn):
    t0 = stamp(); fn(); return 1e3 * (stamp() - t0)
for it in range(3):
    torch.cuda.synchronize()
    cc = cd.detach().requires_grad_(True); x = f(cc); x.backward(gd)
    a = t_ms(lambda: ph.copy_(plain, non_blocking=True))
    b_ = t_ms(lambda: gch.copy_(cc.grad, non_blocking=True))
    c_ = t_ms(lambda: xh.copy_(x.detach(), non_blocking=True))
    xc = x.detach().clone(); d_ = t_ms(lambda: xh.copy_(xc, non_blocking=True))
    print("plain %.2f  grad %.2f  x %.2f  x.clone %.2f   x.storage_offset %d x.stride %s numel %d" % (a, b_, c_, d_, x.storage_offset(), x.stride(), x.numel()))

print("--- staging variant")
stage_x = torch.empty(1024, 1440, device="cuda"); stage_g = torch.empty(1024, 1440, device="cuda")
def with_stage():
    cc = ch.to("cuda", non_blocking=True).requires_grad_(True); gg = gh.to("cuda", non_blocking=True); x = f(cc); x.backward(gg)
    stage_x.copy_(x.detach()); stage_g.copy_(cc.grad)
    xh.copy_(stage_x, non_blocking=True); gch.copy_(stage_g, non_blocking=True)
timeit("+ h2d + staged d2h same stream", with_stage)
for it in range(2):
    torch.cuda.synchronize()
    cc = cd.detach().requires_grad_(True); x = f(cc); x.backward(gd)
    a = t_ms(lambda: stage_x.copy_(x.detach()))
    b_ = t_ms(lambda: xh.copy_(stage_x, non_blocking=True))
    print("d2d x %.2f  d2h staged %.2f  | x.data_ptr %x is_leaf %s grad_fn %s" % (a, b_, x.data_ptr(), x.is_leaf, type(x.grad_fn).__name__))
