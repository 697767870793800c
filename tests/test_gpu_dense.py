"""GPU parity tests of the lock-step dense path (csrc/intopt_dense.cu: batched DMMA assembly, tiled Cholesky, blocked
substitution, per-instance masked iteration) against the CPU oracle, the reference's own golden outputs and the
persistent-CTA path.  Tolerances are BASELINE.json's: 1e-6 on x* / objective, 1e-5 on dL/dc, equal iteration counts for
early-stopped runs."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import hsd_numpy as O
from tests import golden_util as GU
from tests.test_gpu_parity import _check_case, _rel2, _relinf, _tt

pytestmark = pytest.mark.gpu


def gen_dense(m, B, seed=None):
    """feasible and bounded dense LP, n = 2m, costs c = A^T y0 + s0."""
    rng = np.random.default_rng(m if seed is None else seed)
    n = 2 * m
    A = rng.standard_normal((m, n)).astype(np.float32)
    x0 = rng.uniform(0.1, 1.0, n)
    b = (A.astype(np.float64) @ x0).astype(np.float32)
    y0 = rng.standard_normal((B, m))
    s0 = rng.uniform(0.1, 1.0, (B, n))
    c = (y0 @ A.astype(np.float64) + s0).astype(np.float32)
    return A, b, c


@pytest.fixture
def dense_path(monkeypatch):
    from neuripsintopt_b200 import ipofunc
    monkeypatch.setenv("INTOPT_PATH", "dense")
    ipofunc.clear_cache()
    yield
    ipofunc.clear_cache()


def _handle(A, b=None):
    from neuripsintopt_b200 import _capi
    L = _capi.lib()
    A = sp.csr_matrix(A)
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    val = np.ascontiguousarray(A.data, dtype=np.float64)
    h = ctypes.c_void_p()
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    _capi.check(L.intopt_problem_create(ctypes.byref(h), A.shape[0], A.shape[1], rp.ctypes.data, ci.ctypes.data,
                                        val.ctypes.data, None if bb is None else bb.ctypes.data, 0), "create")
    info = _capi.IntoptInfo()
    _capi.check(L.intopt_problem_info(h, ctypes.byref(info)), "info")
    return L, h, info.as_dict()


@pytest.mark.parametrize("m,n", [(40, 100), (64, 128), (200, 330), (256, 512)])
def test_normal_solve_matches_numpy(dense_path, m, n):
    """Assembly (DMMA SYRK) + tiled Cholesky + blocked substitution against numpy, including shapes that need padding
    in both dimensions and weights spanning twelve orders of magnitude (d = x / t near convergence)."""
    rng = np.random.default_rng(m + n)
    A = rng.standard_normal((m, n))
    L, h, info = _handle(A)
    assert info["path"] == 1
    B = 5
    d = np.exp(rng.uniform(-14, 14, (B, n)))
    rhs = rng.standard_normal((B, m))
    delta = 1e-6
    dd = torch.tensor(d, device="cuda")
    rr = torch.tensor(rhs, device="cuda")
    fail = torch.zeros(B, dtype=torch.int32, device="cuda")
    nbytes = L.intopt_workspace_bytes(h, B)
    ws = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device="cuda")
    from neuripsintopt_b200 import _capi
    _capi.check(L.intopt_dense_normal_solve(h, B, dd.data_ptr(), delta, rr.data_ptr(), fail.data_ptr(), ws.data_ptr(),
                                            ws.numel() * 8, None), "normal_solve")
    torch.cuda.synchronize()
    z = rr.cpu().numpy()
    assert int(fail.sum()) == 0
    for i in range(B):
        M = (A * d[i]) @ A.T + delta * np.eye(m)
        ref = np.linalg.solve(M, rhs[i])
        # backward-stable solve: the residual is at rounding level relative to |M| |z|
        res = np.abs(M @ z[i] - rhs[i]).max() / (np.abs(M).max() * np.abs(z[i]).max() * m)
        err, cond = _rel2(z[i], ref), np.linalg.cond(M)
        print("normal_solve m=%d n=%d inst %d: scaled residual %.2e, error vs numpy %.2e, cond %.1e" % (m, n, i, res, err, cond))
        assert res < 1e-13, (m, n, i, res)
        assert err < 1e-13 * cond + 1e-12, (m, n, i, err, cond)
    L.intopt_problem_destroy(h)


@pytest.mark.parametrize("name", ["dense_m16", "dense_m64"])
def test_golden_cases_on_the_dense_path(dense_path, name):
    """The reference's own outputs (tests/golden) and the oracle, early-stopped and converged, whole and alter."""
    z = GU.load(name)
    inp = GU.problem_inputs(name, z, 0)
    for case in GU.cases(z):
        _check_case(name, z, case, inp, lambda i: inp)


@pytest.mark.parametrize("m", [128, 200, 256])
@pytest.mark.parametrize("thr", [0.1, None])
def test_config5_against_oracle(m, thr):
    """Config 5 shapes that take the dense path by default (m >= 128), incl. one that is not a multiple of the tile."""
    from neuripsintopt_b200 import ipofunc
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    ipofunc.clear_cache()
    B = 6
    A, b, c = gen_dense(m, B)
    f = IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=thr,
                damping=1e-6)
    ct = torch.tensor(c, device="cuda", requires_grad=True)
    x = f(ct)
    x.sum().backward()
    torch.cuda.synchronize()
    status, iters = IPOfunc.status()
    info = [v.info for v in ipofunc._PROBLEM_CACHE.values() if hasattr(v, "info")][-1]
    assert info["path"] == 1
    assert int((status != 0).sum()) == 0
    for i in range(3):
        r = O.ipofunc(c[i], np.ones(2 * m, np.float32), A=A, b=b, bounds=[(0.0, None)], thr=thr, damping=1e-6)
        xi = x[i].detach().double().cpu().numpy()
        c64 = c[i].astype(np.float64)
        assert abs(c64 @ xi - c64 @ r["x"]) <= 1e-6 * max(1.0, abs(c64 @ r["x"])), (m, thr, i)
        if thr:
            assert int(iters[i]) == r["iters"], (int(iters[i]), r["iters"])
            assert _relinf(xi, r["x"]) < 1e-6, (m, thr, i, _relinf(xi, r["x"]))
            assert _rel2(ct.grad[i].double().cpu().numpy(), r["grad"]) < 1e-5
        else:
            assert abs(int(iters[i]) - r["iters"]) <= 1
            assert _relinf(xi, r["x"]) < 1e-5, (m, thr, i, _relinf(xi, r["x"]))


def test_dense_and_cta_paths_agree(monkeypatch):
    """The same dense problem (m = 96) through both paths: different kernels, same LP layer."""
    from neuripsintopt_b200 import ipofunc
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    A, b, c = gen_dense(96, 9)
    outs = []
    for path in ("cta", "dense"):
        monkeypatch.setenv("INTOPT_PATH", path)
        ipofunc.clear_cache()
        f = IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=0.1,
                    damping=1e-6)
        ct = torch.tensor(c, device="cuda", requires_grad=True)
        x = f(ct)
        (x * torch.arange(1, 193, device="cuda")).sum().backward()
        torch.cuda.synchronize()
        status, iters = IPOfunc.status()
        outs.append((x.detach().cpu().numpy(), ct.grad.cpu().numpy(), iters.cpu().numpy()))
    ipofunc.clear_cache()
    assert (outs[0][2] == outs[1][2]).all()
    assert _relinf(outs[0][0], outs[1][0]) < 1e-6
    assert _rel2(outs[0][1], outs[1][1]) < 1e-5


def test_dense_path_per_instance_b_and_warm_start(dense_path):
    """Per-instance right-hand sides and init_val (end_vectors of a previous solve) on the dense path."""
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    m, B = 24, 4
    A, b, c = gen_dense(m, B, seed=5)
    rng = np.random.default_rng(1)
    bb = np.stack([(A.astype(np.float64) @ rng.uniform(0.1, 1.0, 2 * m)).astype(np.float32) for _ in range(B)])
    f = IPOfunc(torch.tensor(A), torch.tensor(bb), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=0.1,
                damping=1e-6)
    ct = torch.tensor(c, device="cuda", requires_grad=True)
    x = f(ct)
    x.sum().backward()
    torch.cuda.synchronize()
    for i in range(B):
        r = O.ipofunc(c[i], np.ones(2 * m, np.float32), A=A, b=bb[i], bounds=[(0.0, None)], thr=0.1, damping=1e-6)
        assert _relinf(x[i].detach().double().cpu().numpy(), r["x"]) < 1e-6
        assert _rel2(ct.grad[i].double().cpu().numpy(), r["grad"]) < 1e-5
    warm = IPOfunc.end_vectors()
    f2 = IPOfunc(torch.tensor(A), torch.tensor(bb), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=1e-3,
                 damping=1e-6, init_val=warm)
    x2 = f2(ct.detach())
    f3 = IPOfunc(torch.tensor(A), torch.tensor(bb), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=1e-3,
                 damping=1e-6)
    x3 = f3(ct.detach())
    torch.cuda.synchronize()
    # both end below the same mu threshold: objectives agree to the threshold's accuracy
    o2 = (x2.double() * ct.detach().double()).sum(1)
    o3 = (x3.double() * ct.detach().double()).sum(1)
    assert torch.allclose(o2, o3, rtol=5e-2, atol=5e-2)


@pytest.mark.parametrize("m", [512, 1024])
def test_large_dense_against_oracle(m):
    """Config 5 at m = 512 and 1024 (n = 2m): early-stopped forward + backward of a small batch against the oracle —
    1e-6 on x, 1e-5 on dL/dc, equal iteration counts."""
    from neuripsintopt_b200 import ipofunc
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    ipofunc.clear_cache()
    B = 3
    A, b, c = gen_dense(m, B)
    f = IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=0.1,
                damping=1e-6)
    ct = torch.tensor(c, device="cuda", requires_grad=True)
    x = f(ct)
    x.sum().backward()
    torch.cuda.synchronize()
    status, iters = IPOfunc.status()
    assert int((status != 0).sum()) == 0
    r = O.ipofunc(c[0], np.ones(2 * m, np.float32), A=A, b=b, bounds=[(0.0, None)], thr=0.1, damping=1e-6)
    assert int(iters[0]) == r["iters"]
    ex = _relinf(x[0].detach().double().cpu().numpy(), r["x"])
    eg = _rel2(ct.grad[0].double().cpu().numpy(), r["grad"])
    print("dense m=%d: x %.2e grad %.2e iters %d" % (m, ex, eg, r["iters"]))
    assert ex < 1e-6 and eg < 1e-5, (m, ex, eg)
    ipofunc.clear_cache()


def test_m2048_smoke():
    """m = 2048, n = 4096 (the largest shape of the sweep; the gather-list path refused anything beyond m ~ 600): the
    normal-equations primitive against numpy, and one early-stopped forward + backward that ends with status 0 and a
    finite gradient."""
    from neuripsintopt_b200 import _capi, ipofunc
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    ipofunc.clear_cache()
    m, n, B = 2048, 4096, 2
    A, b, c = gen_dense(m, B)
    L, h, info = _handle(A.astype(np.float64))
    assert info["path"] == 1
    rng = np.random.default_rng(0)
    d = np.exp(rng.uniform(-3, 3, (1, n)))
    rhs = rng.standard_normal((1, m))
    dd, rr = torch.tensor(d, device="cuda"), torch.tensor(rhs, device="cuda")
    ws = torch.empty(L.intopt_workspace_bytes(h, 1) // 8 + 1, dtype=torch.float64, device="cuda")
    _capi.check(L.intopt_dense_normal_solve(h, 1, dd.data_ptr(), 1e-6, rr.data_ptr(), None, ws.data_ptr(), ws.numel() * 8,
                                            None), "normal_solve")
    torch.cuda.synchronize()
    Md = torch.tensor(A.astype(np.float64), device="cuda")
    M = (Md * dd) @ Md.T + 1e-6 * torch.eye(m, dtype=torch.float64, device="cuda")
    res = (M @ rr[0] - torch.tensor(rhs[0], device="cuda")).abs().max() / (M.abs().max() * rr[0].abs().max() * m)
    assert float(res) < 1e-13, float(res)
    L.intopt_problem_destroy(h)
    del ws, M, Md
    f = IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=0.1,
                damping=1e-6)
    ct = torch.tensor(c, device="cuda", requires_grad=True)
    x = f(ct)
    x.sum().backward()
    torch.cuda.synchronize()
    status, iters = IPOfunc.status()
    assert status.tolist() == [0, 0] and 2 <= int(iters.min()) and int(iters.max()) <= 6
    assert torch.isfinite(x).all() and torch.isfinite(ct.grad).all() and float(ct.grad.abs().max()) > 0
    ipofunc.clear_cache()


def test_dense_path_failure_ladder(dense_path):
    """_get_delta's failure ladder (ip_model_whole.py:798-822) on the lock-step path: a rank-deficient dense A with a
    slightly negative damping makes every factorisation hit a non-positive pivot; the two retry rungs (diagonal raised by
    1e-12, 1e-8 x its bound) recover it and the run ends like one with a small positive damping.  A damping that no rung
    can repair ends with status 3 and the iterate unchanged."""
    from neuripsintopt_b200 import ip_model_alter
    m, B = 24, 5
    A, b, c = gen_dense(m, B, seed=11)
    A[m - 1] = A[0] + A[1]                      # dependent row: A D A^T is singular for every D
    b[m - 1] = b[0] + b[1]
    c = (np.random.default_rng(2).standard_normal((B, m)) @ A.astype(np.float64) + np.random.default_rng(3).uniform(0.1, 1.0, (B, 2 * m))).astype(np.float32)
    ct = torch.tensor(c, device="cuda")
    outs = {}
    for damping in (-1e-9, 1e-6, -1e3):
        f = ip_model_alter.IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)],
                                   thr=0.1, damping=damping, max_iter=50)
        x = f(ct)
        torch.cuda.synchronize()
        status, iters = ip_model_alter.IPOfunc.status()
        outs[damping] = (x.double().cpu().numpy(), status.cpu().numpy(), iters.cpu().numpy())
    x_neg, st_neg, it_neg = outs[-1e-9]
    x_pos, st_pos, it_pos = outs[1e-6]
    assert (st_pos == 0).all() and (st_neg == 0).all(), (st_pos, st_neg)
    assert (it_neg == it_pos).all(), (it_neg, it_pos)
    err = max(_relinf(x_neg[i], x_pos[i]) for i in range(B))
    print("dense ladder: x vs the damping=1e-6 run, rel-inf %.2e" % err)
    assert err < 1e-3, err
    x_bad, st_bad, it_bad = outs[-1e3]
    assert (st_bad == 3).all() and np.isfinite(x_bad).all() and (it_bad == 1).all()
