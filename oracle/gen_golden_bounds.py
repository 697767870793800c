"""TEST INFRASTRUCTURE — writes tests/golden/bounds.npz: x and dL/dc of the UNTOUCHED reference (through
oracle/ref_shim.py) on small LPs with mixed variable bounds — finite boxes, shifted lower bounds, upper-only bounds,
free variables — a singleton equality row (a presolve-fixed variable) and a singleton inequality row, both variants,
early-stopped and converged; and the outcome of a problem that presolve decides completely (no non-trivial
constraint left, ip_model_whole.py:1113-1144: x follows the sign of c).
Run here (needs /root/reference):  python -m oracle.gen_golden_bounds"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ref_shim  # noqa: E402
from tests import golden_util as GU  # noqa: E402

BOUNDS = [(0, None), (0, 2.0), (0, 3), (None, 5.0), (-1.0, None), (None, None), (1.0, 4.0), (0, None), (None, None)]


def mixed_lp(seed):
    """A feasible, bounded LP in the layout of tests/test_problem_host.py::test_random_lps_match_oracle_restatement."""
    rng = np.random.default_rng(seed)
    n, me, mu = 9, 3, 4
    A = rng.integers(-2, 3, (me, n)).astype(np.float32)
    A[0] = 0
    A[0, 2] = 2.0  # singleton equality row fixes x2
    G = rng.integers(-2, 3, (mu, n)).astype(np.float32)
    G[1] = 0
    G[1, 4] = -1.0  # singleton inequality row: lower bound on x4
    x0 = np.array([1.0, 1.0, 0.0, 2.0, 0.5, -0.5, 2.0, 1.0, 0.3])  # strictly inside the bounds
    b = (A.astype(np.float64) @ x0).astype(np.float32)
    x0[2] = b[0] / 2.0
    b = (A.astype(np.float64) @ x0).astype(np.float32)
    h = (G.astype(np.float64) @ x0 + rng.uniform(0.5, 1.5, mu)).astype(np.float32)
    # bounded objective: c = A^T y + G^T z + (reduced costs that push every variable towards a finite bound or the box)
    c = (rng.standard_normal(me) @ A + rng.uniform(0.1, 1.0, mu) @ G).astype(np.float64)
    c += np.array([0.5, 0.3, 0.0, -0.4, 0.6, 0.0, 0.2, 0.7, 0.0])
    return A, b, G, h, c.astype(np.float32)


def main():
    out = {}
    kept = 0
    for seed in range(40):
        A, b, G, h, c = mixed_lp(seed)
        try:
            from scipy.optimize import linprog
            lp = linprog(c.astype(np.float64), A_ub=G, b_ub=h, A_eq=A, b_eq=b, bounds=BOUNDS, method="highs")
        except Exception:
            continue
        if lp.status != 0:
            continue
        g = np.random.default_rng(100 + seed).standard_normal(9).astype(np.float32)
        ok = True
        rec = {}
        for variant in ("whole", "alter"):
            for thr in (0.1, None):
                kw = dict(bounds=BOUNDS, thr=thr, damping=1e-3, max_iter=200)
                try:
                    x, grad, cap = ref_shim.run_ipofunc(variant, c, g, A=A, b=b, G=G, h=h, **kw)
                except Exception as e:  # the reference itself fails on this instance: not a fixture
                    ok = False
                    break
                if not np.isfinite(x).all() or not np.isfinite(grad).all():
                    ok = False
                    break
                tag = "lp%d_%s_%s" % (kept, variant, "thr" if thr else "conv")
                rec[tag + "_x"], rec[tag + "_grad"], rec[tag + "_iters"] = x.astype(np.float32), grad.astype(np.float32), np.int32(cap["iters"])
            if not ok:
                break
        if not ok:
            continue
        out.update(rec)
        for k, v in (("A", A), ("b", b), ("G", G), ("h", h), ("c", c), ("g", g)):
            out["lp%d_%s" % (kept, k)] = v
        out["lp%d_highs_obj" % kept] = np.float64(lp.fun)
        kept += 1
        if kept == 3:
            break
    out["n_lps"] = np.int32(kept)
    # presolve decides everything: the only constraint is a singleton row (:1020-1041), then nothing is left (:1113-1144)
    A = np.array([[0.0, 2.0, 0.0, 0.0]], dtype=np.float32)
    b = np.array([3.0], dtype=np.float32)
    tb = [(0, 5.0), (0, 4.0), (-2.0, 6.0), (1.0, 3.0)]
    cs = np.array([[1.0, 1.0, -1.0, 0.0], [-2.0, 0.5, 0.0, 3.0], [0.0, -1.0, 2.0, -0.5]], dtype=np.float32)
    xs = []
    import torch
    W = ref_shim.load("whole")
    for c in cs:
        f = W.IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=tb, thr=0.1)
        xs.append(f(torch.tensor(c)).detach().numpy().astype(np.float32))
    out["trivial_A"], out["trivial_b"], out["trivial_c"], out["trivial_x"] = A, b, cs, np.stack(xs)
    out["trivial_bounds"] = np.array(tb, dtype=np.float64)
    np.savez_compressed(os.path.join(GU.GOLDEN, "bounds.npz"), **out)
    print("wrote", len(out), "arrays;", kept, "mixed-bounds LPs; trivial x =", np.stack(xs).tolist())


if __name__ == "__main__":
    main()
