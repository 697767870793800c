"""TEST INFRASTRUCTURE — golden-vector generator.  Runs only in the build container.

Runs the untouched reference (via oracle/ref_shim.py) on the experiment LPs of
SURVEY.md §8(d) and writes small fixtures to tests/golden/:

    python oracle/gen_golden.py            # all configs
    python oracle/gen_golden.py energy     # one config

Every fixture holds the user-form inputs (so the GPU box needs nothing from
/root/reference), the reference's float32 outputs (x, dL/dc), the fp64 iterate
(x,y,t,tau,kappa) captured inside the reference run, its iteration count, and the
standard form the reference built.  Problem inputs that are data, not source
(ICON instance parameters, the Twitter graph's edge list, knapsack cost rows) are
stored as arrays as well.
"""
import io
import json
import os
import pickle
import sys
import warnings
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = ref_shim.REF_ROOT


def _coo(M):
    M = np.asarray(M)
    r, c = np.nonzero(M)
    return dict(shape=np.array(M.shape), row=r.astype(np.int32), col=c.astype(np.int32),
                val=M[r, c].astype(np.float32))


def _pack(prefix, d):
    return {"%s_%s" % (prefix, k): v for k, v in d.items()}


def _run_cases(name, cases, A=None, b=None, G=None, h=None, per_instance=None):
    """cases: list of dict(variant, kw, cs (k×n), gs (k×n)); returns flat dict for npz."""
    out = {}
    for ci, case in enumerate(cases):
        xs, grads, iters, states = [], [], [], []
        for i, (c, g) in enumerate(zip(case["cs"], case["gs"])):
            args = dict(A=A, b=b, G=G, h=h)
            if per_instance is not None:
                args.update(per_instance(i))
            x, dc, cap = ref_shim.run_ipofunc(case["variant"], c, g, **args, **case["kw"])
            xs.append(x)
            grads.append(dc)
            iters.append(cap["iters"])
            st = cap["state"]  # (x, y, t, tau, kappa) as _do_step returns them
            states.append(np.concatenate([st[0], st[1], st[2], np.ravel(st[3]), np.ravel(st[4])]))
            if ci == 0 and i == 0:
                out["std_A"] = cap["A_std"].astype(np.float32) if cap["A_std"].size < 600000 else np.zeros(0)
                out.update(_pack("stdA", _coo(cap["A_std"])))
                out["std_b"] = cap["b_std"]
                out["std_c"] = cap["c_std"]
        tag = "case%d" % ci
        out[tag + "_variant"] = np.array(case["variant"])
        out[tag + "_kw"] = np.array(json.dumps(case["kw"]))
        out[tag + "_c"] = np.asarray(case["cs"], dtype=np.float32)
        out[tag + "_g"] = np.asarray(case["gs"], dtype=np.float32)
        out[tag + "_x"] = np.asarray(xs, dtype=np.float32)
        out[tag + "_grad"] = np.asarray(grads, dtype=np.float32)
        out[tag + "_iters"] = np.asarray(iters, dtype=np.int32)
        out[tag + "_state"] = np.asarray(states, dtype=np.float64)
        print("  %s %s %s iters=%s" % (name, case["variant"], case["kw"], iters))
    out["n_cases"] = np.array(len(cases))
    out.pop("std_A", None)
    return out


def gen_energy(k=3):
    """Config 2: Energy ICON load1: reference builders data_reading (EnergyCost/ICON.py:5-47)
    and make_matrix_intopt (Interior/intopt_energy_mlp.py:124-168), real price days from Data.npz."""
    (data_reading,) = ref_shim.lift_functions("EnergyCost/ICON.py", ["data_reading"])
    (make_matrix_intopt,) = ref_shim.lift_functions("Interior/intopt_energy_mlp.py", ["make_matrix_intopt"])
    param = data_reading(os.path.join(REF, "EnergyCost/load1/day01.txt"))
    A, b, G, h, F = [np.asarray(v.numpy(), dtype=np.float64) for v in make_matrix_intopt(**param)]
    y = np.load(os.path.join(REF, "Data.npz"))["y_train"].reshape(-1, 48)
    y = np.maximum(y, 0.0)  # get_energy.py:27-28 clamps at 0
    rng = np.random.default_rng(0)
    P_true = y[:k]
    P_pred = P_true * np.abs(1 + 0.2 * rng.standard_normal(P_true.shape))
    cs = (P_pred @ F.T).astype(np.float32)
    gs = (P_true @ F.T / 1440.0).astype(np.float32)  # loss=(x*c_true).mean()  (:426)
    base = dict(pc=True, bounds=[(0.0, None)], max_iter=None, method=1)
    cases = [
        dict(variant="whole", kw=dict(base, thr=0.1, damping=1e-6), cs=cs, gs=gs),
        dict(variant="whole", kw=dict(base, thr=None, damping=1e-6), cs=cs[:2], gs=gs[:2]),
        dict(variant="whole", kw=dict(base, thr=1e-3, damping=1e-5), cs=cs[:2], gs=gs[:2]),
        dict(variant="alter", kw=dict(base, thr=0.1, damping=1e-6), cs=cs[:2], gs=gs[:2]),
    ]
    out = _run_cases("energy", cases, A=A, b=b, G=G, h=h)
    out.update(_pack("A", _coo(A)))
    out.update(_pack("G", _coo(G)))
    out.update(_pack("F", _coo(F)))
    out["b"], out["h"] = b.astype(np.float32), h.astype(np.float32)
    out["price_days"] = y[:64].astype(np.float32)
    out["icon_param"] = np.array(json.dumps(param))
    np.savez_compressed(os.path.join(OUT, "energy_load1.npz"), **out)


def gen_twitter(k=3):
    """Config 3: Twitter shortest path (shortespath/shortespath.py:377-449)."""
    z = zipfile.ZipFile(os.path.join(REF, "experiments/Twitter Shortest Path/data.zip"))
    name = [n for n in z.namelist() if n.endswith("Graph_exp.npy")][0]
    fobj = io.BytesIO(z.read(name))
    # exp_run.py:17-22: three arrays saved back to back: A, data (features), c
    A = np.asarray(np.load(fobj), dtype=np.float64)
    _features = np.load(fobj)
    c_true = np.asarray(np.load(fobj), dtype=np.float64).reshape(-1)
    with open(os.path.join(REF, "experiments/Twitter Shortest Path/instances.pickle"), "rb") as f:
        inst = pickle.load(f)
    W = ref_shim.load("whole")
    import contextlib
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        rows = np.asarray(W._remove_redundant_rows(A))  # shortespath.py:392
    A_rr = np.delete(A, rows, axis=0)
    pairs = [tuple(int(v) for v in inst["train"][i]) for i in range(len(inst["train"]))]
    train = pairs[:k]
    rng = np.random.default_rng(0)
    c_pred = (c_true * np.abs(1 + 0.3 * rng.standard_normal((k, c_true.size)))).astype(np.float32)
    gs = np.tile((c_true / c_true.size).astype(np.float32), (k, 1))  # loss=(y*x).mean()
    bs = np.zeros((k, A.shape[0]))
    for i, (s, d) in enumerate(train):
        bs[i, s] = 1
        bs[i, d] = -1
    bs_rr = np.delete(bs, rows, axis=1)
    base = dict(bounds=[(0.0, None)], max_iter=100, method=1)
    cases = [
        dict(variant="whole", kw=dict(base, thr=0.1), cs=c_pred, gs=gs),
        dict(variant="whole", kw=dict(base, thr=None), cs=c_pred[:2], gs=gs[:2]),
    ]
    out = _run_cases("twitter", cases, A=A_rr, per_instance=lambda i: dict(b=bs_rr[i]))
    out.update(_pack("Afull", _coo(A)))
    out["removed_rows"] = rows.astype(np.int32)
    out["b_rr"] = bs_rr.astype(np.float32)
    out["c_true"] = c_true.astype(np.float32)
    out["train_pairs"] = np.asarray(pairs, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "twitter_sp.npz"), **out)


def gen_knapsack(k=4):
    """Config 1: Building knapsack (experiments/Building Knapsack/optimizer.py:385-405, ModelRun.py:17-60).
    The cost row is passed in the (A,b) = equality slot."""
    import pandas as pd
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prop = pd.read_pickle(os.path.join(REF, "experiments/Building Knapsack/scaled_property_features.pkl"))
    train = prop.sample(n=310, random_state=10).sort_values(["Sl"], ascending=[True])
    y = train.iloc[:, 13].values.astype(np.float32)
    cost = train.iloc[:, 14].values.astype(np.float32)
    rng = np.random.default_rng(0)
    costs = cost.reshape(10, 31)[:k]
    ys = y.reshape(10, 31)[:k]
    cs = (-(ys + 0.3 * rng.standard_normal(ys.shape))).astype(np.float32)
    gs = (-ys / 31.0).astype(np.float32)  # loss=-(x*y).mean()
    hval = np.array([31 * 20.0])
    base = dict(bounds=[(0.0, 1.0)], max_iter=1000, smoothing=False)
    cases = [
        dict(variant="whole", kw=dict(base, thr=0.1, damping=0.1), cs=cs, gs=gs),
        dict(variant="whole", kw=dict(base, thr=None, damping=0.1), cs=cs[:2], gs=gs[:2]),
        dict(variant="alter", kw=dict(base, thr=0.1, damping=1e-3), cs=cs[:2], gs=gs[:2]),
    ]
    out = _run_cases("knapsack", cases, per_instance=lambda i: dict(A=costs[i][None, :], b=hval))
    out["costs"] = np.asarray(cost.reshape(10, 31), dtype=np.float32)
    out["ys"] = np.asarray(y.reshape(10, 31), dtype=np.float32)
    out["budget_rhs"] = hval
    np.savez_compressed(os.path.join(OUT, "knapsack.npz"), **out)


def dense_problem(m, B, seed=None):
    """Config 5 generator: feasible & bounded dense LP, equality-only."""
    rng = np.random.default_rng(m if seed is None else seed)
    n = 2 * m
    A = rng.standard_normal((m, n)).astype(np.float32).astype(np.float64)
    x0 = rng.uniform(0.1, 1.0, n)
    b = (A @ x0).astype(np.float32).astype(np.float64)
    y0 = rng.standard_normal((B, m))
    s0 = rng.uniform(0.1, 1.0, (B, n))
    c = (y0 @ A + s0).astype(np.float32)
    return A, b, c


def gen_dense(ms=(16, 64), k=2):
    for m in ms:
        A, b, cs = dense_problem(m, k)
        gs = np.ones_like(cs)
        base = dict(bounds=[(0.0, None)], max_iter=1000)
        cases = [
            dict(variant="whole", kw=dict(base, thr=0.1, damping=1e-6), cs=cs, gs=gs),
            dict(variant="whole", kw=dict(base, thr=None, damping=1e-6), cs=cs, gs=gs),
            dict(variant="alter", kw=dict(base, thr=0.1, damping=1e-6), cs=cs, gs=gs),
        ]
        out = _run_cases("dense%d" % m, cases, A=A, b=b)
        out["A"] = A.astype(np.float32)
        out["b"] = b.astype(np.float32)
        np.savez_compressed(os.path.join(OUT, "dense_m%d.npz" % m), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["energy", "twitter", "knapsack", "dense"]
    for w in which:
        print("generating", w)
        {"energy": gen_energy, "twitter": gen_twitter, "knapsack": gen_knapsack, "dense": gen_dense}[w]()
