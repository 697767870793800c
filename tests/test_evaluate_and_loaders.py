"""CPU tests of the §8 f-3 / f-4 host pieces: data loaders (ICON, Twitter graph + row removal, knapsack frames)
and the Gurobi-free regret evaluation.  Reference data files are read only when /root/reference is present
(this container); everything else runs from committed fixtures and small synthetic cases."""
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from neuripsintopt_b200 import evaluate as EV
from neuripsintopt_b200.data import icon, knapsack, twitter

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference/experiments"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference data files not present")


def _golden_twitter():
    z = np.load(os.path.join(GOLD, "twitter_sp.npz"), allow_pickle=True)
    A = sp.csr_matrix((z["Afull_val"].astype(np.float64), (z["Afull_row"], z["Afull_col"])),
                      shape=tuple(z["Afull_shape"])).toarray()
    return z, A


def test_row_removal_matches_reference_fixture():
    z, A = _golden_twitter()
    assert twitter.redundant_rows(A) == [int(v) for v in z["removed_rows"]]
    A_rr, removed = twitter.prepare(A)
    assert A_rr.shape == (A.shape[0] - len(removed), A.shape[1])


def test_row_removal_full_rank_and_small_cases():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(5, 9))
    assert twitter.redundant_rows(A) == []
    B = np.vstack([A, A[1] + A[3]])                      # one dependent row
    removed = twitter.redundant_rows(B)
    assert len(removed) == 1
    assert np.linalg.matrix_rank(np.delete(B, removed, axis=0)) == 5
    assert twitter.redundant_rows(np.zeros((0, 4))) == []


def test_large_row_nullspace_takes_the_pivoting_removal():
    """More than five dependent rows: upstream switches to its pivoting removal (Andersen 1995; its own call into it
    raises).  The restatement must find exactly the rows SciPy's implementation of the same algorithm finds,
    dense and sparse, count zero rows as dependent and flag an inconsistent right-hand side."""
    import warnings
    from scipy.optimize._remove_redundancy import _remove_redundancy_pivot_dense
    rng = np.random.default_rng(0)
    for trial in range(4):
        m, n, r = 30, 45, 21 - 2 * trial
        base = rng.standard_normal((r, n))
        A = np.vstack([base, rng.standard_normal((m - r, r)) @ base])[rng.permutation(m)]
        if trial % 2:
            A[int(rng.integers(m))] = 0.0
        b = A @ rng.standard_normal(n)
        d = twitter.dependent_rows_pivot(A, b)
        assert d == twitter.dependent_rows_pivot(sp.csr_matrix(A), b) == twitter.redundant_rows(A)
        keep = [i for i in range(m) if i not in d]
        assert len(d) == m - np.linalg.matrix_rank(A) and np.linalg.matrix_rank(A[keep]) == len(keep)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            A2, _, status, _ = _remove_redundancy_pivot_dense(A.copy(), b.copy())
        assert status == 0 and np.array_equal(A2, A[keep])
    b_bad = b.copy()
    b_bad[d[0]] += 1.0
    with pytest.raises(ValueError):
        twitter.dependent_rows_pivot(A, b_bad)


def test_graph_file_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    A, X, c = rng.integers(-1, 2, (4, 6)).astype(float), rng.normal(size=(6, 3)), rng.uniform(size=6)
    path = tmp_path / "Graph_exp.npy"
    with open(path, "wb") as f:
        np.save(f, A), np.save(f, X), np.save(f, c)
    A2, X2, c2 = twitter.load_graph(str(path))
    assert np.array_equal(A, A2) and np.array_equal(X, X2) and np.array_equal(c, c2)
    ipath = tmp_path / "instances.pickle"
    with open(ipath, "wb") as f:
        pickle.dump({"train": {1: (np.int64(2), np.int64(0)), 0: (np.int64(3), np.int64(1))}, "test": {}}, f)
    assert twitter.load_instances(str(ipath)) == {"train": [(3, 1), (2, 0)], "test": []}


@needs_ref
def test_twitter_files_of_the_reference():
    A, X, c = twitter.load_graph(REF + "/Twitter Shortest Path/data.zip")
    assert A.shape == (231, 2861) and X.shape == (2861, 1050) and c.shape == (2861,)
    z, A_gold = _golden_twitter()
    assert np.array_equal(A, A_gold)
    ins = twitter.load_instances(REF + "/Twitter Shortest Path/instances.pickle")
    assert {k: len(v) for k, v in ins.items()} == {"train": 80, "validation": 15, "test": 20}
    assert np.array_equal(np.asarray(ins["train"]), z["train_pairs"])


@needs_ref
def test_knapsack_frames_of_the_reference():
    prop, econ = knapsack.load_frames(REF + "/Building Knapsack/scaled_property_features.pkl",
                                      REF + "/Building Knapsack/scaled_economic_features.pkl")
    tr_p, tr_e, te_p, te_e = knapsack.split(prop, econ)
    assert len(tr_p) == 310 and len(te_p) == 62 and len(tr_e) == 5 * 310
    mb = knapsack.minibatches(knapsack.arrays(tr_p, tr_e))
    assert mb["cost"].shape == (10, 31) and mb["x_ts"].shape == (10, 31, 5, 19)
    assert mb["cost"].min() == 20.0 and mb["cost"].max() == 1000.0
    up = knapsack.arrays(tr_p, tr_e, upstream_indexing=True)["x_ts"]
    assert up.shape == (310, 5, 19) and not up[1].any() and up[5].any()


def test_knapsack_minibatches_drop_the_ragged_tail():
    arr = {"y": np.arange(70, dtype=np.float32), "cost": np.ones(70, np.float32), "x": np.zeros((70, 3))}
    mb = knapsack.minibatches(arr, items=31)
    assert mb["y"].shape == (2, 31) and mb["x"].shape == (2, 31, 3) and mb["y"][1, 0] == 31


def test_icon_program_vs_lp_layer_matrices():
    """On load1 every start on machine m>0 outside the window is ruled out by the scheduling program but not by
    the LP layer's rows; the once-per-task rows coincide."""
    p = icon.load1_day01()
    A, b, G, h, ub = EV.icon_program(**p)
    A_lp = icon.make_matrix_intopt(**p)[0]
    T = p["nbTasks"]
    assert (A != A_lp[:T]).nnz == 0 and np.all(b == 1)
    assert G.shape == (p["nbResources"] * p["nbMachines"] * 48, A.shape[1])
    assert set(np.unique(G.data)) <= {float(u[0]) for u in p["U"]}
    f, m, N = 1, 2, 48
    base = (f * p["nbMachines"] + m) * N
    assert not ub[base:base + p["E"][f]].any() and ub[base + p["E"][f]] == 1.0


def test_energy_schedule_and_regret():
    p = icon.load1_day01()
    rng = np.random.default_rng(0)
    price = rng.uniform(20, 120, 48)
    es = EV.EnergySchedule(p)
    x, load = es.solve(price)
    assert set(np.unique(x)) <= {0.0, 1.0} and x.sum() == p["nbTasks"]
    assert np.all(es.G @ x <= es.h + 1e-9) and np.allclose(es.A @ x, 1.0) and not x[es.ub == 0].any()
    bill = float(load @ price)
    assert abs(bill - float(x @ (es.F @ price)) * es.scale) < 1e-6 * bill
    relaxed = EV.EnergySchedule(p, relax=True).solve(price)[1] @ price
    assert relaxed <= bill * (1 + 1e-9)
    assert es.regret(price, price) == 0.0
    assert es.regret(price[::-1].copy(), price) >= 0.0


def test_shortest_path_regret_small_graph():
    # 0 -> 1 -> 3 (cost 1+1) and 0 -> 2 -> 3 (cost 1+5), arcs: 01, 13, 02, 23
    A = np.array([[1, 0, 1, 0], [-1, 1, 0, 0], [0, 0, -1, 1], [0, -1, 0, -1]], float)
    c_true = np.array([1.0, 1.0, 1.0, 5.0])
    total, skipped = EV.shortest_path_cost(A, c_true, c_true, [(0, 3)])
    assert total == 2.0 and skipped == 0
    assert EV.shortest_path_regret(A, np.array([3.0, 3.0, 1.0, 1.0]), c_true, [(0, 3)]) == 4.0
    assert EV.shortest_path_regret(A, c_true * 2, c_true, [(0, 3), (0, 1)]) == 0.0
    assert EV.shortest_path_cost(A, c_true, c_true, [(3, 0)]) == (0.0, 1)      # no path back: skipped


def test_shortest_path_binary_program_equals_lp_on_fixture_graph():
    z, A = _golden_twitter()
    c = z["c_true"].astype(np.float64)
    pairs = [tuple(int(v) for v in p) for p in z["train_pairs"][:3]]
    total, skipped = EV.shortest_path_cost(A, c, c, pairs)
    lp = 0.0
    for s, d in pairs:
        r = linprog(c, A_eq=sp.csr_matrix(A), b_eq=EV.path_rhs(A.shape[0], s, d), bounds=(0, 1), method="highs")
        lp += r.fun
    assert skipped == 0 and abs(total - lp) <= 1e-7 * abs(lp)


def test_knapsack_regret():
    costs = np.array([[2.0, 3.0, 5.0, 4.0]])
    true = np.array([[3.0, 1.0, 4.0, 1.0]])
    # budget*items = 7: feasible subsets {0,2} (value 7), {1,3} (value 2)
    assert EV.knapsack_value(true, true, costs, budget=7 / 4)[0] == 7.0
    bad = np.array([[0.0, 5.0, 0.0, 5.0]])
    assert EV.knapsack_regret(bad, true, costs, budget=7 / 4)[0] == 5.0
    assert EV.knapsack_regret(true, true, costs, budget=7 / 4)[0] == 0.0
    assert np.isnan(EV.knapsack_value(true, true, costs, budget=0.25)[0])            # cost.x = 1 impossible
    assert EV.knapsack_value(true, true, costs, budget=2.0, equality=False)[0] == 7.0   # cost.x <= 8


def test_energy_loader_clamps_and_splits(tmp_path):
    """data/energy.py: `get_energy` on a Data.npz-style file clamps negative prices (get_energy.py:27-28) and
    `experiment_split` reproduces exp_run.py:29-39."""
    from neuripsintopt_b200.data import energy
    rng = np.random.default_rng(0)
    n_tr, n_te = 48 * 6, 48 * 70
    Xtr = np.column_stack([np.repeat(np.arange(6), 48), rng.standard_normal((n_tr, 8))])
    Xte = np.column_stack([np.repeat(np.arange(70), 48), rng.standard_normal((n_te, 8))])
    ytr = rng.uniform(-50, 200, n_tr)
    yte = rng.uniform(5, 300, n_te)
    f = tmp_path / "Data.npz"
    np.savez(f, X_1gtrain=Xtr, y_train=ytr, X_1gtest=Xte, y_test=yte)
    a, b, c, d = energy.get_energy(str(f))
    assert a.shape == Xtr.shape and (b >= 0).all() and np.array_equal(b, np.maximum(ytr, 0)) and (ytr < 0).any()
    Xt, yt, Xv, yv, Xs, ys = energy.experiment_split(a, b, c, d)
    assert Xt.shape == (n_tr, 8) and Xv.shape == (2880, 8) and Xs.shape == (n_te - 2880, 8)
    assert np.array_equal(yv, yte[:2880]) and np.array_equal(ys, yte[2880:])
    X, y = energy.synthetic_days(3)
    assert X.shape == (144, 8) and y.min() >= 20 and y.max() <= 120


def test_fit_loop_visits_every_minibatch_once_per_epoch():
    """IntoptEnergyLearner.fit on CPU with a surrogate LP layer: n_batches = T // (batchsize * 48) minibatches per epoch
    (intopt_energy_mlp.py:386), shuffled order reproducible by seed, scaler fitted on the training features."""
    import torch
    from neuripsintopt_b200.data import energy, icon
    from neuripsintopt_b200.energy_learner import IntoptEnergyLearner
    torch.manual_seed(0)
    surrogate = lambda c: torch.softmax(-c / c.abs().mean(dim=1, keepdim=True), dim=1)  # noqa: E731
    X, y = energy.synthetic_days(7, n_features=4, seed=3)
    runs = []
    for _ in range(2):
        torch.manual_seed(0)
        lrn = IntoptEnergyLearner(icon.load1_day01(), input_size=4, device="cpu", lp_layer=surrogate,
                                  optimizer=torch.optim.SGD, lr=1e-4)
        hist = lrn.fit(X, y, epochs=2, batchsize=2, seed=11)
        runs.append(hist)
        assert len(hist) == 2 * 3  # 7 days -> 3 whole minibatches of 2 days, last day dropped like upstream
        for e in (0, 1):
            assert sorted(h["first_day"] for h in hist if h["epoch"] == e) == [0, 2, 4]
        assert abs(lrn.scaler.transform(X).mean()) < 1e-12
        assert lrn.predict(X).shape == (7 * 48,)
    assert [h["first_day"] for h in runs[0]] == [h["first_day"] for h in runs[1]]
    assert [h["loss"] for h in runs[0]] == [h["loss"] for h in runs[1]]
