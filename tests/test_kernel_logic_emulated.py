"""CPU tests of the solver's device code compiled as single-threaded C++ (tests/host_emu): the same
source the CUDA build uses (intopt_solver.cuh + intopt_structure.hpp), checked against the oracle on the
reference's golden problems.  Catches arithmetic/indexing bugs without a GPU; the GPU parity tests
(-m gpu) remain the parity proof of the shipped path."""
import numpy as np
import pytest

from oracle import hsd_numpy as O
from tests import golden_util as GU
from tests.host_emu import emu_py


@pytest.mark.parametrize("name", GU.ALL)
@pytest.mark.parametrize("order", [-1, 0, 1])
def test_emulated_kernel_matches_oracle(name, order):
    z = GU.load(name)
    for case in GU.cases(z):
        kw = case["kw"]
        thr = kw.get("thr") or 0.0
        d = kw.get("damping", 1e-2)
        inp = GU.problem_inputs(name, z, 0)
        r = O.ipofunc(case["c"][0], case["g"][0], variant=case["variant"], bounds=kw.get("bounds"),
                      thr=kw.get("thr"), max_iter=kw.get("max_iter", 1000), damping=d, **inp)
        sf = r["sf"]
        m, n = sf["A"].shape
        g = case["g"][0].astype(np.float64)[sf["var_index"]]
        whole = case["variant"] == "whole"
        e = emu_py.solve(sf["A"], sf["b"], sf["c"], g=g, thr=thr, max_iter=kw.get("max_iter") or 1000,
                         damping_fwd=1e-6 if whole else d, mode=0 if whole else 1, damping_bwd=d,
                         nvar=sf["nvar"], force_order=order)
        xh = r["state"]["x"] / r["state"]["tau"]
        assert np.abs(e["xhat"][0] - xh).max() / max(1.0, np.abs(xh).max()) < 1e-6
        assert abs(int(e["iters"][0]) - r["iters"]) <= (0 if thr > 0 else 1)
        st = GU.split_state(e["state"][0], m, n)
        fn = O.backward_whole if whole else O.backward_alter
        dc = fn(sf["A"], sf["b"], sf["c"], st, g, d, sf["nvar"])
        assert np.linalg.norm(e["gc"][0] - dc) / np.linalg.norm(dc) < 1e-6


def test_emulated_per_instance_b_and_shared_c():
    z = GU.load("twitter_sp")
    case = GU.cases(z)[0]
    inp = GU.problem_inputs("twitter_sp", z, 0)
    f32 = lambda v: np.asarray(v, np.float32).astype(np.float64)  # noqa: E731
    sfs = [O.standard_form(case["c"][0].astype(np.float64), f32(inp["A"]), f32(z["b_rr"][i]),
                           bounds=[(0.0, None)]) for i in range(3)]
    bb = np.stack([s["b"] for s in sfs])
    e = emu_py.solve(sfs[0]["A"], None, sfs[0]["c"], b_batch=bb, thr=0.1, max_iter=100)
    for i in range(3):
        r = O.forward(sfs[i]["A"], sfs[i]["b"], sfs[i]["c"], thr=0.1, max_iter=100)
        assert np.abs(e["xhat"][i] - r["x"] / r["tau"]).max() < 1e-8
        assert e["iters"][i] == r["iters"]


def test_breakdown_is_reported_not_hidden():
    # an indefinite "normal matrix" cannot arise from A D A^T with D>0; force failure with NaN costs instead
    A = np.array([[1.0, 1.0, 1.0]])
    e = emu_py.solve(A, np.array([1.0]), np.array([1.0, np.nan, 2.0]))
    assert e["status"][0] == 4


def test_emulated_per_instance_values_knapsack():
    """Per-instance values of A (same pattern): the knapsack driver's equality row is the minibatch's cost vector
    (experiments/Building Knapsack/optimizer.py:396-405); every instance is checked against its own oracle solve."""
    import scipy.sparse as sp
    z = GU.load("knapsack")
    case = GU.cases(z)[0]
    kw = case["kw"]
    B = min(4, case["c"].shape[0], z["costs"].shape[0])
    sfs, refs = [], []
    for i in range(B):
        inp = GU.problem_inputs("knapsack", z, i)
        r = O.ipofunc(case["c"][i], case["g"][i], variant="whole", bounds=kw["bounds"], thr=kw["thr"],
                      max_iter=kw["max_iter"], damping=kw["damping"], **inp)
        refs.append(r)
        sfs.append(r["sf"])
    A0 = sp.csr_matrix(sfs[0]["A"])
    A0.sort_indices()
    vals = []
    for s in sfs:  # same pattern, different numbers
        Ai = sp.csr_matrix(s["A"])
        Ai.sort_indices()
        assert np.array_equal(Ai.indptr, A0.indptr) and np.array_equal(Ai.indices, A0.indices)
        vals.append(Ai.data)
    cc = np.stack([s["c"] for s in sfs])
    gg = np.stack([case["g"][i].astype(np.float64)[sfs[i]["var_index"]] for i in range(B)])
    e = emu_py.solve(sfs[0]["A"], sfs[0]["b"], cc, g=gg, thr=kw["thr"], max_iter=kw["max_iter"], damping_fwd=1e-6,
                     mode=0, damping_bwd=kw["damping"], nvar=sfs[0]["nvar"], a_vals=np.stack(vals))
    for i in range(B):
        r = refs[i]
        xh = r["state"]["x"] / r["state"]["tau"]
        assert np.abs(e["xhat"][i] - xh).max() / max(1.0, np.abs(xh).max()) < 1e-6
        assert int(e["iters"][i]) == r["iters"]
        m, n = sfs[i]["A"].shape
        st = GU.split_state(e["state"][i], m, n)
        dc = O.backward_whole(sfs[i]["A"], sfs[i]["b"], sfs[i]["c"], st, gg[i], kw["damping"], sfs[i]["nvar"])
        assert np.linalg.norm(e["gc"][i] - dc) / np.linalg.norm(dc) < 1e-6


@pytest.mark.parametrize("max_iter", [1000, 3])
def test_emulated_infeasible_unbounded_and_capped_runs(max_iter):
    """Same stopping iteration, status word and last iterate as the oracle for an infeasible LP, an unbounded LP
    and a run cut off by max_iter (which executes max_iter + 1 iterations, ip_model_whole.py:1414)."""
    cases = [(np.array([[1., 1., 0.], [0., 1., 1.]]), np.array([-1., 2.]), np.array([1., 2., 0.5])),
             (np.array([[1., -1., 0.], [0., 1., -1.]]), np.array([0., 0.]), np.array([-1., 0., 0.])),
             (np.array([[1., 1., 1., 0.], [1., -1., 0., 1.]]), np.array([4., 1.]), np.array([-1., -2., 0., 0.]))]
    seen = set()
    for A, b, c in cases:
        r = O.forward(A, b, c, thr=0.0, max_iter=max_iter)
        e = emu_py.solve(A, b, c, thr=0.0, max_iter=max_iter)
        assert e["iters"][0] == r["iters"] and e["status"][0] == r["status"]
        xo = r["x"] / r["tau"]
        assert np.abs(e["xhat"][0] - xo).max() <= 1e-9 * max(1.0, np.abs(xo).max())
        seen.add(int(r["status"]))
    assert seen == ({0, 2} if max_iter == 1000 else {1, 2})


@pytest.mark.parametrize("name", ["knapsack", "energy_load1", "dense_m16"])
@pytest.mark.parametrize("thr", [0.1, 0.0])
def test_emulated_plain_path_following(name, thr):
    """pc=False (IntoptOpts.plain = 1): same iteration count and iterate as the oracle, early-stopped and converged."""
    z = GU.load(name)
    inp = GU.problem_inputs(name, z, 0)
    case = GU.cases(z)[0]
    kw = case["kw"]
    mi = kw.get("max_iter") or 1000
    r = O.ipofunc(case["c"][0], case["g"][0], bounds=kw.get("bounds"), thr=thr, max_iter=mi, damping=kw["damping"],
                  pc=False, **inp)
    sf = r["sf"]
    g = case["g"][0].astype(np.float64)[sf["var_index"]]
    e = emu_py.solve(sf["A"], sf["b"], sf["c"], g=g, thr=thr, max_iter=mi, damping_bwd=kw["damping"], nvar=sf["nvar"],
                     plain=1, beta=0.1)
    assert e["iters"][0] == r["iters"] and e["status"][0] == 0
    xo = r["state"]["x"] / r["state"]["tau"]
    assert np.abs(e["xhat"][0] - xo).max() <= 1e-6 * max(1.0, np.abs(xo).max())
    if thr > 0:
        assert np.linalg.norm(e["gc"][0] - r["grad"][sf["var_index"]]) <= 1e-5 * np.linalg.norm(r["grad"])
