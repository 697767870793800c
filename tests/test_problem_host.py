"""CPU tests of the host side: presolve/standard form (neuripsintopt_b200/problem.py) against the
reference's own standard forms (tests/golden) and against the oracle restatement on random LPs."""
import numpy as np
import pytest

from neuripsintopt_b200 import problem as PR
from oracle import hsd_numpy as O
from tests import golden_util as GU

f32 = lambda v: None if v is None else np.asarray(v, np.float32).astype(np.float64)  # noqa: E731


def _build(name, z, case, b=True):
    inp = GU.problem_inputs(name, z, 0)
    n_x = case["c"].shape[1]
    zc = PR.zero_columns(f32(inp["A"]), f32(inp["G"]), n_x)
    sgn = np.sign(case["c"][0].astype(np.float64)[zc])
    sf = PR.build_standard_form(n_x, f32(inp["A"]), f32(inp["b"]) if b else None, f32(inp["G"]), f32(inp["h"]),
                                case["kw"].get("bounds"), sgn)
    return sf, inp


@pytest.mark.parametrize("name", GU.ALL)
def test_standard_form_equals_reference(name):
    z = GU.load(name)
    case = GU.cases(z)[0]
    sf, _ = _build(name, z, case)
    np.testing.assert_array_equal(sf.A.toarray(), GU.dense(z, "stdA"))
    np.testing.assert_array_equal(sf.b, z["std_b"])
    c = case["c"][0].astype(np.float64)
    c_std = np.where(sf.c_src >= 0, c[np.maximum(sf.c_src, 0)] * sf.c_sign, 0.0)
    np.testing.assert_array_equal(c_std, z["std_c"])


def test_per_instance_b_is_a_linear_map():
    z = GU.load("twitter_sp")
    case = GU.cases(z)[0]
    sf, inp = _build("twitter_sp", z, case, b=False)
    assert sf.b is None
    bb = sf.b_std(f32(z["b_rr"]))
    assert bb.shape == (len(z["b_rr"]), sf.m)
    for i in range(len(z["b_rr"])):
        ref = O.standard_form(case["c"][i].astype(np.float64), f32(inp["A"]), f32(z["b_rr"][i]),
                              bounds=case["kw"]["bounds"])
        np.testing.assert_allclose(bb[i], ref["b"], rtol=0, atol=1e-15)
        np.testing.assert_array_equal(sf.fixed_idx, ref["fixed_idx"])
        np.testing.assert_allclose(sf.fixed_values(f32(z["b_rr"][i])), ref["fixed_val"], atol=1e-15)
    assert not sf.infeasible_mask(f32(z["b_rr"])).any()


def test_zero_column_rule_depends_on_sign_of_c():
    # ip_model_whole.py:995-1016 — Twitter's A has one unconstrained column
    z = GU.load("twitter_sp")
    case = GU.cases(z)[0]
    inp = GU.problem_inputs("twitter_sp", z, 0)
    n_x = case["c"].shape[1]
    zc = PR.zero_columns(f32(inp["A"]), None, n_x)
    assert zc.tolist() == [1852]
    pos = PR.build_standard_form(n_x, f32(inp["A"]), f32(inp["b"]), None, None, [(0.0, None)], [1.0])
    zero = PR.build_standard_form(n_x, f32(inp["A"]), f32(inp["b"]), None, None, [(0.0, None)], [0.0])
    assert pos.n + 1 == zero.n and 1852 in pos.fixed_idx and 1852 not in zero.fixed_idx
    with pytest.raises(PR.PresolveComplete) as e:  # c<0 and no upper bound: unbounded, status 3
        PR.build_standard_form(n_x, f32(inp["A"]), f32(inp["b"]), None, None, [(0.0, None)], [-1.0])
    assert e.value.status == 3


@pytest.mark.parametrize("seed", range(6))
def test_random_lps_match_oracle_restatement(seed):
    """Mixed bounds (finite, one-sided below, free), inequality + equality rows, a singleton row."""
    rng = np.random.default_rng(seed)
    n, me, mu = 9, 3, 4
    A = rng.integers(-2, 3, (me, n)).astype(float)
    A[0] = 0
    A[0, 2] = 2.0  # singleton equality row fixes x2
    G = rng.integers(-2, 3, (mu, n)).astype(float)
    G[1] = 0
    G[1, 4] = -1.0  # singleton inequality row: lower bound on x4
    b = rng.uniform(0.5, 2.0, me)
    h = rng.uniform(1.0, 3.0, mu)
    bounds = [(0, None), (0, 2.0), (0, 3), (None, 5.0), (-1.0, None), (None, None), (1.0, 4.0), (0, None), (None, None)]
    c = rng.standard_normal(n)
    ref = O.standard_form(c, A, b, G, h, bounds)
    zc = PR.zero_columns(A, G, n)
    sf = PR.build_standard_form(n, A, b, G, h, bounds, np.sign(c[zc]))
    assert not ref["complete"]
    np.testing.assert_allclose(sf.A.toarray(), ref["A"], atol=0)
    np.testing.assert_allclose(sf.b, ref["b"], atol=1e-14)
    c_std = np.where(sf.c_src >= 0, c[np.maximum(sf.c_src, 0)] * sf.c_sign, 0.0)
    np.testing.assert_allclose(c_std, ref["c"], atol=0)
    np.testing.assert_array_equal(sf.var_index, ref["var_index"])
    np.testing.assert_allclose(sf.fixed_const, ref["fixed_val"], atol=1e-15)


@pytest.mark.parametrize("seed", range(4))
def test_presolve_false_keeps_everything(seed):
    """`presolve=False` (ip_model_whole.py:1352): the same LPs as above, with a zero row, a zero column, singleton
    rows and a fixed variable left in place — standard form equal to the oracle restatement (itself checked
    against the reference's `_get_Abc` output for presolve=False on the Energy / knapsack / dense fixtures)."""
    rng = np.random.default_rng(seed)
    n, me, mu = 9, 4, 4
    A = rng.integers(-2, 3, (me, n)).astype(float)
    A[0] = 0
    A[0, 2] = 2.0
    A[3] = 0                # zero row (rhs 0)
    G = rng.integers(-2, 3, (mu, n)).astype(float)
    G[1] = 0
    G[1, 4] = -1.0
    A[:, 7] = 0
    G[:, 7] = 0             # zero column
    b = np.append(rng.uniform(0.5, 2.0, me - 1), 0.0)
    h = rng.uniform(1.0, 3.0, mu)
    bounds = [(0, None), (0, 2.0), (0, 3), (None, 5.0), (-1.0, None), (None, None), (2.0, 2.0), (0, None), (None, None)]
    c = rng.standard_normal(n)
    ref = O.standard_form(c, A, b, G, h, bounds, presolve=False)
    sf = PR.build_standard_form(n, A, b, G, h, bounds, None, presolve=False)
    assert ref["A"].shape[0] == me + mu + 3 and sf.fixed_idx.size == 0 and sf.nvar == n  # 3 finite upper bounds
    np.testing.assert_allclose(sf.A.toarray(), ref["A"], atol=0)
    np.testing.assert_allclose(sf.b, ref["b"], atol=1e-14)
    c_std = np.where(sf.c_src >= 0, c[np.maximum(sf.c_src, 0)] * sf.c_sign, 0.0)
    np.testing.assert_allclose(c_std, ref["c"], atol=0)


def test_presolve_false_on_energy_fixture():
    z = GU.load("energy_load1")
    inp = GU.problem_inputs("energy_load1", z, 0)
    sf = PR.build_standard_form(1440, f32(inp["A"]), f32(inp["b"]), f32(inp["G"]), f32(inp["h"]), [(0.0, None)], None,
                                presolve=False)
    assert sf.A.shape == (174, 1584)      # what the reference builds with presolve=False (172 x 1582 with presolve)
    ref = O.standard_form(np.ones(1440), f32(inp["A"]), f32(inp["b"]), f32(inp["G"]), f32(inp["h"]), [(0.0, None)],
                          presolve=False)
    np.testing.assert_array_equal(sf.A.toarray(), ref["A"])
    np.testing.assert_array_equal(sf.b, ref["b"])


def test_presolve_outcomes():
    with pytest.raises(PR.PresolveComplete) as e:  # zero row, non-zero rhs (:957-970)
        PR.build_standard_form(3, np.array([[0.0, 0, 0], [1, 1, 1]]), np.array([1.0, 1.0]))
    assert e.value.status == 2
    with pytest.raises(PR.PresolveComplete) as e:  # singleton row outside the bounds (:1026-1034)
        PR.build_standard_form(3, np.array([[1.0, 0, 0], [1, 1, 1]]), np.array([-1.0, 1.0]))
    assert e.value.status == 2
    sf = PR.build_standard_form(3, np.array([[1.0, 0, 0], [1, 1, 1]]), None)
    assert sf.infeasible_mask(np.array([[-1.0, 1.0], [0.5, 1.0]])).tolist() == [True, False]
    with pytest.raises(ValueError):
        PR.clean_bounds([(2.0, 1.0)], 3)
    with pytest.raises(ValueError):
        PR.build_standard_form(3, np.array([[1.0, np.nan, 0]]), np.array([1.0]))


def test_trivial_presolve_outcome_follows_the_sign_of_c():
    """No non-trivial constraint left after presolve (ip_model_whole.py:1113-1144): x = ub where c < 0, lb where c > 0,
    the finite upper bound (else 0) where c == 0.  Pinned by the reference's own output (tests/golden/bounds.npz,
    oracle/gen_golden_bounds.py): upstream returns only the variables presolve did not fix, in order."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bounds.npz"))
    A, b = z["trivial_A"].astype(np.float64), z["trivial_b"].astype(np.float64)
    tb = [tuple(r) for r in z["trivial_bounds"]]
    for c, x_ref in zip(z["trivial_c"].astype(np.float64), z["trivial_x"]):
        zc = PR.zero_columns(A, None, 4)
        with pytest.raises(PR.PresolveComplete) as e:
            PR.build_standard_form(4, A, b, None, None, tb, np.sign(c[zc]))
        x = e.value.solution(c)
        assert e.value.status == 0 and x[1] == 1.5  # the singleton row 2 x1 = 3
        with pytest.raises(PR.PresolveComplete) as e2:  # right-hand side supplied later (per-instance b): same answer
            PR.build_standard_form(4, A, None, None, None, tb, np.sign(c[zc]))
        np.testing.assert_allclose(e2.value.solution(c, b), x, atol=0)
        np.testing.assert_allclose(x[e.value.trivial[0]], x_ref.astype(np.float64), atol=0)
        lo, hi = np.array(tb).T
        assert ((x >= lo) & (x <= hi)).all()


def test_batched_A_with_number_dependent_presolve_is_split_per_instance():
    """A batch of equality matrices is analysed once on its shared pattern - unless presolve would act on the numbers
    (singleton row -> fixed variable, non-zero lower bound -> shifted b): then one function per instance is built."""
    import torch
    from neuripsintopt_b200.ipofunc import make_ipofunc
    rng = np.random.default_rng(0)
    A = rng.uniform(0.5, 2.0, (3, 1, 5))
    b = A.sum(axis=2) * 1.1
    _, hooks = make_ipofunc("whole", torch.as_tensor(A), torch.as_tensor(b), None, None, bounds=[(0.0, None)])
    assert "instances" not in hooks
    _, hooks = make_ipofunc("whole", torch.as_tensor(A), torch.as_tensor(b), None, None, bounds=[(0.25, None)])
    assert len(hooks["instances"]) == 3
    A2 = np.zeros((2, 2, 4))
    A2[:, 0, 1] = [1.0, 2.0]
    A2[:, 1, :] = 1.0
    _, hooks = make_ipofunc("alter", torch.as_tensor(A2), torch.as_tensor(np.ones((2, 2))), None, None, bounds=[(0.0, 1.0)])
    assert len(hooks["instances"]) == 2
