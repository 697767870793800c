"""Regret evaluation without Gurobi: the reference scores a predictor by solving each test
instance with the PREDICTED costs in Gurobi and pricing that decision at the TRUE costs
(`Gurobi_ICON.solve_model`, EnergyCost/ICON.py:156-319; `get_loss`, shortespath/shortespath.py:80-102;
`val_loss`, experiments/Building Knapsack/optimizer.py:416-447).  gurobipy is not installable offline, so the
same integer programs go through HiGHS (`scipy.optimize.milp`).  Host-side, evaluation only: nothing here is
on the training hot path, and nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp

from .data import icon


def solve_program(c, A_eq=None, b_eq=None, A_ub=None, b_ub=None, lb=0.0, ub=1.0, integral=True, time_limit=None):
    """min c.x  s.t.  A_eq x = b_eq, A_ub x <= b_ub, lb <= x <= ub, x integer when `integral`.
    Returns (x, status) with status 0 = optimal, 1 = limit reached with an incumbent, else x is None."""
    c = np.asarray(c, dtype=np.float64)
    cons = []
    if A_eq is not None and A_eq.shape[0]:
        b_eq = np.asarray(b_eq, dtype=np.float64)
        cons.append(LinearConstraint(sp.csr_matrix(A_eq), b_eq, b_eq))
    if A_ub is not None and A_ub.shape[0]:
        cons.append(LinearConstraint(sp.csr_matrix(A_ub), -np.inf, np.asarray(b_ub, dtype=np.float64)))
    opts = {} if time_limit is None else {"time_limit": float(time_limit)}
    res = milp(c, constraints=cons, bounds=Bounds(lb, ub), integrality=np.full(c.size, 1 if integral else 0),
               options=opts)
    if res.x is None:
        return None, int(res.status)
    return (np.round(res.x) if integral else res.x), int(res.status)


# ---------------------------------------------------------------------------------------------- Energy (ICON)
def icon_program(nbMachines, nbTasks, nbResources, MC, U, D, E, L, P, q, **_):
    """The scheduling program `Gurobi_ICON.make_model` states (ICON.py:90-141), as sparse matrices over the
    variable order of `make_matrix_intopt` (task f, machine m, start slot t -> (f*M + m)*N + t):
    every task starts exactly once (:130); no start before E[f] or after L[f]-D[f] on ANY machine (:126-127,
    returned as per-variable upper bounds); for every resource r, machine m and slot t the running tasks use at
    most MC[m][r], each weighted by its use U[f][r] (:133-137).  Unlike the LP layer's matrices (which keep the
    training code's simplifications) this is the program regret is measured on."""
    N = 1440 // q
    M, T = nbMachines, nbTasks
    n = T * M * N
    ub = np.ones(n)
    a_r, a_c = [], []
    for f in range(T):
        for m in range(M):
            base = (f * M + m) * N
            ub[base:base + E[f]] = 0.0
            ub[base + max(0, L[f] - D[f] + 1):base + N] = 0.0
        a_r.append(np.full(M * N, f))
        a_c.append(np.arange(f * M * N, (f + 1) * M * N))
    A = sp.csr_matrix((np.ones(T * M * N), (np.concatenate(a_r), np.concatenate(a_c))), shape=(T, n))
    g_r, g_c, g_v, h = [], [], [], []
    row = 0
    for r in range(nbResources):
        for m in range(M):
            for t in range(N):
                for f in range(T):
                    base = (f * M + m) * N
                    cols = np.arange(base + max(0, t - D[f] + 1), base + t + 1)
                    g_r.append(np.full(cols.size, row))
                    g_c.append(cols)
                    g_v.append(np.full(cols.size, float(U[f][r])))
                h.append(float(MC[m][r]))
                row += 1
    G = sp.csr_matrix((np.concatenate(g_v), (np.concatenate(g_r), np.concatenate(g_c))), shape=(row, n))
    return A, np.ones(T), G, np.array(h), ub


class EnergySchedule:
    """`Gurobi_ICON` (ICON.py:48-319) on HiGHS: binary start indicators x[f, m, t]; objective
    sum x[f,m,t] * P[f] * q/60 * sum(price[t:t+D[f]])  (:183-184); `solve` returns the energy drawn per slot
    (`solver`, :256-258) like `solve_model` does."""

    def __init__(self, param, relax=False, time_limit=None):
        self.A, self.b, self.G, self.h, self.ub = icon_program(**param)
        self.F = icon.make_matrix_intopt(**param)[4]          # (n, N): P[f] on the slots a start occupies
        self.scale = param["q"] / 60.0
        self.relax = relax
        self.time_limit = time_limit

    def solve(self, price):
        c = (self.F @ np.asarray(price, dtype=np.float64)) * self.scale
        x, status = solve_program(c, self.A, self.b, self.G, self.h, 0.0, self.ub, integral=not self.relax,
                                  time_limit=self.time_limit)
        if x is None:
            raise RuntimeError("ICON schedule: no solution (HiGHS status %d)" % status)
        return x, (self.F.T @ x) * self.scale

    def regret(self, price_pred, price_true):
        """`sum(V_target * (sol - sol_target))` (intopt_energy_mlp.py:224-225): the bill, at the true prices, of
        the schedule chosen for the predicted prices, minus the best bill."""
        price_true = np.asarray(price_true, dtype=np.float64)
        _, load_pred = self.solve(price_pred)
        _, load_true = self.solve(price_true)
        return float(price_true @ (load_pred - load_true))


# ------------------------------------------------------------------------------------------ shortest path
def path_rhs(n_nodes, source, destination):
    b = np.zeros(n_nodes)
    b[source] = 1.0
    b[destination] = -1.0
    return b


def shortest_path_cost(A, c_decide, c_price, pairs):
    """Sum over (source, destination) pairs of  c_price . x*(c_decide)  with x* the binary min-cost path of
    `A x = b` (get_loss, shortespath.py:80-102).  A is the FULL node-arc incidence matrix.  Infeasible pairs are
    skipped, as the reference does (it only prints them)."""
    A = sp.csr_matrix(A)
    total, skipped = 0.0, 0
    for s, d in pairs:
        x, _ = solve_program(c_decide, A, path_rhs(A.shape[0], s, d), lb=0.0, ub=1.0, integral=True)
        if x is None:
            skipped += 1
            continue
        total += float(np.asarray(c_price) @ x)
    return total, skipped


def shortest_path_regret(A, c_pred, c_true, pairs):
    """`test_rslt[1] - test_obj` of the Twitter experiment (exp_run.py:30-50, :75)."""
    got, _ = shortest_path_cost(A, c_pred, c_true, pairs)
    best, _ = shortest_path_cost(A, c_true, c_true, pairs)
    return got - best


# ----------------------------------------------------------------------------------------------- knapsack
def knapsack_value(prices_decide, prices_value, costs, budget=20.0, integral=True, equality=True):
    """Per minibatch: value at `prices_value` of  argmax prices_decide.x  s.t.  cost.x (=|<=) items*budget,
    x binary (val_loss, optimizer.py:433-441 — the reference states the budget as an EQUALITY).  Rows for which
    the program is infeasible give NaN."""
    prices_decide = np.atleast_2d(np.asarray(prices_decide, dtype=np.float64))
    prices_value = np.atleast_2d(np.asarray(prices_value, dtype=np.float64))
    costs = np.atleast_2d(np.asarray(costs, dtype=np.float64))
    out = np.full(costs.shape[0], np.nan)
    for k in range(costs.shape[0]):
        row = sp.csr_matrix(costs[k][None, :])
        rhs = np.array([costs.shape[1] * float(budget)])
        x, _ = solve_program(-prices_decide[k], row if equality else None, rhs if equality else None,
                             None if equality else row, None if equality else rhs, 0.0, 1.0, integral=integral)
        if x is not None:
            out[k] = float(prices_value[k] @ x)
    return out


def knapsack_regret(prices_pred, prices_true, costs, budget=20.0, integral=True, equality=True):
    """`test_obj - obj_list` (optimizer.py:447): best achievable value minus the value of the predicted choice."""
    best = knapsack_value(prices_true, prices_true, costs, budget, integral, equality)
    got = knapsack_value(prices_pred, prices_true, costs, budget, integral, equality)
    return best - got
