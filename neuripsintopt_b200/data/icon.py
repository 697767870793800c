"""ICON energy-scheduling instances: parser, LP builder and a synthetic-instance generator.

Host-side problem setup for the Energy experiment:
  * `data_reading` parses the ICON text format like the reference's EnergyCost/ICON.py:5-47;
  * `make_matrix_intopt` builds (A, b, G, h, F) with the structure of
    Interior/intopt_energy_mlp.py:124-168, including its quirks (capacity rows get coefficient 1, not the
    resource use U[f][r], :155; the E/L window rows index machine 0's block only, :164-165 — SURVEY C.10),
    but as scipy sparse matrices built from index ranges instead of dense torch loops;
  * `synthesize` draws larger instances in the same format (BASELINE.json config 4).
`load1_day01.json` holds the parameters of the one instance the reference ships
(EnergyCost/load1/day01.txt) so benchmarks need nothing from the reference tree.
"""
from __future__ import annotations

import json
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))


def load1_day01():
    with open(os.path.join(_HERE, "load1_day01.json")) as f:
        return json.load(f)


def data_reading(filename):
    """ICON text format -> parameter dict (same keys as EnergyCost/ICON.py:36-47)."""
    with open(filename) as f:
        lines = f.read().splitlines()
    q, n_res, n_mach = int(lines[0]), int(lines[1]), int(lines[2])
    idle, up, down, MC = [], [], [], []
    for m in range(n_mach):
        head = lines[3 + 2 * m].split()
        idle.append(int(head[1]))
        up.append(float(head[2]))
        down.append(float(head[3]))
        MC.append([int(v) for v in lines[4 + 2 * m].split()])
    pos = 3 + 2 * n_mach
    n_tasks = int(lines[pos])
    U, D, E, L, P = [], [], [], [], []
    for f_ in range(n_tasks):
        head = lines[pos + 1 + 2 * f_].split()
        D.append(int(head[1]))
        E.append(int(head[2]))
        L.append(int(head[3]))
        P.append(float(head[4]))
        U.append([int(v) for v in lines[pos + 2 + 2 * f_].split()])
    return {"nbMachines": n_mach, "nbTasks": n_tasks, "nbResources": n_res, "MC": MC, "U": U, "D": D, "E": E,
            "L": L, "P": P, "idle": idle, "up": up, "down": down, "q": q}


def write_icon(param, filename):
    """Inverse of data_reading."""
    out = [str(param["q"]), str(param["nbResources"]), str(param["nbMachines"])]
    for m in range(param["nbMachines"]):
        out.append("%d %d %s %s" % (m, param["idle"][m], param["up"][m], param["down"][m]))
        out.append(" ".join(str(v) for v in param["MC"][m]))
    out.append(str(param["nbTasks"]))
    for f_ in range(param["nbTasks"]):
        out.append("%d %d %d %d %s" % (f_, param["D"][f_], param["E"][f_], param["L"][f_], param["P"][f_]))
        out.append(" ".join(str(v) for v in param["U"][f_]))
    with open(filename, "w") as f:
        f.write("\n".join(out) + "\n")


def make_matrix_intopt(nbMachines, nbTasks, nbResources, MC, U, D, E, L, P, idle, up, down, q, **_):
    """(A, b, G, h, F): A (3T x n) equalities, G (M*N x n) capacity rows, F (n x N) maps a price vector of N
    slots to per-variable costs `c = F @ price`; variable (task f, machine m, start slot s) sits at column
    (f*M + m)*N + s.  float32-representable values, returned as scipy CSR (+ dense 1-D b, h)."""
    N = 1440 // q
    M, T = nbMachines, nbTasks
    n = T * M * N
    g_r, g_c, f_r, f_c, f_v = [], [], [], [], []
    for m in range(M):
        for t in range(N):
            for f_ in range(T):
                base = (f_ * M + m) * N
                lo, hi = base + max(0, t - D[f_] + 1), base + t + 1
                cols = np.arange(lo, hi)
                g_r.append(np.full(cols.size, t + m * N))
                g_c.append(cols)
                f_r.append(np.full(cols.size, t))
                f_c.append(cols)
                f_v.append(np.full(cols.size, P[f_]))
    g_r, g_c = np.concatenate(g_r), np.concatenate(g_c)
    G = sp.csr_matrix((np.ones(g_r.size), (g_r, g_c)), shape=(M * N, n))
    G.data[:] = 1.0  # duplicates cannot occur, but keep the 0/1 contract explicit
    h = np.array([float(MC[m][0]) for m in range(M) for _ in range(N)])
    # F[t, col] is *assigned* (not accumulated) per machine in the reference; rows differ per m only in columns
    fr, fc, fv = np.concatenate(f_r), np.concatenate(f_c), np.concatenate(f_v)
    Ft = sp.csr_matrix((fv, (fr, fc)), shape=(N, n))
    a_r, a_c = [], []
    for f_ in range(T):
        blk = f_ * N * M
        for row, (lo, hi) in ((f_, (blk, blk + N * M)), (T + f_, (blk, blk + E[f_])),
                              (2 * T + f_, (blk + L[f_] - D[f_] + 1, blk + N * M))):
            cols = np.arange(lo, max(lo, hi))
            a_r.append(np.full(cols.size, row))
            a_c.append(cols)
    a_r, a_c = np.concatenate(a_r), np.concatenate(a_c)
    A = sp.csr_matrix((np.ones(a_r.size), (a_r, a_c)), shape=(3 * T, n))
    b = np.concatenate([np.ones(T), np.zeros(2 * T)])
    F = sp.csr_matrix(Ft.T.astype(np.float32).astype(np.float64))
    return A, b, G, h.astype(np.float32).astype(np.float64), F


def synthesize(n_machines, n_tasks, q=30, seed=0):
    """A random instance in the ICON format with the value ranges of load1."""
    rng = np.random.default_rng(seed)
    N = 1440 // q
    D = rng.integers(5, 31, n_tasks)
    E = rng.integers(1, 21, n_tasks)
    L = np.minimum(N, E + D + rng.integers(0, N, n_tasks))
    P = np.round(rng.uniform(100, 750, n_tasks), 2)
    U = rng.integers(500, 4200, n_tasks)
    MC = rng.integers(5400, 5600, n_machines)
    return {"nbMachines": int(n_machines), "nbTasks": int(n_tasks), "nbResources": 1,
            "MC": [[int(v)] for v in MC], "U": [[int(v)] for v in U], "D": [int(v) for v in D],
            "E": [int(v) for v in E], "L": [int(v) for v in L], "P": [float(v) for v in P],
            "idle": [int(v) for v in rng.integers(100, 200, n_machines)], "up": [0.1] * n_machines,
            "down": [0.1] * n_machines, "q": int(q)}
