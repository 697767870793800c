"""Twitter shortest-path instances: the graph file, the instance file and the row removal the
reference's driver performs before it builds `IPOfunc`.

  * `load_graph` reads `Graph_exp.npy` — three arrays saved back to back: node-arc incidence matrix A (nodes x
    edges, +1 tail / -1 head), edge features, true edge costs (experiments/Twitter Shortest Path/exp_run.py:17-22)
    — from the file itself or from the `data.zip` that ships it;
  * `load_instances` reads `instances.pickle`: {'train' | 'validation' | 'test': {k: (source, destination)}}
    (exp_run.py:24-28);
  * `redundant_rows` restates `_remove_redundant_rows` (shortespath/shortespath.py:36-78) for the small-nullspace
    branch the experiment takes: rank test, then the SVD loop of Interior/remove_redundancy.py:360-415.  Like the
    reference, the indices it returns refer to the SHRINKING matrix of each pass but are afterwards applied to
    the original (shortespath.py:398-399, :410), so the result can stay rank-deficient (rows [0, 112]
    for the shipped graph, rank 228 of 229) — kept, because it decides which LP is trained on.
"""
from __future__ import annotations

import io
import pickle
import zipfile

import numpy as np


def _open(path, member):
    if str(path).endswith(".zip"):
        with zipfile.ZipFile(path) as z:
            return io.BytesIO(z.read(member))
    return open(path, "rb")


def load_graph(path):
    """-> (A, features, c)."""
    with _open(path, "Graph_exp.npy") as f:
        A = np.load(f)
        features = np.load(f)
        c = np.load(f)
    return A, features, c


def load_instances(path):
    """-> {'train': [(s, d), ...], 'validation': [...], 'test': [...]} in key order."""
    with _open(path, "instances.pickle") as f:
        raw = pickle.load(f)
    return {split: [(int(s), int(d)) for _, (s, d) in sorted(items.items())] for split, items in raw.items()}


def redundant_rows(A_eq, small_nullspace=5):
    """Row indices the reference deletes from `A_eq` and from every right-hand side."""
    A = np.array(A_eq, dtype=np.float64)
    if A.size == 0:
        return []
    rank = np.linalg.matrix_rank(A)
    if rank >= A.shape[0]:
        return []
    if A.shape[0] - rank > small_nullspace:
        # upstream switches to its pivoting removal here (ip_model_whole.py:1192-1194) - and raises a TypeError, because
        # that function still has SciPy's (A, rhs) signature (remove_redundancy.py:107).  What it was meant
        # to return is restated in dependent_rows_pivot: the dependent rows in the ORIGINAL numbering.
        return dependent_rows_pivot(A)
    eps = np.finfo(float).eps
    removed = []
    U, s, _ = np.linalg.svd(A)
    tol = s.max() * max(A.shape) * eps          # fixed from the first decomposition (:399)
    s_min = s[-1] if A.shape[0] <= A.shape[1] else 0.0
    while abs(s_min) < tol:
        eligible = np.abs(U[:, -1]) > tol * 10e6
        counts = (np.abs(A) > 1e-13).sum(axis=1)
        i = int(np.argmax(counts * eligible))   # densest eligible row, first on ties (:34-54)
        A = np.delete(A, i, axis=0)
        removed.append(i)
        U, s, _ = np.linalg.svd(A)
        s_min = s[-1] if A.shape[0] <= A.shape[1] else 0.0
    return removed


def dependent_rows_pivot(A_eq, rhs=None, tolapiv=1e-8, tolprimal=1e-8):
    """Linearly dependent rows of `A_eq` by the basis-pivoting method of Andersen (1995) that
    `remove_redundancy.py:107-230` (dense, Bartels-Golub updates) and `:233-355` (sparse LU) implement: start from the
    basis of artificial columns [I | A]; for row i, pi = B^-T e_i; if some structural column outside the basis has
    |A_j . pi| > tolapiv it replaces artificial column i, otherwise row i is a combination of the others - dependent, or,
    if pi . rhs is not ~0, the system is inconsistent (upstream: status 2).  Accepts dense or scipy.sparse input (the
    sparse branch factors the basis with SuperLU, the dense one with LAPACK; the basis is refactored per row instead of
    updated, which only costs time on a host-side one-off).  Returns the list of dependent row indices; zero rows count
    as dependent (upstream removes them first, :139).  Raises ValueError for an inconsistent system."""
    import scipy.linalg as sla
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    sparse = sp.issparse(A_eq)
    A = sp.csc_matrix(A_eq, dtype=np.float64) if sparse else np.array(A_eq, dtype=np.float64)
    m, n = A.shape
    if m == 0:
        return []
    nz_row = np.asarray((abs(A) > 1e-13).sum(axis=1)).ravel() > 0
    rhs_v = None if rhs is None else np.asarray(rhs, dtype=np.float64).ravel()
    if rhs_v is not None and np.any(~nz_row & (np.abs(rhs_v) > tolprimal)):
        raise ValueError("a zero row of A_eq has a non-zero right-hand side: infeasible")
    rows = np.flatnonzero(nz_row)                # the algorithm runs on the non-zero rows
    Ar = A[rows, :]
    mr = rows.size
    basis = list(range(mr))                      # columns of [I | Ar]; < mr: artificial
    in_basis = np.zeros(mr + n, dtype=bool)
    in_basis[:mr] = True
    dependent = [int(i) for i in np.flatnonzero(~nz_row)]

    def column(j):
        if j < mr:
            e = np.zeros(mr)
            e[j] = 1.0
            return e
        col = Ar[:, j - mr]
        return np.asarray(col.todense()).ravel() if sparse else col

    bnorm = 0.0 if rhs_v is None else float(np.linalg.norm(rhs_v[rows]))
    for i in range(mr):
        Bm = np.column_stack([column(j) for j in basis])
        e = np.zeros(mr)
        e[i] = 1.0
        if sparse:
            pi = spla.splu(sp.csc_matrix(Bm)).solve(e, trans="T")
        else:
            pi = sla.lu_solve(sla.lu_factor(Bm), e, trans=1)
        score = np.abs(np.asarray(Ar.T @ pi).ravel())
        score[in_basis[mr:]] = 0.0
        if score.size and score.max() > tolapiv:
            j = mr + int(np.argmax(score))       # a very independent column takes the place of artificial column i
            in_basis[basis[i]] = False
            basis[i] = j
            in_basis[j] = True
        else:
            if rhs_v is not None and abs(float(pi @ rhs_v[rows])) / (1.0 + bnorm) > tolprimal:
                raise ValueError("a combination of rows of A_eq vanishes while the same combination of b_eq does not: "
                                 "infeasible")
            dependent.append(int(rows[i]))
    return sorted(dependent)


def prepare(A):
    """-> (A_rr, removed_rows): what `intopt.fit` hands to `IPOfunc` (shortespath.py:397-399)."""
    removed = redundant_rows(A)
    return np.delete(np.asarray(A), removed, axis=0), removed
