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
        raise NotImplementedError("large row nullspace: the reference switches to its dense pivoting removal")
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


def prepare(A):
    """-> (A_rr, removed_rows): what `intopt.fit` hands to `IPOfunc` (shortespath.py:397-399)."""
    removed = redundant_rows(A)
    return np.delete(np.asarray(A), removed, axis=0), removed
