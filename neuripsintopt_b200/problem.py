"""Host-side problem setup: presolve + standard form, done ONCE per (A, b, G, h, bounds).

Mirrors the behaviour of the reference's host prep — `_clean_inputs` (Interior/ip_model_whole.py
:67-329), `_presolve` (:918-1161, single pass, rule order kept) and `_get_Abc` (:333-523: bound rows
first, slack columns last) — which the reference re-runs inside every `forward` call (:1324-1380;
the Energy driver re-creates IPOfunc per instance, Interior/intopt_energy_mlp.py:422).  Here the result
is a cached `StandardForm`: a sparse standard-form A, linear maps that turn the user's `c` and
(optionally per-instance) `b_eq` into the standard-form `c`/`b`, and the index maps that put the
solution and gradient back (`_postprocess` :1241-1285, backward :1518,1559-1560).

The only part of presolve that depends on the cost vector is the zero-column rule (:995-1016, a
variable without any constraint is fixed at a bound according to sign(c)); `zero_columns` exposes those
columns so the caller can group a batch by sign pattern.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

TOL = 1e-9  # presolve tolerance (:918)


class PresolveComplete(Exception):
    """Presolve decided the problem (:964-970,1003-1011,1026-1034,1113-1144)."""

    def __init__(self, status, x, trivial=None, fixed=None):
        super().__init__("presolve complete, status %d" % status)
        self.status = status
        self.x = x
        # (fixed_idx, const, map): values of the presolve-fixed variables = const + map @ b_eq (singleton rows, :1020-1041)
        self.fixed = fixed
        # (var_index, lb, ub) of the variables that are left when no non-trivial constraint is (:1113-1144): their values
        # follow the sign of c — x = ub where c < 0, lb where c > 0, the upper bound (or 0 if infinite) where c == 0
        self.trivial = trivial

    def solution(self, c, b_eq=None):
        """x (n_x) for the cost vector c (numpy): the presolve values, with the sign rule of :1136-1142 applied to the
        variables no constraint is left for.  (Upstream returns only those variables, a vector of their number, from
        its `complete` branch :1483-1488; here every variable is reported, like every other outcome does.)"""
        x = np.array(self.x, dtype=np.float64)
        if self.fixed is not None and b_eq is not None and self.fixed[0].size:
            x[self.fixed[0]] = self.fixed[1] + self.fixed[2] @ np.asarray(b_eq, dtype=np.float64)
        if self.trivial is not None:
            vi, lb, ub = self.trivial
            cv = np.asarray(c, dtype=np.float64)[vi]
            zero_c = np.where(np.isinf(ub), 0.0, ub)
            x[vi] = np.where(cv < 0, ub, np.where(cv > 0, lb, zero_c))
        return x


def raw_bounds_count(bounds, n_x):
    """Number of leading variables upstream's post-solve loop visits: the length of the caller's bounds list when it is
    a list of pairs (ip_model_whole.py:1263), every variable for the other accepted forms."""
    if bounds is None or isinstance(bounds, str) or len(bounds) == 0:
        return n_x
    try:
        if all(len(bd) == 2 for bd in bounds):
            return min(len(bounds), n_x)
    except TypeError:
        pass
    return n_x


def clean_bounds(bounds, n_x):
    """`_clean_inputs` bounds handling (:256-317) -> (lb, ub) float arrays with ±inf for None."""
    if isinstance(bounds, str):
        raise TypeError("bounds must be a sequence of (min, max) pairs")
    if bounds is None or len(bounds) == 0:
        bounds = [(0, None)] * n_x
    elif len(bounds) == 1:
        if len(bounds[0]) != 2:
            raise ValueError("exactly one lower and one upper bound must be specified for each element of x")
        bounds = [tuple(bounds[0])] * n_x
    elif len(bounds) == n_x:
        try:
            len(bounds[0])
        except TypeError:
            bounds = [(bounds[0], bounds[1])] * n_x
    elif len(bounds) == 2 and np.isreal(bounds[0]) and np.isreal(bounds[1]):
        bounds = [(bounds[0], bounds[1])] * n_x
    else:
        raise ValueError("exactly one lower and one upper bound must be specified for each element of x")
    lb = np.empty(n_x)
    ub = np.empty(n_x)
    for i, bd in enumerate(bounds):
        if len(bd) != 2:
            raise ValueError("bound %d %s: need (lb, ub)" % (i, bd))
        lo, hi = bd
        if lo is not None and hi is not None and lo > hi:
            raise ValueError("bound %d %s: a lower bound must be <= the upper bound" % (i, bd))
        if lo == np.inf:
            raise ValueError("bound %d %s: infinity is not a valid lower bound" % (i, bd))
        if hi == -np.inf:
            raise ValueError("bound %d %s: negative infinity is not a valid upper bound" % (i, bd))
        lb[i] = -np.inf if lo is None or lo == -np.inf else float(lo)
        ub[i] = np.inf if hi is None or hi == np.inf else float(hi)
    return lb, ub


def _as_matrix(M, n_x, name):
    if M is None or np.size(M) == 0:
        return sp.csr_matrix((0, n_x), dtype=np.float64)
    M = sp.csr_matrix(M, dtype=np.float64) if sp.issparse(M) else sp.csr_matrix(np.asarray(M, dtype=np.float64))
    if M.ndim != 2 or M.shape[1] != n_x:
        raise ValueError("%s must be 2-D with %d columns" % (name, n_x))
    if not np.isfinite(M.data).all():
        raise ValueError("%s must not contain inf, nan or None" % name)
    M.eliminate_zeros()
    M.sort_indices()
    return M


def _as_vector(v, size, name):
    if v is None or np.size(v) == 0:
        v = np.zeros(0)
    v = np.asarray(v, dtype=np.float64)
    if v.ndim == 2 and v.shape[-1] == size:  # per-instance right-hand sides, B x size
        pass
    else:
        v = v.reshape(-1)
        if v.shape != (size,):
            raise ValueError("%s must have %d entries" % (name, size))
    if not np.isfinite(v).all():
        raise ValueError("%s must not contain inf, nan or None" % name)
    return v


@dataclass
class StandardForm:
    n_x: int
    A: sp.csr_matrix            # m x n standard-form matrix
    m: int
    n: int
    nvar: int                   # leading columns that are the kept user variables
    var_index: np.ndarray       # user index of each kept variable (ascending)
    c_src: np.ndarray           # n: user index feeding column j of c_std, -1 for slacks
    c_sign: np.ndarray          # n: +-1 (0 for slacks)
    fixed_idx: np.ndarray       # user indices removed by presolve (`undo[0]`)
    fixed_const: np.ndarray     # their values when b_eq is the problem's own (n_fixed)
    fixed_from_beq: sp.csr_matrix  # n_fixed x n_eq: fixed = fixed_const0 + fixed_from_beq @ b_eq
    fixed_const0: np.ndarray
    b_const: np.ndarray         # m: b_std = b_const + b_from_beq @ b_eq
    b_from_beq: sp.csr_matrix   # m x n_eq
    b: np.ndarray | None        # standard-form b for the problem's own b_eq (None if it had none)
    zero_rows_eq: np.ndarray    # equality rows without entries: |b_eq| must be <= tol (:957-970)
    singleton: tuple            # (rows, cols, coef, lb_at_check, ub_at_check) of singleton equality rows
    x_shift: np.ndarray         # nvar: lower-bound shift added back in postsolve (:1274)
    x_flip_ub: np.ndarray       # nvar: nan, or ub for variables substituted x = ub - x' (:1272)
    free_cols: np.ndarray       # standard-form column of the negative part of each free variable, -1 if none
    c0_const: float = 0.0
    n_eq_user: int = 0
    meta: dict = field(default_factory=dict)

    def b_std(self, b_eq=None):
        """Standard-form right-hand side(s); `b_eq` may be (n_eq,) or (B, n_eq)."""
        if b_eq is None:
            if self.b is None:
                raise ValueError("the problem was built without b; pass b per call")
            return self.b
        b_eq = np.asarray(b_eq, dtype=np.float64)
        out = self.b_from_beq @ b_eq.reshape(-1, self.n_eq_user).T  # m x B
        out = (out + self.b_const[:, None]).T
        return out[0] if b_eq.ndim == 1 else out

    def fixed_values(self, b_eq=None):
        if b_eq is None:
            return self.fixed_const
        b_eq = np.asarray(b_eq, dtype=np.float64)
        out = (self.fixed_from_beq @ b_eq.reshape(-1, self.n_eq_user).T + self.fixed_const0[:, None]).T
        return out[0] if b_eq.ndim == 1 else out

    def infeasible_mask(self, b_eq):
        """Per-instance presolve infeasibility (status 2): zero equality row with non-zero rhs (:959-970) or a
        singleton equality row whose value violates the bounds (:1026-1034)."""
        b_eq = np.asarray(b_eq, dtype=np.float64).reshape(-1, self.n_eq_user)
        bad = np.zeros(b_eq.shape[0], dtype=bool)
        if self.zero_rows_eq.size:
            bad |= (np.abs(b_eq[:, self.zero_rows_eq]) > TOL).any(axis=1)
        rows, cols, coef, lo, hi = self.singleton
        if rows.size:
            val = b_eq[:, rows] / coef
            bad |= ((val < lo - TOL) | (val > hi + TOL)).any(axis=1)
        return bad


def zero_columns(A_eq, A_ub, n_x):
    """User columns with no entry in any constraint (after zero-row removal nothing changes this)."""
    A_eq = _as_matrix(A_eq, n_x, "A_eq")
    A_ub = _as_matrix(A_ub, n_x, "A_ub")
    nnz = np.asarray((A_eq != 0).sum(axis=0)).ravel() + np.asarray((A_ub != 0).sum(axis=0)).ravel()
    if A_eq.shape[0] + A_ub.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    return np.nonzero(nnz == 0)[0]


def build_standard_form(n_x, A_eq=None, b_eq=None, A_ub=None, b_ub=None, bounds=None, zero_col_sign=None,
                        presolve=True):
    """Presolve + standard form for every cost vector whose sign pattern on the zero columns is
    `zero_col_sign` (array over `zero_columns(...)`, entries -1/0/+1; ignored if there are none).

    `b_eq` may be None (per-instance right-hand sides are supplied later through `b_std`).
    Raises PresolveComplete when presolve decides the problem independently of `c`'s values.
    `presolve=False` (ip_model_whole.py:1352): no row / column is removed and no variable fixed; the standard form
    is `_get_Abc` of the inputs as given.
    """
    A_eq = _as_matrix(A_eq, n_x, "A_eq")
    A_ub = _as_matrix(A_ub, n_x, "A_ub")
    n_eq, n_ubr = A_eq.shape[0], A_ub.shape[0]
    have_beq = b_eq is not None and np.size(b_eq) > 0
    b_eq_v = _as_vector(b_eq, n_eq, "b_eq") if (have_beq or n_eq == 0) else None
    b_ub_v = _as_vector(b_ub, n_ubr, "b_ub")
    if b_ub_v.ndim != 1:
        raise ValueError("per-instance b_ub (h) is not supported; only b_eq may vary per instance")
    if b_eq_v is not None and b_eq_v.ndim != 1:
        raise ValueError("build_standard_form takes a single b_eq; pass batches to StandardForm.b_std")
    lb, ub = clean_bounds(bounds, n_x)
    lb_raw, ub_raw = lb.copy(), ub.copy()  # the caller's bounds: what upstream's post-solve adjusts with (:1263-1274)
    x_pre = np.zeros(n_x)

    # b_eq enters linearly: track every right-hand side as (const, sparse map of b_eq)
    I_eq = sp.identity(n_eq, format="csr")

    # -- zero rows (:957-991) ---------------------------------------------------------------------
    eq_nnz = np.diff(A_eq.indptr)
    ub_nnz = np.diff(A_ub.indptr)
    if not presolve:
        eq_nnz, ub_nnz = np.ones(n_eq, dtype=np.int64), np.ones(n_ubr, dtype=np.int64)  # every row is kept
    zero_rows_eq = np.nonzero(eq_nnz == 0)[0]
    if zero_rows_eq.size and b_eq_v is not None and (np.abs(b_eq_v[zero_rows_eq]) > TOL).any():
        raise PresolveComplete(2, x_pre)
    eq_keep = np.nonzero(eq_nnz != 0)[0]
    if ((ub_nnz == 0) & (b_ub_v < -TOL)).any():
        raise PresolveComplete(2, x_pre)
    ub_keep = np.nonzero(ub_nnz != 0)[0]

    # -- zero columns (:995-1016): depends on sign(c) ------------------------------------------------
    zc = zero_columns(A_eq, A_ub, n_x) if presolve else np.zeros(0, dtype=np.int64)
    if zc.size:
        sgn = np.zeros(zc.size) if zero_col_sign is None else np.asarray(zero_col_sign, dtype=np.float64)
        neg, pos = zc[sgn < 0], zc[sgn > 0]
        x_pre[neg] = ub[neg]
        x_pre[pos] = lb[pos]
        if np.isinf(x_pre).any():
            raise PresolveComplete(3, x_pre)
        lb[neg] = ub[neg]
        ub[pos] = lb[pos]

    # -- singleton equality rows fix a variable (:1020-1041) ---------------------------------------
    A_eq_k = A_eq[eq_keep]
    sing_local = np.nonzero(np.diff(A_eq_k.indptr) == 1)[0] if presolve else np.zeros(0, dtype=np.int64)
    s_rows = eq_keep[sing_local]
    s_cols = A_eq_k.indices[A_eq_k.indptr[sing_local]] if sing_local.size else np.zeros(0, dtype=np.int64)
    s_coef = A_eq_k.data[A_eq_k.indptr[sing_local]] if sing_local.size else np.zeros(0)
    s_lo, s_hi = lb[s_cols].copy(), ub[s_cols].copy()
    fixed_by_row = {}  # user col -> (row, coef): value = b_eq[row] / coef (the last row wins, like the loop)
    for r, col, a in zip(s_rows, s_cols, s_coef):
        if b_eq_v is not None:
            val = b_eq_v[r] / a
            if not (lb[col] - TOL <= val <= ub[col] + TOL):
                raise PresolveComplete(2, x_pre)
            lb[col] = ub[col] = val
        fixed_by_row[int(col)] = (int(r), float(a))
    if sing_local.size:
        mask = np.ones(eq_keep.size, dtype=bool)
        mask[sing_local] = False
        eq_keep = eq_keep[mask]

    # -- singleton inequality rows tighten bounds (:1048-1072) -------------------------------------
    A_ub_k = A_ub[ub_keep]
    sing_ub = np.nonzero(np.diff(A_ub_k.indptr) == 1)[0] if presolve else np.zeros(0, dtype=np.int64)
    for loc in sing_ub:
        col = A_ub_k.indices[A_ub_k.indptr[loc]]
        a = A_ub_k.data[A_ub_k.indptr[loc]]
        val = b_ub_v[ub_keep[loc]] / a
        if a > 0:
            if val < lb[col] - TOL:
                raise PresolveComplete(2, x_pre)
            if val < ub[col]:
                ub[col] = val
        else:
            if val > ub[col] + TOL:
                raise PresolveComplete(2, x_pre)
            if val > lb[col]:
                lb[col] = val
    if sing_ub.size:
        mask = np.ones(ub_keep.size, dtype=bool)
        mask[sing_ub] = False
        ub_keep = ub_keep[mask]

    # -- fixed variables (:1075-1110) ---------------------------------------------------------------
    i_f = (np.abs(lb - ub) < TOL) if presolve else np.zeros(n_x, dtype=bool)
    for col in fixed_by_row:  # fixed whatever b_eq turns out to be
        i_f[col] = True
    A_eq2, A_ub2 = A_eq[eq_keep], A_ub[ub_keep]
    if i_f.all() and b_eq_v is not None:
        residual = b_eq_v[eq_keep] - A_eq2 @ lb
        slack = b_ub_v[ub_keep] - A_ub2 @ lb
        if (A_ub2.nnz > 0 and (slack < 0).any()) or (A_eq2.nnz > 0 and not np.allclose(residual, 0)):
            raise PresolveComplete(2, x_pre)
    fixed_idx = np.nonzero(i_f)[0]
    var_index = np.nonzero(~i_f)[0]
    nf = fixed_idx.size
    # fixed values = const0 + map @ b_eq
    fixed_const0 = np.where(np.isfinite(lb[fixed_idx]), lb[fixed_idx], 0.0)
    fr, fc, fv = [], [], []
    for k, col in enumerate(fixed_idx):
        if int(col) in fixed_by_row:
            r, a = fixed_by_row[int(col)]
            fixed_const0[k] = 0.0
            fr.append(k)
            fc.append(r)
            fv.append(1.0 / a)
    fixed_from_beq = sp.csr_matrix((fv, (fr, fc)), shape=(nf, n_eq))
    A_eq_f, A_ub_f = A_eq2[:, fixed_idx], A_ub2[:, fixed_idx]
    A_eq3, A_ub3 = A_eq2[:, var_index].tocsr(), A_ub2[:, var_index].tocsr()
    # right-hand sides after removing fixed variables (:1096-1097)
    beq_const = -(A_eq_f @ fixed_const0)
    beq_map = I_eq[eq_keep] - A_eq_f @ fixed_from_beq
    bub_const = b_ub_v[ub_keep] - A_ub_f @ fixed_const0
    bub_map = -(A_ub_f @ fixed_from_beq)
    lb_k, ub_k = lb[var_index], ub[var_index]
    nvar = var_index.size

    if presolve and A_eq3.nnz == 0 and A_ub3.nnz == 0:
        # no non-trivial constraints left (:1113-1144): the answer depends on c only through its sign; callers get the
        # c-independent part (fixed variables) in x and the bounds of the remaining variables for the sign rule
        x_done = x_pre.copy()
        if b_eq_v is not None or not fixed_by_row:
            x_done[fixed_idx] = fixed_const0 + (fixed_from_beq @ b_eq_v if (n_eq and b_eq_v is not None) else 0.0)
        raise PresolveComplete(0, x_done, trivial=(var_index, lb_k.copy(), ub_k.copy()),
                               fixed=(fixed_idx, fixed_const0, fixed_from_beq))

    # -- _get_Abc (:333-523) ------------------------------------------------------------------------
    lbs, ubs = lb_k.copy(), ub_k.copy()
    nolb = np.isinf(lbs) & ~np.isinf(ubs)            # x = -x'  (:453-468)
    col_sign = np.ones(nvar)
    col_sign[nolb] = -1.0
    x_flip_ub = np.full(nvar, np.nan)
    x_flip_ub[nolb] = ubs[nolb]
    lbs[nolb], ubs[nolb] = -ubs[nolb], np.inf
    S = sp.diags(col_sign)
    A_ub4, A_eq4 = (A_ub3 @ S).tocsr(), (A_eq3 @ S).tocsr()
    has_ub = ~np.isinf(ubs)
    i_newub = np.nonzero(has_ub)[0]                  # bound rows FIRST (:479-482)
    nb = i_newub.size
    A_bnd = sp.csr_matrix((np.ones(nb), (np.arange(nb), i_newub)), shape=(nb, nvar))
    A_ubs = sp.vstack([A_bnd, A_ub4], format="csr")
    m_ub = A_ubs.shape[0]
    free = np.isinf(lbs) & np.isinf(ubs)             # x = x+ - x-  (:490-498)
    i_free = np.nonzero(free)[0]
    A1 = sp.vstack([A_ubs, A_eq4], format="csr")
    A_free = -A1[:, i_free]
    A2 = sp.vstack([sp.identity(m_ub, format="csr"), sp.csr_matrix((A_eq4.shape[0], m_ub))], format="csr")
    A_std = sp.hstack([A1, A_free, A2], format="csr")  # [vars | negative parts | slacks] (:505-506)
    A_std.eliminate_zeros()
    A_std.sort_indices()
    m, n = A_std.shape
    c_src = np.concatenate([var_index, var_index[i_free], -np.ones(m_ub, dtype=np.int64)]).astype(np.int64)
    c_sign = np.concatenate([col_sign, -col_sign[i_free], np.zeros(m_ub)])
    free_cols = -np.ones(nvar, dtype=np.int64)
    free_cols[i_free] = nvar + np.arange(i_free.size)
    # lower-bound shift (:510-519)
    shift = np.where(np.isinf(lbs), 0.0, lbs)
    # Post-solve constants (:1263-1274).  Upstream undoes the substitutions with the CALLER's bounds, not with the bounds
    # presolve tightened through singleton inequality rows (:1048-1072): x = x' + lb_user (or ub_user - x').  The same is
    # done here whenever the user's bound is of the same kind as the one _get_Abc substituted with; if presolve changed
    # the kind (a free variable that a singleton row bounded), upstream's loop reads a wrong column, and the
    # mathematically right constant is used instead.
    lbr, ubr = lb_raw[var_index], ub_raw[var_index]
    x_shift = np.where(nolb, 0.0, np.where(np.isfinite(lbr), lbr, shift))  # flipped variables: ub - x' has it all
    same_kind = nolb & np.isinf(lbr) & np.isfinite(ubr)
    x_flip_ub[same_kind] = ubr[same_kind]
    # Upstream's loop runs over the caller's bounds LIST (:1263 `enumerate(bounds)` with the raw argument, :1376): a
    # one-element list - broadcast to every variable by _clean_inputs for the solve - undoes the substitution of x[0]
    # only; every other variable is returned as the standard-form value x' (SURVEY C.8).  With the (0, None) / (0, 1)
    # lists the drivers pass nothing changes (lb = 0); with any other one-element list the result is upstream's, not
    # the LP's.
    n_raw = raw_bounds_count(bounds, n_x)
    beyond = var_index >= n_raw
    x_shift = np.where(beyond, 0.0, x_shift)
    x_flip_ub = np.where(beyond, np.nan, x_flip_ub)
    free_cols = np.where(beyond, -1, free_cols)
    b_const = np.concatenate([ubs[has_ub], bub_const, beq_const]) - A_std[:, :nvar] @ shift
    b_map = sp.vstack([sp.csr_matrix((nb, n_eq)), bub_map, beq_map], format="csr")
    sf = StandardForm(
        n_x=n_x, A=A_std, m=m, n=n, nvar=nvar, var_index=var_index, c_src=c_src, c_sign=c_sign,
        fixed_idx=fixed_idx, fixed_const=None, fixed_from_beq=fixed_from_beq, fixed_const0=fixed_const0,
        b_const=b_const, b_from_beq=b_map, b=None, zero_rows_eq=zero_rows_eq,
        singleton=(s_rows, s_cols, s_coef, s_lo, s_hi), x_shift=x_shift, x_flip_ub=x_flip_ub,
        free_cols=free_cols, n_eq_user=n_eq,
        meta=dict(m_ub=m_ub, n_free=int(i_free.size), n_bound_rows=int(nb), eq_keep=eq_keep, ub_keep=ub_keep,
                  # b was moved by A lb / A ub, or variables were fixed: presolve acted on the NUMBERS of A
                  number_dependent=bool(fixed_idx.size or np.any(shift != 0) or np.any(nolb))))
    if b_eq_v is not None:
        sf.b = b_const + (b_map @ b_eq_v if n_eq else 0.0)
        sf.fixed_const = fixed_const0 + (fixed_from_beq @ b_eq_v if n_eq else 0.0)
    return sf
