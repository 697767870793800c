"""TEST INFRASTRUCTURE — the CPU oracle.  Not part of the product path.

An fp64 NumPy/SciPy restatement of the reference's differentiable interior-point
LP layer (JayMan91/NeurIPSIntopt, `Interior/ip_model_whole.py` and
`Interior/ip_model_alter.py`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module,
and only as the checker / the timed CPU baseline — the product
(`neuripsintopt_b200/`) never does.

Parity pin: the reference ships no tests or golden vectors ("parity unpinned"
upstream).  This restatement is pinned instead against outputs of
the reference itself run in the build container through `oracle/ref_shim.py`
(fixtures under `tests/golden/`, generator `oracle/gen_golden.py`), and the
converged objective against HiGHS (`scipy.optimize.linprog`).

Every function cites the reference lines it follows.  Arithmetic is fp64
throughout; the reference's backward runs in float32 (ip_model_whole.py
:1445-1466,1541-1544) — `backward_*_asref` reproduce that as-executed float32
path, `backward_*` are the same formulas in fp64 (the primary parity target).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

# --------------------------------------------------------------------------------------
# Host preparation: _clean_inputs -> _presolve -> _get_Abc   (ip_model_whole.py:67-329,
# :918-1161, :333-523) restated for dense ndarray inputs.
# --------------------------------------------------------------------------------------


def _broadcast_bounds(bounds, n_x):
    """ip_model_whole.py:256-317: a one-tuple list applies to every variable; None/±inf = no bound."""
    if bounds is None or len(bounds) == 0:
        bounds = [(0, None)] * n_x
    elif len(bounds) == 1:
        bounds = [tuple(bounds[0])] * n_x
    elif len(bounds) == n_x:
        try:
            len(bounds[0])
        except TypeError:
            bounds = [(bounds[0], bounds[1])] * n_x
    elif len(bounds) == 2 and np.isreal(bounds[0]) and np.isreal(bounds[1]):
        bounds = [(bounds[0], bounds[1])] * n_x
    else:
        raise ValueError("bounds: need one (lb, ub) pair per variable")
    lb = np.empty(n_x)
    ub = np.empty(n_x)
    for i, (l, u) in enumerate(bounds):
        if l is not None and u is not None and l > u:
            raise ValueError("bound %d: lb > ub" % i)
        if l == np.inf or u == -np.inf:
            raise ValueError("bound %d: invalid infinite bound" % i)
        lb[i] = -np.inf if (l is None or l == -np.inf) else float(l)
        ub[i] = np.inf if (u is None or u == np.inf) else float(u)
    return lb, ub


def standard_form(c, A_eq=None, b_eq=None, A_ub=None, b_ub=None, bounds=None, tol=1e-9, presolve=True):
    """Restates forward()'s host prep, ip_model_whole.py:1317-1383.

    Returns a dict with the standard-form `A (m×n), b (m), c (n)` of
    `min c·x  s.t. A x = b, x >= 0`, `var_index` (kept original variables, ascending),
    `fixed_idx/fixed_val` (presolve `undo`), `n_x`, `lb_shift` (lower bounds of the kept
    variables; `_postprocess` :1263-1274 adds bounds[0]'s lb to x[0] only — see SURVEY C.8),
    or `complete=True` with the presolve `x` when presolve decides the problem (:1483-1488).
    """
    c = np.array(c, dtype=np.float64).reshape(-1)
    n_x = c.size
    if not np.isfinite(c).all():
        raise ValueError("c must be finite")  # :145-148 -> LinAlgError at :1341
    A_eq = np.zeros((0, n_x)) if A_eq is None or np.size(A_eq) == 0 else np.array(A_eq, dtype=np.float64)
    A_ub = np.zeros((0, n_x)) if A_ub is None or np.size(A_ub) == 0 else np.array(A_ub, dtype=np.float64)
    b_eq = np.zeros(0) if b_eq is None or np.size(b_eq) == 0 else np.array(b_eq, dtype=np.float64).reshape(-1)
    b_ub = np.zeros(0) if b_ub is None or np.size(b_ub) == 0 else np.array(b_ub, dtype=np.float64).reshape(-1)
    lb, ub = _broadcast_bounds(bounds, n_x)
    # the caller's list as given (one tuple, or one per variable): _postprocess adjusts with it (:1263-1274)
    raw_bounds = [tuple(None if (v is None or (isinstance(v, float) and np.isinf(v))) else float(v) for v in bd)
                  for bd in (bounds or [])]
    out = dict(n_x=n_x, complete=False, status=0, raw_bounds=raw_bounds)

    # ---- _presolve (:918-1161), single pass, rule order preserved -----------------------
    x_pre = np.zeros(n_x)
    i_f = np.zeros(n_x, dtype=bool)  # presolve=False (:1352): nothing removed, nothing fixed
    i_nf = ~i_f
    if presolve:
        zr = (A_eq != 0).sum(axis=1) == 0  # :957
        if zr.any():
            if (zr & (np.abs(b_eq) > tol)).any():
                return dict(out, complete=True, status=2, x=x_pre)
            A_eq, b_eq = A_eq[~zr], b_eq[~zr]
        zr = (A_ub != 0).sum(axis=1) == 0  # :977
        if zr.any():
            if (zr & (b_ub < -tol)).any():
                return dict(out, complete=True, status=2, x=x_pre)
            A_ub, b_ub = A_ub[~zr], b_ub[~zr]
        Aall = np.vstack((A_eq, A_ub))  # :995
        if Aall.shape[0] > 0:
            zc = (Aall != 0).sum(axis=0) == 0
            neg, pos = zc & (c < 0), zc & (c > 0)
            x_pre[neg] = ub[neg]
            x_pre[pos] = lb[pos]
            if np.isinf(x_pre).any():
                return dict(out, complete=True, status=3, x=x_pre)
            lb[neg] = ub[neg]
            ub[pos] = lb[pos]
        sr = (A_eq != 0).sum(axis=1) == 1  # :1020
        if sr.any():
            for row in np.where(sr)[0]:
                col = np.nonzero(A_eq[row])[0][0]
                val = b_eq[row] / A_eq[row, col]
                if not (lb[col] - tol <= val <= ub[col] + tol):
                    return dict(out, complete=True, status=2, x=x_pre)
                lb[col] = val
                ub[col] = val
            A_eq, b_eq = A_eq[~sr], b_eq[~sr]
        sr = (A_ub != 0).sum(axis=1) == 1  # :1048
        if sr.any():
            for row in np.where(sr)[0]:
                col = np.nonzero(A_ub[row])[0][0]
                val = b_ub[row] / A_ub[row, col]
                if A_ub[row, col] > 0:
                    if val < lb[col] - tol:
                        return dict(out, complete=True, status=2, x=x_pre)
                    if val < ub[col]:
                        ub[col] = val
                else:
                    if val > ub[col] + tol:
                        return dict(out, complete=True, status=2, x=x_pre)
                    if val > lb[col]:
                        lb[col] = val
            A_ub, b_ub = A_ub[~sr], b_ub[~sr]
        i_f = np.abs(lb - ub) < tol  # :1075
        i_nf = ~i_f
        if i_f.all():  # :1079
            residual = b_eq - A_eq.dot(lb)
            slack = b_ub - A_ub.dot(lb)
            if (A_ub.size > 0 and (slack < 0).any()) or (A_eq.size > 0 and not np.allclose(residual, 0)):
                return dict(out, complete=True, status=2, x=x_pre)
    fixed_idx = np.zeros(0, dtype=np.int64)
    fixed_val = np.zeros(0)
    c0 = 0.0
    lb_mod, ub_mod = lb, ub
    if i_f.any():  # :1094-1110
        c0 += c[i_f].dot(lb[i_f])
        b_eq = b_eq - A_eq[:, i_f].dot(lb[i_f])
        b_ub = b_ub - A_ub[:, i_f].dot(lb[i_f])
        c = c[i_nf]
        x_pre = x_pre[i_nf]
        A_eq = A_eq[:, i_nf]
        A_ub = A_ub[:, i_nf]
        fixed_idx = np.nonzero(i_f)[0]
        fixed_val = lb[i_f].copy()
        lb_mod, ub_mod = lb[i_nf], ub[i_nf]
    out.update(fixed_idx=fixed_idx, fixed_val=fixed_val,
               var_index=np.setdiff1d(np.arange(n_x), fixed_idx))
    if presolve and A_eq.size == 0 and A_ub.size == 0:  # :1113-1144 trivial problem
        x = x_pre.copy()
        status = 0
        if c.size and (((c < 0) & (ub_mod == np.inf)).any() or ((c > 0) & (lb_mod == -np.inf)).any()):
            status = 3
        x[c < 0] = ub_mod[c < 0]
        x[c > 0] = lb_mod[c > 0]
        xz = ub_mod[c == 0].copy()
        xz[np.isinf(xz)] = 0
        x[c == 0] = xz
        full = np.zeros(n_x)
        full[out["var_index"]] = x
        full[fixed_idx] = fixed_val
        return dict(out, complete=True, status=status, x=full)

    # ---- _get_Abc (:333-523) ---------------------------------------------------------------
    lbs, ubs = lb_mod.copy(), ub_mod.copy()
    n = c.size
    c = c.copy()
    A_ub = A_ub.copy()
    A_eq = A_eq.copy()
    lb_none, ub_none = np.isinf(lbs), np.isinf(ubs)
    nolb = lb_none & ~ub_none  # :453-468  x -> -x'
    lbs[nolb], ubs[nolb] = -ubs[nolb], np.inf
    c[nolb] *= -1
    A_ub[:, nolb] *= -1
    A_eq[:, nolb] *= -1
    lb_none, ub_none = np.isinf(lbs), np.isinf(ubs)
    i_newub = np.nonzero(~ub_none)[0]  # :471-482  bound rows go FIRST
    nb = i_newub.size
    Abnd = np.zeros((nb, n))
    Abnd[np.arange(nb), i_newub] = 1
    A_ub = np.vstack((Abnd, A_ub))
    b_ub = np.concatenate((ubs[~ub_none], b_ub))
    m_ub = A_ub.shape[0]
    A1 = np.vstack((A_ub, A_eq))  # :484
    b = np.concatenate((b_ub, b_eq))
    i_free = np.nonzero(lb_none & ub_none)[0]  # :490-498  x = x+ - x-
    A1 = np.hstack((A1, -A1[:, i_free]))
    c_std = np.concatenate((c, -c[i_free], np.zeros(m_ub)))
    A2 = np.vstack((np.eye(m_ub), np.zeros((A_eq.shape[0], m_ub))))  # :505 slacks last
    A = np.hstack((A1, A2))
    i_shift = np.nonzero(~lb_none)[0]  # :510-519  x = x' + lb
    lb_shift = lbs[~lb_none]
    c0 += float(np.sum(lb_shift * c_std[i_shift]))
    b = b - (A[:, i_shift] * lb_shift).sum(axis=1)
    out.update(A=A, b=b, c=c_std, c0=c0, nvar=n, m_ub=m_ub, n_free=i_free.size, n_bound_rows=nb,
               lb0=(lb_mod[0] if n and np.isfinite(lb_mod[0]) else 0.0))
    return out


def postprocess(x_hat, sf):
    """_postprocess (:1241-1285): re-insert fixed variables in ascending index order, undo the variable substitutions,
    truncate to n_x.  The adjustment loop (:1261-1274) runs over the CALLER's raw bounds list: with the one-element list
    every driver passes, only x[0] is adjusted (lb = 0 there), with one entry per variable every
    non-fixed variable is — by the caller's bounds, not by the ones presolve tightened."""
    x = list(np.asarray(x_hat, dtype=np.float64))
    for i, v in zip(sf["fixed_idx"], sf["fixed_val"]):
        x.insert(int(i), float(v))
    x = np.array(x)
    n_x = sf["n_x"]
    no_adjust = set(int(i) for i in sf["fixed_idx"])
    n_unbounded = 0
    for i, bnds in enumerate(sf.get("raw_bounds") or []):
        if i in no_adjust:
            continue
        lb, ub = bnds
        if lb is None and ub is None:
            n_unbounded += 1
            x[i] = x[i] - x[n_x + n_unbounded - 1]
        elif lb is None:
            x[i] = ub - x[i]
        else:
            x[i] += lb
    return x[:n_x]


# --------------------------------------------------------------------------------------
# Forward: homogeneous self-dual predictor-corrector (ip_model_whole.py:647-917,1385-1447)
# --------------------------------------------------------------------------------------


def _step(x, dx, t, dt, tau, dtau, kappa, dkappa, alpha0):
    """_get_step, :870-878."""
    ix, it = dx < 0, dt < 0
    ax = alpha0 * np.min(x[ix] / -dx[ix]) if ix.any() else 1.0
    at = alpha0 * np.min(t[it] / -dt[it]) if it.any() else 1.0
    atau = alpha0 * tau / -dtau if dtau < 0 else 1.0
    akap = alpha0 * kappa / -dkappa if dkappa < 0 else 1.0
    return min(1.0, ax, atau, at, akap)


STATUS_OK, STATUS_MAXITER, STATUS_INFEAS, STATUS_BREAKDOWN, STATUS_NONFINITE = 0, 1, 2, 3, 4


def forward(A, b, c, alpha0=0.9995, tol=1e-8, thr=0.0, max_iter=1000, damping_fwd=1e-6,
            init=None, trace=None, pc=True, beta=0.1):
    """One LP: returns dict(x,y,t,tau,kappa,iters,status,mu).  `A` dense (m×n) fp64.

    Follows the loop at ip_model_whole.py:1385-1423 with `_get_delta` (:761-859, pc=True),
    `_get_step` (:870), `_do_step` (:861), `_indicators` (:880-917).  The `(p,q)` solve is
    done once per iteration (the reference repeats it identically in both passes, :800).
    damping_fwd = 1e-6 for `whole` (:774), = `damping` kwarg for `alter` (ip_model_alter.py:775).
    """
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    m, n = A.shape
    if init is None:  # :647-654
        x, y, t, tau, kappa = np.ones(n), np.zeros(m), np.ones(n), 1.0, 1.0
    else:
        x, y, t = (np.array(init[k], dtype=np.float64) for k in ("x", "y", "t"))
        tau, kappa = float(np.squeeze(init["tau"])), float(np.squeeze(init["kappa"]))
    # cold-start residual norms for the indicators (:884,907-910) — constants per (A,b,c)
    rp0 = max(1.0, np.linalg.norm(b - A.dot(np.ones(n))))
    rd0 = max(1.0, np.linalg.norm(c - np.ones(n)))
    rg0 = max(1.0, abs(1.0 + c.sum()))
    iters, status, mu = 0, STATUS_OK, 1.0
    go = True
    while go:
        iters += 1
        r1 = -(A.dot(x) - b * tau)  # :765
        r2 = -(A.T.dot(y) + t - c * tau)  # :766
        r3 = -(-c.dot(x) + b.dot(y) - kappa)  # :767
        mu = (x.dot(t) + tau * kappa) / (n + 1)  # :770
        D = x / t  # :772
        M = A.dot(D.reshape(-1, 1) * A.T)  # :773
        M[np.diag_indices(m)] += damping_fwd  # :774-775
        solved = True
        # _get_solver (:683-759): Cholesky first; on failure the reference's retry ladder (:798-822)
        # re-solves with scipy.linalg.solve(assume_a = 'sym', then 'pos', then 'gen').
        solver = None
        try:
            Lc = sla.cho_factor(M)  # :742
            solver = lambda r: sla.cho_solve(Lc, r)  # noqa: E731
        except (sla.LinAlgError, ValueError):
            for assume in ("sym", "pos", "gen"):  # :819-822
                try:
                    sla.solve(M, b, assume_a=assume)
                    solver = (lambda a_: lambda r: sla.solve(M, r, assume_a=a_))(assume)
                    break
                except (sla.LinAlgError, ValueError):
                    continue
            if solver is None:
                solved = False
        if solved:
            def sym_solve(ra, rb):  # :663-681
                v = solver(rb + A.dot(D * ra))
                return D * (A.T.dot(v) - ra), v

            with np.errstate(all="ignore"):
                p, q = sym_solve(c, b)  # :800
            if not (np.isfinite(p).all() and np.isfinite(q).all()):
                solved = False
        if solved:
            gamma = 0.0 if pc else beta * np.mean(t * x)  # :1392
            dx = dt = None
            dtau = dkappa = 0.0
            for i in ((0, 1) if pc else (0,)):  # :783, :846-847
                eta = 1.0 - gamma
                rxt = -(x * t - gamma * mu)  # :788
                rtk = -(tau * kappa - gamma * mu)  # :789
                if i == 1:  # Mehrotra :792-795
                    rxt = rxt - dx * dt
                    rtk = rtk - dtau * dkappa
                u, v = sym_solve(eta * r2 - (1 / x) * rxt, eta * r1)  # :801 (operation order kept)
                dtau = ((eta * r3 + 1 / tau * rtk - (-c.dot(u) + b.dot(v))) /
                        (1 / tau * kappa + (-c.dot(p) + b.dot(q))))  # :837-838
                dx = u + p * dtau  # :839-842
                dy = v + q * dtau
                dt = (1 / x) * (rxt - t * dx)
                dkappa = (rtk - kappa * dtau) / tau
                alpha = _step(x, dx, t, dt, tau, dtau, kappa, dkappa, 1.0)  # :843
                gamma = (1 - alpha) ** 2 * min(0.1, 1 - alpha)  # :844
            alpha = _step(x, dx, t, dt, tau, dtau, kappa, dkappa, alpha0)  # :1402
            x = x + alpha * dx  # :861-868
            tau = tau + alpha * dtau
            t = t + alpha * dt
            kappa = kappa + alpha * dkappa
            y = y + alpha * dy
        if trace is not None:
            trace.append(dict(x=x.copy(), y=y.copy(), t=t.copy(), tau=tau, kappa=kappa))
        # _indicators :880-917
        bty, ctx = b.dot(y), c.dot(x)
        rho_A = abs(ctx - bty) / (tau + abs(bty))
        rho_p = np.linalg.norm(b * tau - A.dot(x)) / rp0
        rho_d = np.linalg.norm(c * tau - A.T.dot(y) - t) / rd0
        rho_g = abs(kappa + ctx - bty) / rg0
        mu = (x.dot(t) + tau * kappa) / (n + 1)
        rho_mu = mu  # mu_0 = 1
        go = (rho_p > tol or rho_d > tol or rho_A > tol) and (mu > thr)  # :1408
        inf1 = rho_p < tol and rho_d < tol and rho_g < tol and tau < tol * max(1, kappa)
        inf2 = rho_mu < tol and tau < tol * min(1, kappa)
        if inf1 or inf2:  # :1412
            status = STATUS_INFEAS
            break
        if max_iter < iters:  # :1414
            status = STATUS_MAXITER
            break
        if not solved:  # :1419
            status = STATUS_BREAKDOWN
            break
    return dict(x=x, y=y, t=t, tau=tau, kappa=kappa, iters=iters, status=status, mu=mu)


# --------------------------------------------------------------------------------------
# Backward
# --------------------------------------------------------------------------------------


def backward_whole(A, b, c, st, g, damping, nvar=None):
    """fp64 VJP form of ip_model_whole.py:1515-1560 (HSD embedding), SURVEY Appendix B.1.
    Uses the un-normalised x,t,tau,kappa (:1445-1447).  `g` has length nvar (kept user
    variables); returns dc of length nvar."""
    x, t, tau, kappa = st["x"], st["t"], st["tau"], st["kappa"]
    m, n = A.shape
    nvar = n if nvar is None else nvar
    gh = np.zeros(n)
    gh[:nvar] = g
    mu = (x.dot(t) + tau * kappa) / (n + 1)  # :1517
    x2 = x * x
    M = (A * x2).dot(A.T)  # :1525-1527
    M[np.diag_indices(m)] += damping  # :1535
    L = sla.cho_factor(M)
    q = sla.cho_solve(L, mu * b + A.dot(x2 * c))  # :1530,1542
    p = x2 * (A.T.dot(q) - c) / mu  # :1548
    w = sla.cho_solve(L, tau * A.dot(x2 * gh))  # v @ g, :1529,1541
    ug = x2 * (A.T.dot(w) - tau * gh) / mu  # u @ g, :1549
    s = (x.dot(gh) + c.dot(ug) - b.dot(w)) / (b.dot(q) - c.dot(p) + mu / tau ** 2)  # :1552-1553
    return (ug + p * s)[:nvar]  # :1554-1556


def backward_alter(A, b, c, st, g, damping, nvar=None):
    """fp64 VJP form of ip_model_alter.py:1524-1620 (log-barrier).  Uses x̂=x/τ, t̂=t/τ
    (ip_model_alter.py:1459-1461); mu = x̂·t̂/n (:1525)."""
    xh, th = st["x"] / st["tau"], st["t"] / st["tau"]
    m, n = A.shape
    nvar = n if nvar is None else nvar
    gh = np.zeros(n)
    gh[:nvar] = g
    mu = xh.dot(th) / n
    x2 = xh * xh
    M = (A * x2).dot(A.T)
    M[np.diag_indices(m)] += damping
    w = sla.cho_solve(sla.cho_factor(M), A.dot(x2 * gh))
    return (-x2 * (gh - A.T.dot(w)) / mu)[:nvar]  # u = -X2 (I - Aᵀv); dc = u[:nvar,:nvar] g / mu


def backward_whole_asref(A, b, c, st, g, damping, nvar=None, f=np.float32):
    """As-executed restatement of ip_model_whole.py:1506-1556: float32 tensors, n right-hand
    sides, the dense n×n Jacobian and the logging-only matrix_rank SVD (:1536).  This is what
    the reference CPU path costs; used for the secondary parity report and the CPU baseline.
    With f=np.float64 it is the same dense-Jacobian formula in fp64 (checks the VJP algebra)."""
    x, t = st["x"].astype(f), st["t"].astype(f)
    tau, kappa = f(st["tau"]), f(st["kappa"])
    A, b, c = A.astype(f), b.astype(f), c.astype(f)
    m, n = A.shape
    nvar = n if nvar is None else nvar
    g = np.asarray(g, dtype=f)
    mu = (x.dot(t) + tau * kappa) / f(n + 1)
    x2 = x ** 2
    a = A * x
    M = a.dot(a.T)
    r1 = tau * A * x2
    r2 = mu * b + A.dot(x2 * c)
    M[np.diag_indices(m)] += f(damping)
    np.linalg.matrix_rank(M, tol=1e-3)
    v = sla.solve(M, r1, assume_a="sym").astype(f)
    q = sla.solve(M, r2, assume_a="sym").astype(f)
    p = x2 * (A.T.dot(q) - c) / mu
    u = x2.reshape(-1, 1) * (A.T.dot(v) - tau * np.eye(n, dtype=f)) / mu
    dtau = (x + u.T.dot(c) - v.T.dot(b)) / (b.dot(q) - c.dot(p) + mu / tau ** 2)
    J = u + np.outer(p, dtau)
    return J[:nvar, :nvar].dot(g)


def backward_alter_asref(A, b, c, st, g, damping, nvar=None, f=np.float32):
    """As-executed float32 restatement of ip_model_alter.py:1524-1620 (f=np.float64: same in fp64)."""
    x, t = (st["x"] / st["tau"]).astype(f), (st["t"] / st["tau"]).astype(f)
    A = A.astype(f)
    m, n = A.shape
    nvar = n if nvar is None else nvar
    g = np.asarray(g, dtype=f)
    mu = x.dot(t) / f(n)
    x2 = x ** 2
    a = A * x
    M = a.dot(a.T)
    M[np.diag_indices(m)] += f(damping)
    v = sla.solve(M, A * x2, assume_a="sym").astype(f)
    u = -x2.reshape(-1, 1) * (np.eye(n, dtype=f) - A.T.dot(v))
    return u[:nvar, :nvar].dot(g) / mu


# --------------------------------------------------------------------------------------
# Whole layer on user-form inputs, like IPOfunc(A,b,G,h,...)(c) + backward
# --------------------------------------------------------------------------------------


def ipofunc(c, g=None, A=None, b=None, G=None, h=None, bounds=None, variant="whole",
            alpha0=0.9995, tol=1e-8, thr=None, max_iter=1000, damping=1e-2, asref=False,
            init=None, method=1, presolve=True, pc=True, beta=0.1):
    """fp64 oracle of `IPOfunc(A,b,G,h,...)(c)` followed by `(x*g).sum().backward()`.

    Inputs are rounded to float32 first, like the torch tensors the drivers pass
    (ip_model_whole.py:1299-1316).  Returns dict(x (n_x), grad (n_x), iters, status, state, sf)."""
    f32 = lambda v: None if v is None or np.size(v) == 0 else np.asarray(v, dtype=np.float32).astype(np.float64)
    c = np.asarray(c, dtype=np.float32).astype(np.float64)
    if bounds is None:  # whole :1318, alter ip_model_alter.py:1331
        bounds = [(0.0, None)] if variant == "whole" else [(0.0, 1.0)]
    thr = 0.0 if thr is None else thr
    max_iter = 1000 if max_iter is None else max_iter
    sf = standard_form(c, f32(A), f32(b), f32(G), f32(h), bounds, presolve=presolve)
    n_x = sf["n_x"]
    if sf["complete"]:
        return dict(x=sf["x"], grad=np.zeros(n_x), iters=0, status=sf["status"], state=None, sf=sf)
    dfwd = 1e-6 if variant == "whole" else damping
    st = forward(sf["A"], sf["b"], sf["c"], alpha0, tol, thr, max_iter, dfwd, init, pc=pc, beta=beta)
    x = postprocess(st["x"] / st["tau"], sf)
    grad = None
    if g is not None:
        gk = np.asarray(g, dtype=np.float64)[sf["var_index"]]
        if variant == "whole":
            fn = backward_whole_asref if asref else backward_whole
        else:
            fn = backward_alter_asref if asref else backward_alter
        A_s, b_s, c_s, st_s = sf["A"], sf["b"], sf["c"], st
        if method in (2, 3):
            # saved-tensor slices (ip_model_whole.py:1467-1482): the backward sees only the last n_ub + n_eq rows
            # (the rows that state upper bounds are dropped) and, for method 2, only the first n_x_afterpre
            # columns (slack and split columns dropped)
            r0 = sf["n_bound_rows"]
            nc = sf["nvar"] if method == 2 else sf["A"].shape[1]
            A_s, b_s, c_s = sf["A"][r0:, :nc], sf["b"][r0:], sf["c"][:nc]
            st_s = dict(st, x=st["x"][:nc], t=st["t"][:nc], y=st["y"][r0:])
        dc = fn(A_s, b_s, c_s, st_s, gk, damping, sf["nvar"])
        grad = np.zeros(n_x)
        grad[sf["var_index"]] = dc
    return dict(x=x, grad=grad, iters=st["iters"], status=st["status"], state=st, sf=sf)
