"""`IPOfunc` — the reference's differentiable LP layer, batched, on B200.

Host-side mirror of `IPOfunc` (Interior/ip_model_whole.py:1287-1592 and Interior/ip_model_alter.py
:1287-1655): same factory signature, same positional order `(A, b, G, h)` = (equality, rhs, inequality,
rhs), same defaults, returns `torch.autograd.Function.apply` mapping a cost vector `c` to the solution
`x` with a hand-written backward returning dL/dc.  Additive extensions: `c` may be `(B, n_x)` (a minibatch
of cost vectors, solved concurrently, one CTA each) and `b` may be `(B, n_eq)` (per-instance right-hand
sides, shortespath/shortespath.py:407-413).

All arithmetic of the solver loop and of the backward runs in fp64 inside libintopt_b200.so (hand-written
sm_100a CUDA, reached through the C ABI of include/intopt_b200.h).  PyTorch is only used for device
memory, the current stream, and the index gathers around the call.  No CPU fallback exists: a CPU tensor
or a missing library raises.

The drivers re-create `IPOfunc(...)` for every instance (Interior/intopt_energy_mlp.py:422,
shortespath/shortespath.py:421); the expensive part — presolve, standard form, structure analysis, the
device-resident problem — is therefore cached on the *content* of (A, G, h, bounds), and the right-hand
side `b` only enters through a cached linear map.
"""
from __future__ import annotations

import ctypes
import hashlib
import weakref
from collections import OrderedDict

import numpy as np
import scipy.sparse as sp
import torch
from scipy.linalg import LinAlgError

from . import _capi
from . import problem as _problem

STATUS_OK, STATUS_MAXITER, STATUS_INFEAS, STATUS_BREAKDOWN, STATUS_NONFINITE = 0, 1, 2, 3, 4

_PROBLEM_CACHE: "OrderedDict[tuple, object]" = OrderedDict()
_PROBLEM_CACHE_MAX = 64  # e.g. the knapsack experiment cycles through 10 minibatch cost rows (one problem each)


def clear_cache():
    _PROBLEM_CACHE.clear()


def _to_numpy(v):
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    v = np.asarray(v)
    if v.ndim == 0 or v.shape[0] == 0:  # empty torch.Tensor() means "no such constraints" (:1300-1316)
        return None
    return np.ascontiguousarray(v, dtype=np.float64)


def _digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"<none>")
        else:
            h.update(str(a.shape).encode())
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _gather_convert(src, src_stride, B, n_out, idx, scale, constv, dst, stream):
    """dst[b, j] = idx[j] >= 0 ? scale[j] * src[b, idx[j]] : constv[j] in one launch (float <-> double on the way)."""
    _capi.check(_capi.lib().intopt_gather_convert(
        _ptr(src), 1 if src.dtype == torch.float32 else 0, src_stride, B, n_out, _ptr(idx), _ptr(scale), _ptr(constv),
        _ptr(dst), 1 if dst.dtype == torch.float32 else 0, dst.stride(0), ctypes.c_void_p(stream)), "intopt_gather_convert")


def _glue_ok(t):
    return t.dtype in (torch.float32, torch.float64) and t.dim() == 2 and t.stride(1) == 1


class _Token:
    """Weak-referenceable identity of one IPOfunc closure."""
    __slots__ = ("__weakref__",)


_TAG0 = 1024.0  # tags of the pattern matrix that stands in for a batch of equality matrices (per-instance A)


class _DeviceHandle:
    """An IntoptProblem handle on one device and its workspace."""

    handle = None

    def _create(self, A, device, inst_values):
        """A: scipy CSR (m x n).  Returns its value array in the order the handle (and per-instance values) use."""
        self.device = device
        self.inst_values = inst_values
        L = _capi.lib()
        rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        val = np.ascontiguousarray(A.data, dtype=np.float64)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _capi.check(L.intopt_problem_create(ctypes.byref(self.handle), A.shape[0], A.shape[1], rp.ctypes.data,
                                                ci.ctypes.data, val.ctypes.data, None,
                                                _capi.FLAG_INSTANCE_VALUES if inst_values else 0),
                        "intopt_problem_create")
        info = _capi.IntoptInfo()
        _capi.check(L.intopt_problem_info(self.handle, ctypes.byref(info)), "intopt_problem_info")
        self.info = info.as_dict()
        self.nnz = int(val.size)
        self._ws = {}
        return val

    def __del__(self):
        try:
            if self.handle:
                _capi.lib().intopt_problem_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def workspace(self, B):
        ws = self._ws.get(B)
        if ws is None:
            nbytes = _capi.lib().intopt_workspace_bytes(self.handle, B)
            ws = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=self.device)
            self._ws = {B: ws}  # keep one size only
        return ws


class SlicedProblem(_DeviceHandle):
    """The sub-problem the reference's backward sees for `method` 2 / 3 (saved-tensor slices,
    ip_model_whole.py:1467-1482): the rows of the standard form that state upper bounds are dropped (they come first,
    :471-482), and for method 2 so are all columns beyond the presolved variables (slack and split columns)."""

    def __init__(self, dp, method):
        sf = dp.sf
        self.r0 = int(sf.meta["n_bound_rows"])
        self.nc = int(sf.nvar if method == 2 else sf.n)
        A = sp.csr_matrix(sf.A)
        tagged = sp.csr_matrix((np.arange(1, A.nnz + 1, dtype=np.float64), A.indices, A.indptr), shape=A.shape)
        sub = tagged[self.r0:, :self.nc].tocsr()
        sub.sort_indices()
        self.entry = (sub.data - 1).astype(np.int64)   # entry e of the slice = entry self.entry[e] of the full matrix
        sub.data = np.asarray(A.data, dtype=np.float64)[self.entry]
        self.m, self.n = sub.shape
        self.trivial = self.r0 == 0 and self.nc == sf.n
        if not self.trivial:
            self._create(sub, dp.device, dp.inst_values)
            self.entry_dev = torch.as_tensor(self.entry, dtype=torch.long, device=dp.device)

    def slices(self, sf, c_std, b_std, st, a_vals):
        """(c, b, state, values) of the slice, contiguous."""
        n, m, r0, nc = sf.n, sf.m, self.r0, self.nc
        st_s = torch.cat([st[:, :nc], st[:, n + r0:n + m], st[:, n + m:n + m + nc], st[:, 2 * n + m:]], dim=1)
        return (c_std[:, :nc].contiguous(), b_std[:, r0:].contiguous(), st_s,
                None if a_vals is None else a_vals[:, self.entry_dev].contiguous())


class DeviceProblem(_DeviceHandle):
    """A StandardForm plus its device-resident IntoptProblem handle and index tensors (one per device)."""

    def __init__(self, sf, device, inst_values=False):
        self.sf = sf
        val = self._create(sf.A, device, inst_values)
        self._sliced = {}
        self.per_closure = weakref.WeakKeyDictionary()  # IPOfunc closure token -> its right-hand-side tensors
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
        kept = sf.c_src >= 0
        cols = np.nonzero(kept)[0]
        self.c_cols = t(cols, torch.long)                 # standard-form columns fed by user costs
        self.c_src = t(sf.c_src[kept], torch.long)
        self.c_sign = t(sf.c_sign[kept], torch.float64)
        self.identity_vars = bool(sf.nvar == sf.n_x and sf.fixed_idx.size == 0)
        self.c_plain = bool((sf.c_sign[kept] == 1).all() and cols.size == sf.nvar)
        self.var_index = t(sf.var_index, torch.long)
        self.fixed_idx = t(sf.fixed_idx, torch.long)
        self.x_shift = t(sf.x_shift, torch.float64) if np.any(sf.x_shift != 0) else None
        self.x_flip = t(sf.x_flip_ub, torch.float64) if np.isfinite(sf.x_flip_ub).any() else None
        self.free_cols = t(sf.free_cols, torch.long) if (sf.free_cols >= 0).any() else None
        # maps of the fused host glue (intopt_gather_convert): standard-form column j takes c_sign_full[j] * c[c_src_full[j]]
        # (-1: slack / split column, cost 0); user variable i takes xhat[x_src[i]] or, if fixed by presolve, its value
        self.c_src_full = t(sf.c_src, torch.int32)
        self.c_sign_full = t(sf.c_sign, torch.float64)
        x_src = np.full(sf.n_x, -1, dtype=np.int32)
        x_src[sf.var_index] = np.arange(sf.nvar, dtype=np.int32)
        self.x_src = t(x_src, torch.int32)
        self.var_index32 = t(sf.var_index, torch.int32)
        self.simple_post = self.x_shift is None and self.x_flip is None and self.free_cols is None
        if inst_values:
            # standard-form entry e of instance i = val_const[e] + val_sign[e] * A_i.flat[val_src[e]]: the entries
            # built from the tagged matrix (|value| >= _TAG0) are copies of user entries, the rest are constants
            tagged = np.abs(val) >= _TAG0 - 0.5
            src = np.where(tagged, np.rint(np.abs(val)) - _TAG0, 0).astype(np.int64)
            self.val_src = t(sf.tag_flatpos[src], torch.long)
            self.val_sign = t(np.where(tagged, np.sign(val), 0.0), torch.float64)
            self.val_const = t(np.where(tagged, 0.0, val), torch.float64)

    def sliced(self, method):
        if method not in self._sliced:
            self._sliced[method] = SlicedProblem(self, method)
        return self._sliced[method]

    def instance_values(self, A_flat):
        """(B, nnz) fp64 values of the standard-form A for the user matrices A_flat (B, n_eq * n_x) on the device."""
        return (self.val_const + self.val_sign * A_flat[:, self.val_src]).contiguous()


def _needs_instance_loop(n_x, A_tagged, G_np, h_np, bounds, presolve):
    """True when presolve of the tagged pattern fixes variables or shifts / flips them: what it then does to b depends on
    the per-instance numbers of A (ip_model_whole.py:1020-1041 singleton rows, :1096-1097 b - A lb)."""
    zero_cols = _problem.zero_columns(A_tagged, G_np, n_x) if presolve else np.zeros(0, dtype=np.int64)
    try:
        sf = _problem.build_standard_form(n_x, A_tagged, None, G_np, h_np, bounds,
                                          None if not zero_cols.size else np.ones(zero_cols.size), presolve=bool(presolve))
    except Exception:  # (decided in presolve, or rejected on the tag numbers: the instances give the real answer)
        return True
    return bool(sf.meta["number_dependent"])


def _instance_loop(variant, A_all, b_np, G, h, kw):
    """IPOfunc for a batch of equality matrices as B independent functions (see make_ipofunc).  Gradients flow through
    torch.stack; the hooks aggregate over the instances."""
    B = A_all.shape[0]
    if kw.get("init_val") is not None:
        raise ValueError("init_val is not supported for a batch of equality matrices with number-dependent presolve")
    if b_np is not None and b_np.ndim == 2 and b_np.shape[0] != B:
        raise ValueError("b has %d rows for %d equality matrices" % (b_np.shape[0], B))
    subs = [make_ipofunc(variant, A_all[i], None if b_np is None else (b_np[i] if b_np.ndim == 2 else b_np), G, h, **kw)
            for i in range(B)]

    def apply(c):
        if c.dim() == 1 or (c.dim() == 2 and c.shape[0] == 1 and B > 1):
            rows = [c.reshape(-1)] * B
        elif c.dim() == 2 and c.shape[0] == B:
            rows = [c[i] for i in range(B)]
        else:
            raise ValueError("c must have one row, or one row per equality matrix (%d)" % B)
        return torch.stack([subs[i][0](rows[i]).reshape(-1) for i in range(B)])

    def status():
        pairs = [sub[1]["status"]() for sub in subs]
        return torch.cat([p[0] for p in pairs]), torch.cat([p[1] for p in pairs])

    def end_vectors():
        ev = [sub[1]["end_vectors"]() for sub in subs]
        if len({tuple(e["x"].shape) for e in ev}) != 1:
            raise RuntimeError("end_vectors(): the instances have different standard forms")
        return {k: torch.cat([e[k] for e in ev]) for k in ev[0]}

    hooks = dict(Runtime=lambda: sum(sub[1]["Runtime"]() for sub in subs), end_vectors=end_vectors,
                 forward_solved=lambda: all(sub[1]["forward_solved"]() for sub in subs), status=status,
                 get_problem=None, instances=subs)
    return apply, hooks


def make_ipofunc(variant, A=None, b=None, G=None, h=None, alpha0=0.9995, beta=0.1, pc=True, tol=1e-8,
                 max_iter=1000, init_val=None, method=1, mu0=None, smoothing=False, bounds=None, thr=None,
                 new_grad=True, presolve=True, A_eq_rr=True, damping=1e-2, check_finite=True):
    """Shared implementation behind ip_model_whole.IPOfunc / ip_model_alter.IPOfunc.

    `pc=False` runs one direction per iteration with centring `beta * mean(x t)` (:1392-1394).  Unused upstream
    keywords are accepted and ignored exactly as upstream does: `mu0`, `new_grad`, `A_eq_rr` are never read, `smoothing` is only logged (ip_model_whole.py:1339).
    `check_finite` (additive): True reproduces upstream's LinAlgError on non-finite `c` (:1332-1341) at the
    price of one device->host sync per call; False leaves it to the per-instance status word (4).
    """
    if variant not in ("whole", "alter"):
        raise ValueError(variant)
    if method not in (1, 2, 3):
        raise ValueError("method must be 1 (full), 2 (rows and columns sliced) or 3 (rows sliced), :1459-1482")
    A_np, b_np, G_np, h_np = _to_numpy(A), _to_numpy(b), _to_numpy(G), _to_numpy(h)
    assert (A_np is not None or G_np is not None) and (b_np is not None or h_np is not None)  # :1209
    if A_np is not None and b_np is None:
        raise ValueError("A was given without b")
    if bounds is None:  # whole :1318, alter ip_model_alter.py:1331
        bounds = [(0.0, None)] if variant == "whole" else [(0.0, 1.0)]
    thr_ = 0.0 if thr is None else float(thr)
    max_iter_ = 1000 if max_iter is None else int(max_iter)
    damping_fwd = 1e-6 if variant == "whole" else float(damping)  # :774 | ip_model_alter.py:775
    mode = 0 if variant == "whole" else 1
    # Additive: A of shape (B, n_eq, n_x) = one equality matrix per instance with a common sparsity pattern (the
    # knapsack driver builds a fresh IPOfunc per minibatch because its equality row is the minibatch's cost vector,
    # experiments/Building Knapsack/optimizer.py:396-405).  The structure is analysed once on a tagged pattern
    # matrix; the numbers travel to the kernels per instance.
    a_batched = A_np is not None and A_np.ndim == 3
    A_all = None
    tag_flatpos = None
    if a_batched:
        A_all = np.ascontiguousarray(A_np, dtype=np.float64)
        pattern = (A_all != 0).any(axis=0)
        tag_flatpos = np.flatnonzero(pattern)
        if tag_flatpos.size == 0:
            raise ValueError("the batch of equality matrices is identically zero")
        A_np = np.zeros(pattern.shape, dtype=np.float64)
        A_np.flat[tag_flatpos] = _TAG0 + np.arange(tag_flatpos.size)
    n_x = (A_np if A_np is not None else G_np).shape[1]
    b_batched = b_np is not None and b_np.ndim == 2
    if a_batched and _needs_instance_loop(n_x, A_np, G_np, h_np, bounds, presolve):
        # Presolve outcomes that depend on the NUMBERS of A (a singleton row fixes x_j = b_i / a_ij; a non-zero lower bound
        # moves b by A lb) cannot be read off the shared pattern.  Such a batch runs the way upstream runs every batch of
        # matrices - one function object per instance - each on the device path; only the batching is lost.
        return _instance_loop(variant, A_all, b_np, G, h, dict(
            alpha0=alpha0, beta=beta, pc=pc, tol=tol, max_iter=max_iter, init_val=init_val, method=method, mu0=mu0,
            smoothing=smoothing, bounds=bounds, thr=thr, new_grad=new_grad, presolve=presolve, A_eq_rr=A_eq_rr,
            damping=damping, check_finite=check_finite))
    a_dev_cache = {}
    structure_key = _digest(A_np, G_np, h_np) + repr([tuple(bd) for bd in bounds]) + ("" if presolve else "|raw")
    zero_cols = _problem.zero_columns(A_np, G_np, n_x) if presolve else np.zeros(0, dtype=np.int64)
    state = {"events": [], "time_ms": 0.0, "last": []}
    closure_id = _Token()  # key of this closure's b-dependent tensors inside a (shared, cached) DeviceProblem

    def get_problem(device, sign_key):
        """Cached DeviceProblem (or the PresolveComplete outcome) for this structure on `device`."""
        key = (structure_key, str(device), sign_key)
        hit = _PROBLEM_CACHE.get(key)
        if hit is None:
            try:
                sf = _problem.build_standard_form(
                    n_x, A_np, None, G_np, h_np, bounds,
                    None if not zero_cols.size else np.asarray(sign_key, dtype=np.float64), presolve=bool(presolve))
                if a_batched:
                    # (number-dependent presolve outcomes were routed to _instance_loop at construction)
                    assert not sf.meta["number_dependent"]
                    sf.tag_flatpos = tag_flatpos
                hit = DeviceProblem(sf, device, inst_values=a_batched)
            except _problem.PresolveComplete as done:
                hit = done
            _PROBLEM_CACHE[key] = hit
            while len(_PROBLEM_CACHE) > _PROBLEM_CACHE_MAX:
                _PROBLEM_CACHE.popitem(last=False)
        else:
            _PROBLEM_CACHE.move_to_end(key)
        return hit

    def b_on_device(dp, device):
        """Standard-form right-hand side(s) (1 or B rows, fp64) and the per-instance infeasibility mask."""
        b_cache = dp.per_closure.setdefault(closure_id, {})  # lives and dies with the DeviceProblem (no id() reuse)
        key = "b"
        if key not in b_cache:
            sf = dp.sf
            bs = np.atleast_2d(sf.b_std(b_np) if sf.n_eq_user else sf.b_const)
            bad = sf.infeasible_mask(b_np) if sf.n_eq_user else np.zeros(1, dtype=bool)
            fixed = np.atleast_2d(sf.fixed_values(b_np) if sf.n_eq_user else sf.fixed_const0)
            b_cache[key] = (torch.as_tensor(np.ascontiguousarray(bs), dtype=torch.float64, device=device),
                            bad,
                            torch.as_tensor(np.ascontiguousarray(fixed), dtype=torch.float64, device=device))
        return b_cache[key]

    def pack_init(dp, B, device):
        if init_val is None:
            return None
        sf = dp.sf
        n, m = sf.n, sf.m
        buf = torch.empty(B, 2 * n + m + 2, dtype=torch.float64, device=device)
        for lo, hi, k in ((0, n, "x"), (n, n + m, "y"), (n + m, 2 * n + m, "t"), (2 * n + m, 2 * n + m + 1, "tau"),
                          (2 * n + m + 1, 2 * n + m + 2, "kappa")):
            v = init_val[k]
            v = v.detach() if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))
            buf[:, lo:hi] = v.to(device=device, dtype=torch.float64).reshape(-1, hi - lo)
        return buf

    def solve_group(dp, c2, sel, out_dtype):
        """c2: (Bg, n_x) costs on device (float or double).  Returns x_user (Bg, n_x) and the tensors backward needs."""
        sf = dp.sf
        L = _capi.lib()
        Bg = c2.shape[0]
        device = c2.device
        stream = torch.cuda.current_stream(device).cuda_stream
        um = _capi.IntoptUserMap()  # all-null: no mapping inside the kernel
        map_in = _glue_ok(c2)
        if map_in:  # the kernel reads the caller's c itself: cast, sign, scatter into the standard-form columns
            c_std = torch.empty(Bg, sf.n, dtype=torch.float64, device=device)  # (filled by the kernel, kept for backward)
            um.in_dev, um.in_is_f32, um.in_stride = c2.data_ptr(), int(c2.dtype == torch.float32), c2.stride(0)
            um.in_idx_dev, um.in_scale_dev = dp.c_src_full.data_ptr(), dp.c_sign_full.data_ptr()
        else:
            c2 = c2.double()
            c_std = torch.zeros(Bg, sf.n, dtype=torch.float64, device=device)
            if dp.c_plain:
                c_std[:, :sf.nvar] = c2 if dp.identity_vars else c2[:, dp.c_src]
            else:
                c_std[:, dp.c_cols] = c2[:, dp.c_src] * dp.c_sign
        b_all, bad, fixed_all = b_on_device(dp, device)
        if b_all.shape[0] == 1:
            b_std, b_stride, fixed = b_all, 0, fixed_all
        else:
            b_std = b_all if sel is None else b_all[sel].contiguous()
            fixed = fixed_all if sel is None else fixed_all[sel]
            b_stride = sf.m
            if b_std.shape[0] != Bg:
                raise ValueError("b has %d instances but c has %d" % (b_std.shape[0], Bg))
        a_vals = None
        if a_batched:
            if str(device) not in a_dev_cache:
                a_dev_cache[str(device)] = torch.as_tensor(A_all.reshape(A_all.shape[0], -1), device=device)
            A_flat = a_dev_cache[str(device)]
            if sel is None and A_flat.shape[0] != Bg:
                raise ValueError("A has %d instances but c has %d" % (A_flat.shape[0], Bg))
            a_vals = dp.instance_values(A_flat if sel is None else A_flat[sel])
        map_out = dp.simple_post and fixed.shape[0] == 1 and not bad.any()
        x_user = None
        if map_out:
            # the drivers' case (lb = 0, no upper-bound flips, no free variables, fixed values shared by the batch): the
            # kernel writes x in the caller's layout and dtype, fixed variables re-inserted (_postprocess :1241-1285)
            x_user = torch.empty(Bg, n_x, dtype=out_dtype, device=device)
            b_cache = dp.per_closure.setdefault(closure_id, {})
            xkey = "xconst"
            if xkey not in b_cache:  # values of the presolve-fixed variables in user positions (constant per problem)
                xc = torch.zeros(n_x, dtype=torch.float64, device=device)
                if sf.fixed_idx.size:
                    xc[dp.fixed_idx] = fixed[0]
                b_cache[xkey] = xc
            um.out_dev, um.out_is_f32, um.out_stride, um.n_out = x_user.data_ptr(), int(out_dtype == torch.float32), n_x, n_x
            um.out_idx_dev, um.out_const_dev = dp.x_src.data_ptr(), b_cache[xkey].data_ptr()
        xhat = None if map_out else torch.empty(Bg, sf.n, dtype=torch.float64, device=device)
        st = torch.empty(Bg, 2 * sf.n + sf.m + 2, dtype=torch.float64, device=device)
        status = torch.empty(Bg, dtype=torch.int32, device=device)
        iters = torch.empty(Bg, dtype=torch.int32, device=device)
        ws = dp.workspace(Bg)
        init = pack_init(dp, Bg, device)
        opts = _capi.IntoptOpts(alpha0, tol, thr_, damping_fwd, max_iter_, 0 if pc else 1, float(beta))
        # (solve_group runs under torch.cuda.device(c.device): the library launches on the CURRENT device)
        _capi.check(L.intopt_forward_mapped(dp.handle, Bg, _ptr(c_std), sf.n, _ptr(b_std), b_stride, _ptr(a_vals),
                                            dp.nnz if a_batched else 0, ctypes.byref(opts), _ptr(init), ctypes.byref(um),
                                            _ptr(xhat), _ptr(st), _ptr(status), _ptr(iters), _ptr(ws), ws.numel() * 8,
                                            ctypes.c_void_p(stream)), "intopt_forward")
        if map_out:
            return x_user, (c_std, b_std, b_stride, st, status, iters, a_vals)
        # _postprocess (:1241-1285): undo substitutions, re-insert fixed variables
        xk = xhat[:, :sf.nvar]
        if dp.free_cols is not None:
            neg = torch.where(dp.free_cols >= 0, xhat[:, dp.free_cols.clamp(min=0)], torch.zeros_like(xk))
            xk = xk - neg
        if dp.x_flip is not None:
            xk = torch.where(torch.isnan(dp.x_flip), xk, dp.x_flip - xk)
        if dp.x_shift is not None:
            xk = xk + dp.x_shift
        if dp.identity_vars:
            x_user = xk
        else:
            x_user = torch.zeros(Bg, n_x, dtype=torch.float64, device=device)
            x_user[:, dp.var_index] = xk
            if sf.fixed_idx.size:
                x_user[:, dp.fixed_idx] = fixed
        if bad.any():  # presolve-infeasible instances return the presolve x (zeros) upstream (:1483-1488)
            badsel = bad if sel is None or bad.size == 1 else bad[sel.cpu().numpy()]
            x_user[torch.as_tensor(np.broadcast_to(badsel, (Bg,)).copy(), device=device)] = 0.0
        return x_user, (c_std, b_std, b_stride, st, status, iters, a_vals)

    def _tick():
        """Start a timed region for Runtime(): folds the pairs that have completed into the total first, so the list
        of pending CUDA events stays short however long one IPOfunc lives."""
        if torch.cuda.is_current_stream_capturing():
            return None  # inside a CUDA-graph capture: no timing events (they could not be read back anyway)
        ev = state["events"]
        while ev and ev[0][1].query():
            e0, e1 = ev.pop(0)
            state["time_ms"] += e0.elapsed_time(e1)
        if len(ev) >= 256:  # never let it grow: wait for the oldest pair
            e0, e1 = ev.pop(0)
            e1.synchronize()
            state["time_ms"] += e0.elapsed_time(e1)
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def _tock(e0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        state["events"].append((e0, e1))

    class IPOfunc_cls(torch.autograd.Function):
        @staticmethod
        def forward(ctx, c):
            if not isinstance(c, torch.Tensor) or not c.is_cuda:
                raise RuntimeError("IPOfunc (B200): c must be a CUDA tensor; there is no CPU fallback")
            # every launch below (kernels, the lazily built cold-start factor, the timing events) goes to c's device,
            # whatever the caller's current device is
            with torch.cuda.device(c.device):
                return IPOfunc_cls._forward(ctx, c)

        @staticmethod
        def _forward(ctx, c):
            device = c.device
            ev0 = _tick()
            squeeze = c.dim() == 1
            c2 = c.detach().reshape(-1, c.shape[-1])  # (kept in the caller's dtype: the pack kernel converts)
            if c2.dtype not in (torch.float32, torch.float64):
                c2 = c2.double()
            if c2.shape[1] != n_x:
                raise ValueError("c must have %d entries per instance" % n_x)
            if check_finite and not bool(torch.isfinite(c2).all()):
                raise LinAlgError("c contains inf, nan, or None")  # :1332-1341
            B = c2.shape[0]
            ctx.c_shared = False
            if b_batched and B == 1 and b_np.shape[0] > 1:  # one shared c, per-instance b (Twitter)
                B = b_np.shape[0]
                c2 = c2.expand(B, -1)
                squeeze = False
                ctx.c_shared = True  # (1-D or (1, n_x) c: the gradients of the instances add up in backward)
            if B == 0:  # nothing to solve: no launch, empty result, empty gradient
                ctx.saved, ctx.in_shape, ctx.B, ctx.in_dtype = [], c.shape, 0, c.dtype
                _tock(ev0)
                return torch.empty_like(c2, dtype=c.dtype)
            # group by the sign pattern of c on unconstrained columns (presolve zero-column rule, :995-1016)
            if zero_cols.size:
                signs = torch.sign(c2[:, torch.as_tensor(zero_cols, device=device)]).to(torch.int8).cpu().numpy()
                keys, inverse = np.unique(signs, axis=0, return_inverse=True)
                inverse = np.asarray(inverse).reshape(-1)
                groups = [(tuple(int(v) for v in k), np.nonzero(inverse == gi)[0]) for gi, k in enumerate(keys)]
            else:
                groups = [((), None)]
            single = len(groups) == 1
            x_out = None if single else torch.empty(B, n_x, dtype=torch.float64, device=device)
            saved = []
            for sign_key, idx in groups:
                dp = get_problem(device, sign_key)
                sel = None if single else torch.as_tensor(idx, device=device)
                cg = c2 if single else c2[sel]
                if isinstance(dp, _problem.PresolveComplete):
                    # decided in presolve: upstream returns presolve's x and saves nothing (:1483-1488);
                    # the gradient of such instances is zero here
                    if dp.trivial is not None:  # x follows the sign of c (:1136-1142), instance by instance
                        vi = torch.as_tensor(dp.trivial[0], dtype=torch.long, device=device)
                        lbt = torch.as_tensor(dp.trivial[1], dtype=torch.float64, device=device)
                        ubt = torch.as_tensor(dp.trivial[2], dtype=torch.float64, device=device)
                        cv = cg.double()[:, vi]
                        zero_c = torch.where(torch.isinf(ubt), torch.zeros_like(ubt), ubt)
                        xg = torch.as_tensor(dp.x, dtype=torch.float64, device=device).repeat(cg.shape[0], 1)
                        if dp.fixed is not None and dp.fixed[0].size and b_np is not None:  # singleton rows: b / a
                            fv = dp.fixed[1] + (dp.fixed[2] @ np.atleast_2d(b_np).T).T  # (1 or B, n_fixed)
                            fv = fv if fv.shape[0] == 1 or sel is None else fv[idx]
                            xg[:, torch.as_tensor(dp.fixed[0], dtype=torch.long, device=device)] = torch.as_tensor(
                                np.ascontiguousarray(fv), dtype=torch.float64, device=device)
                        xg[:, vi] = torch.where(cv < 0, ubt, torch.where(cv > 0, lbt, zero_c))
                    else:
                        xg = torch.as_tensor(dp.x, dtype=torch.float64, device=device).expand(cg.shape[0], -1)
                    if single:
                        xg = xg.to(c.dtype)
                    saved.append((None, sel, None))
                else:
                    # (an expanded shared c keeps its stride 0: the pack kernel reads the one row for every instance)
                    xg, pack = solve_group(dp, cg if (single and cg.stride(1) == 1) else cg.contiguous(), sel,
                                           c.dtype if (single and c.dtype in (torch.float32, torch.float64)) else torch.float64)
                    saved.append((dp, sel, pack))
                if single:
                    x_out = xg
                else:
                    x_out[sel] = xg
            ctx.saved = saved
            ctx.in_shape = c.shape
            ctx.B = B
            ctx.in_dtype = c.dtype
            state["last"] = saved
            _tock(ev0)
            out = x_out.to(c.dtype)
            return out[0] if squeeze else out

        @staticmethod
        def backward(ctx, del_x):
            with torch.cuda.device(del_x.device):
                return IPOfunc_cls._backward(ctx, del_x)

        @staticmethod
        def _backward(ctx, del_x):
            ev0 = _tick()
            L = _capi.lib()
            device = del_x.device
            B = ctx.B
            if B == 0:
                return torch.zeros(ctx.in_shape, dtype=ctx.in_dtype, device=device)
            g2 = del_x.detach().reshape(-1, n_x)
            if g2.dtype not in (torch.float32, torch.float64):
                g2 = g2.double()
            if g2.shape[0] != B:
                g2 = g2.expand(B, -1)
            single = len(ctx.saved) == 1
            grad = None if single else torch.zeros(B, n_x, dtype=torch.float64, device=device)
            for dp, sel, pack in ctx.saved:
                if dp is None:
                    if single:
                        grad = torch.zeros(B, n_x, dtype=torch.float64, device=device)
                    continue
                sf = dp.sf
                c_std, b_std, b_stride, st, status, iters, a_vals = pack
                gg = g2 if sel is None else g2[sel]
                if gg.stride(1) != 1 and gg.stride(0) != 0:
                    gg = gg.contiguous()
                Bg = gg.shape[0]
                stream = torch.cuda.current_stream(device).cuda_stream
                fast = sel is None and gg.stride(1) == 1 and ctx.in_dtype in (torch.float32, torch.float64)
                prob, n_std = dp, sf.n
                if method != 1:
                    sub = dp.sliced(method)
                    if not sub.trivial:  # the backward of methods 2 / 3 is the same kernel on the sliced problem
                        prob, n_std = sub, sub.n
                        c_std, b_std, st, a_vals = sub.slices(sf, c_std, b_std, st, a_vals)
                        b_stride = sub.m if b_stride else 0
                um = _capi.IntoptUserMap()
                if fast:  # the kernel gathers del_x[var_index] (:1518) itself, as doubles
                    gk = torch.empty(Bg, sf.nvar, dtype=torch.float64, device=device)
                    um.in_dev, um.in_is_f32, um.in_stride = gg.data_ptr(), int(gg.dtype == torch.float32), gg.stride(0)
                    um.in_idx_dev = dp.var_index32.data_ptr()
                else:
                    gk = (gg if dp.identity_vars else gg[:, dp.var_index]).double().contiguous()
                fast_out = fast and single and not ctx.c_shared
                gfull = None
                if fast_out:  # ... and writes c_grad[var_index] = dc, zeros elsewhere (:1559-1560), in the caller's dtype
                    gfull = torch.empty(Bg, n_x, dtype=ctx.in_dtype, device=device)
                    um.out_dev, um.out_is_f32, um.out_stride, um.n_out = (gfull.data_ptr(), int(ctx.in_dtype == torch.float32),
                                                                           n_x, n_x)
                    um.out_idx_dev = dp.x_src.data_ptr()
                gc = torch.empty(Bg, sf.nvar, dtype=torch.float64, device=device)
                bstat = torch.empty(Bg, dtype=torch.int32, device=device)
                ws = prob.workspace(Bg)
                _capi.check(L.intopt_backward_mapped(prob.handle, Bg, mode, float(damping), sf.nvar, _ptr(st),
                                                     _ptr(c_std), n_std, _ptr(b_std), b_stride, _ptr(a_vals),
                                                     prob.nnz if a_vals is not None else 0, ctypes.byref(um), _ptr(gk),
                                                     _ptr(gc), _ptr(bstat), _ptr(ws), ws.numel() * 8,
                                                     ctypes.c_void_p(stream)),
                            "intopt_backward")
                if fast_out:
                    pass
                elif dp.identity_vars:
                    gfull = gc
                else:
                    gfull = torch.zeros(Bg, n_x, dtype=torch.float64, device=device)
                    gfull[:, dp.var_index] = gc  # c_grad[var_index] = dc (:1559-1560)
                if single:
                    grad = gfull
                else:
                    grad[sel] = gfull
            if ctx.c_shared:  # one shared c, per-instance b: the gradients add up
                grad = grad.sum(dim=0)
            _tock(ev0)
            return grad.reshape(ctx.in_shape).to(ctx.in_dtype)

    def Runtime():
        """Accumulated device time of forward+backward in seconds (CUDA events; upstream: wall clock,
        ip_model_whole.py:1291,1498-1499,1574-1575).  Synchronises."""
        torch.cuda.synchronize()
        for e0, e1 in state["events"]:
            state["time_ms"] += e0.elapsed_time(e1)
        state["events"] = []
        return state["time_ms"] * 1e-3

    def _last_packs():
        return [(dp, sel, pack) for dp, sel, pack in state["last"] if dp is not None]

    def status():
        """(status, iters) int32 tensors of the last forward: 0 ok, 1 max_iter, 2 infeasible/unbounded
        indicators, 3 factorisation breakdown, 4 non-finite c."""
        packs = _last_packs()
        if len(packs) == 1 and packs[0][1] is None:
            return packs[0][2][4], packs[0][2][5]
        raise RuntimeError("status(): the last batch was split by presolve sign patterns")

    def forward_solved():
        """Upstream always returns True (dead plumbing); here: every instance of the last
        forward left the loop through its own `go` test.  Synchronises."""
        return all(bool((pack[4] == 0).all()) for _, _, pack in _last_packs())

    def end_vectors():
        """Final iterate of the last forward as an `init_val` dict (upstream's raises NameError)."""
        packs = _last_packs()
        if len(packs) != 1 or packs[0][1] is not None:
            raise RuntimeError("end_vectors(): the last batch was split by presolve sign patterns")
        dp, _, pack = packs[0]
        n, m, st = dp.sf.n, dp.sf.m, pack[3]
        return {"x": st[:, :n], "y": st[:, n:n + m], "t": st[:, n + m:2 * n + m], "tau": st[:, 2 * n + m],
                "kappa": st[:, 2 * n + m + 1]}

    hooks = dict(Runtime=Runtime, end_vectors=end_vectors, forward_solved=forward_solved, status=status,
                 get_problem=get_problem)
    return IPOfunc_cls.apply, hooks
