"""TEST INFRASTRUCTURE — not part of the product path.

Runs the *untouched* reference implementation (JayMan91/NeurIPSIntopt) from
/root/reference so that golden vectors can be generated from it.  This only
works in the build container (the GPU box has no /root/reference); nothing in
`tests/ -m gpu`, `smoke()` or `bench.py` imports this file.

Two compatibility edits are applied by monkey-patching, never by editing or
copying reference sources:

  * `np.float` was removed from NumPy >= 1.24; the reference uses it at
    Interior/ip_model_whole.py:39,42,44,63,64,130,650-654.
  * `_get_step` (Interior/ip_model_whole.py:870-878) builds a ragged list for
    `np.min` because tau/kappa are shape-(1,) arrays; modern NumPy rejects it.
    The replacement below is the same arithmetic with the scalars squeezed.
"""
import ast
import os
import sys
import warnings

import numpy as np

REF_ROOT = os.environ.get("INTOPT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "Interior", "ip_model_whole.py"))


def _get_step(x, d_x, t, d_t, tau, d_tau, kappa, d_kappa, alpha0):
    # Same maths as Interior/ip_model_whole.py:870-878, scalar-safe.
    i_x = d_x < 0
    i_t = d_t < 0
    tau_ = float(np.squeeze(tau))
    kappa_ = float(np.squeeze(kappa))
    d_tau_ = float(np.squeeze(d_tau))
    d_kappa_ = float(np.squeeze(d_kappa))
    alpha_x = alpha0 * np.min(x[i_x] / -d_x[i_x]) if np.any(i_x) else 1
    alpha_tau = alpha0 * tau_ / -d_tau_ if d_tau_ < 0 else 1
    alpha_t = alpha0 * np.min(t[i_t] / -d_t[i_t]) if np.any(i_t) else 1
    alpha_kappa = alpha0 * kappa_ / -d_kappa_ if d_kappa_ < 0 else 1
    return np.min([1.0, float(alpha_x), float(alpha_tau), float(alpha_t), float(alpha_kappa)])


_MODULES = {}


def load(which: str = "whole"):
    """Import Interior/ip_model_whole.py or ip_model_alter.py with the shim applied."""
    if which in _MODULES:
        return _MODULES[which]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not hasattr(np, "float"):
        np.float = float  # noqa: NPY001  (compat shim, see module docstring)
    interior = os.path.join(REF_ROOT, "Interior")
    if interior not in sys.path:
        sys.path.insert(0, interior)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if which == "whole":
            import ip_model_whole as mod  # type: ignore
        elif which == "alter":
            import ip_model_alter as mod  # type: ignore
        else:
            raise ValueError(which)
    mod._get_step = _get_step
    _MODULES[which] = mod
    return mod


def lift_functions(relpath: str, names):
    """exec() selected top-level FunctionDefs of a reference file without importing
    the module (the drivers import gurobipy, which is absent).  Nothing is written."""
    import torch  # the lifted builders use torch

    path = os.path.join(REF_ROOT, relpath)
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def run_ipofunc(which, c, g, A=None, b=None, G=None, h=None, **kw):
    """Call the reference IPOfunc exactly like its drivers do
    (Interior/intopt_energy_mlp.py:422-428): x = IPOfunc(...)(c); (x*g).sum().backward().

    Returns (x float32 ndarray, dL/dc float32 ndarray, hooks dict with the fp64
    iterates captured from inside the reference run)."""
    import torch

    mod = load(which)
    cap = {"iters": 0}
    orig_do_step = mod._do_step

    def spy_do_step(x, y, t, tau, kappa, d_x, d_y, d_t, d_tau, d_kappa, alpha):
        out = orig_do_step(x, y, t, tau, kappa, d_x, d_y, d_t, d_tau, d_kappa, alpha)
        cap["iters"] += 1
        cap["state"] = tuple(np.array(v, dtype=np.float64).copy() for v in out)
        cap.setdefault("alphas", []).append(float(alpha))
        return out

    orig_get_abc = mod._get_Abc

    def spy_get_abc(*a, **k):
        out = orig_get_abc(*a, **k)
        cap["A_std"], cap["b_std"], cap["c_std"] = (np.array(out[0]).copy(), np.array(out[1]).copy(),
                                                   np.array(out[2]).copy())
        return out

    mod._do_step = spy_do_step
    mod._get_Abc = spy_get_abc
    try:
        def tt(v):
            if v is None:
                return torch.Tensor()
            return torch.as_tensor(np.asarray(v), dtype=torch.float32)

        c_t = torch.tensor(np.asarray(c), dtype=torch.float32, requires_grad=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f = mod.IPOfunc(tt(A), tt(b), tt(G), tt(h), **kw)
            x = f(c_t)
            (x * torch.as_tensor(np.asarray(g), dtype=torch.float32)).sum().backward()
    finally:
        mod._do_step = orig_do_step
        mod._get_Abc = orig_get_abc
    return x.detach().numpy().copy(), c_t.grad.detach().numpy().copy(), cap
