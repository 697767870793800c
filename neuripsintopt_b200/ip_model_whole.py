"""Drop-in for the reference's `Interior/ip_model_whole.py`: `IPOfunc` with the HSD-embedding backward.

Same name, argument order, defaults and module-level hooks as ip_model_whole.py:1287-1290,1579-1592
(`IPOfunc.Runtime()`, `IPOfunc.end_vectors()`, `IPOfunc.forward_solved()` are set on the function object
at every construction, exactly like upstream, because the Energy driver reads them right after each call,
Interior/intopt_energy_mlp.py:431-432).  Forward damping of the normal matrix is the hard-coded 1e-6
of :774; `damping` only enters the backward (:1509,1535).
"""
from .ipofunc import make_ipofunc


def IPOfunc(A=None, b=None, G=None, h=None, alpha0=0.9995, beta=0.1, pc=True,
            tol=1e-8, max_iter=1000, init_val=None, method=1, mu0=None,
            smoothing=False, bounds=None, thr=None, new_grad=True, presolve=True, A_eq_rr=True,
            damping=1e-2, check_finite=True):
    fn, hooks = make_ipofunc("whole", A, b, G, h, alpha0=alpha0, beta=beta, pc=pc, tol=tol, max_iter=max_iter,
                             init_val=init_val, method=method, mu0=mu0, smoothing=smoothing, bounds=bounds,
                             thr=thr, new_grad=new_grad, presolve=presolve, A_eq_rr=A_eq_rr, damping=damping,
                             check_finite=check_finite)
    IPOfunc.Runtime = hooks["Runtime"]
    IPOfunc.end_vectors = hooks["end_vectors"]
    IPOfunc.forward_solved = hooks["forward_solved"]
    IPOfunc.status = hooks["status"]
    return fn
