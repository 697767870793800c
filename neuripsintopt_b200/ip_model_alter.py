"""Drop-in for the reference's `Interior/ip_model_alter.py`: `IPOfunc` with the log-barrier backward.

Differences from ip_model_whole kept as upstream has them: `damping` sits before `smoothing` in the
signature (ip_model_alter.py:1288-1289), it also damps the *forward* normal matrix (:775,1408-1409),
the default bounds are (0, 1) (:1331), the backward uses the iterate normalised by tau (:1459-1461)
and mu = x.t/n (:1525).
"""
from .ipofunc import make_ipofunc


def IPOfunc(A=None, b=None, G=None, h=None, alpha0=0.9995, beta=0.1, pc=True,
            tol=1e-8, max_iter=1000, init_val=None, method=1, mu0=None, damping=1e-2,
            smoothing=False, bounds=None, thr=None, new_grad=True, presolve=True, A_eq_rr=True,
            check_finite=True):
    fn, hooks = make_ipofunc("alter", A, b, G, h, alpha0=alpha0, beta=beta, pc=pc, tol=tol, max_iter=max_iter,
                             init_val=init_val, method=method, mu0=mu0, smoothing=smoothing, bounds=bounds,
                             thr=thr, new_grad=new_grad, presolve=presolve, A_eq_rr=A_eq_rr, damping=damping,
                             check_finite=check_finite)
    IPOfunc.Runtime = hooks["Runtime"]
    IPOfunc.end_vectors = hooks["end_vectors"]
    IPOfunc.forward_solved = hooks["forward_solved"]
    IPOfunc.status = hooks["status"]
    return fn
