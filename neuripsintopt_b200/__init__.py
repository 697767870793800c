"""neuripsintopt_b200 — B200-native batched differentiable interior-point LP layer.

A from-scratch replacement for the hot path of JayMan91/NeurIPSIntopt (`IPOfunc`, Interior/
ip_model_whole.py / ip_model_alter.py): the homogeneous self-dual predictor-corrector forward and the
KKT backward returning dL/dc, batched over a minibatch of cost vectors, fp64, hand-written sm_100a CUDA
behind the C ABI of include/intopt_b200.h.

    from neuripsintopt_b200.ip_model_whole import IPOfunc      # same call as the reference
    x = IPOfunc(A, b, G, h, bounds=[(0., None)], thr=0.1, damping=1e-6)(c)     # c: (n,) or (B, n), CUDA
"""
from . import _capi  # noqa: F401
from .ip_model_whole import IPOfunc  # noqa: F401

__all__ = ["IPOfunc", "ip_model_whole", "ip_model_alter"]
