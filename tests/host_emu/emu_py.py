"""TEST-ONLY ctypes wrapper over tests/host_emu/libintopt_emu.so (the solver's device code compiled as
single-threaded C++).  Used by CPU tests to check the kernel arithmetic without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libintopt_emu.so")


def build():
    src = os.path.join(HERE, "emu.cpp")
    deps = [src, os.path.join(HERE, "../../neuripsintopt_b200/csrc/intopt_solver.cuh"),
            os.path.join(HERE, "../../neuripsintopt_b200/csrc/intopt_structure.hpp"),
            os.path.join(HERE, "../../neuripsintopt_b200/csrc/intopt_bind.hpp")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, src])
    return LIB


def solve(A, b, c, g=None, b_batch=None, alpha0=0.9995, tol=1e-8, thr=0.0, max_iter=1000, damping_fwd=1e-6,
          mode=0, damping_bwd=1e-2, nvar=None, force_order=-1, init=None, a_vals=None, plain=0, beta=0.1):
    lib = ctypes.CDLL(build())
    A = sp.csr_matrix(A)
    A.sort_indices()
    m, n = A.shape
    c = np.ascontiguousarray(np.atleast_2d(c), dtype=np.float64)
    B = c.shape[0] if b_batch is None else len(b_batch)
    c_stride = n if c.shape[0] == B else 0
    nvar = n if nvar is None else nvar
    dp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rp = A.indptr.astype(np.int32)
    ci = A.indices.astype(np.int32)
    val = A.data.astype(np.float64)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    bbatch = None if b_batch is None else np.ascontiguousarray(b_batch, dtype=np.float64)
    xhat = np.zeros((B, n))
    state = np.zeros((B, 2 * n + m + 2))
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    bstatus = np.zeros(B, dtype=np.int32)
    gx = None if g is None else np.ascontiguousarray(np.atleast_2d(g), dtype=np.float64)
    gc = None if g is None else np.zeros((B, nvar))
    ini = None if init is None else np.ascontiguousarray(init, dtype=np.float64)
    av = None if a_vals is None else np.ascontiguousarray(a_vals, dtype=np.float64)  # B x nnz, CSR order of A
    env = ctypes.c_int(0)
    rc = lib.emu_solve(ctypes.c_int(m), ctypes.c_int(n), dp(rp), dp(ci), dp(val), dp(bb), ctypes.c_int(force_order),
                       ctypes.c_int(B), dp(c), ctypes.c_longlong(c_stride), dp(bbatch), ctypes.c_double(alpha0),
                       ctypes.c_double(tol), ctypes.c_double(thr), ctypes.c_double(damping_fwd),
                       ctypes.c_int(max_iter), dp(ini), dp(xhat), dp(state), dp(status), dp(iters),
                       ctypes.c_int(mode), ctypes.c_double(damping_bwd), ctypes.c_int(nvar), dp(gx), dp(gc),
                       dp(bstatus), ctypes.byref(env), dp(av), ctypes.c_int(plain), ctypes.c_double(beta))
    assert rc == 0, rc
    return dict(xhat=xhat, state=state, status=status, iters=iters, gc=gc, bstatus=bstatus, envelope=env.value)
