"""ctypes binding of the C ABI declared in include/intopt_b200.h (libintopt_b200.so).

This is the reference-side stub a maintainer would add: the reference is pure
Python with no FFI, so the binding is ctypes over plain pointers and sizes — no torch types cross the
boundary.  There is deliberately no fallback: if the CUDA library is missing or no device is usable
every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("INTOPT_LIB_OVERRIDE") or os.path.join(_HERE, "lib", "libintopt_b200.so")  # override: A/B tuning only


class IntoptOpts(ctypes.Structure):
    _fields_ = [("alpha0", ctypes.c_double), ("tol", ctypes.c_double), ("thr", ctypes.c_double),
                ("damping_fwd", ctypes.c_double), ("max_iter", ctypes.c_int), ("plain", ctypes.c_int),
                ("beta", ctypes.c_double)]


class IntoptUserMap(ctypes.Structure):
    _fields_ = [("in_dev", ctypes.c_void_p), ("in_is_f32", ctypes.c_int), ("in_stride", ctypes.c_longlong),
                ("in_idx_dev", ctypes.c_void_p), ("in_scale_dev", ctypes.c_void_p), ("out_dev", ctypes.c_void_p),
                ("out_is_f32", ctypes.c_int), ("out_stride", ctypes.c_longlong), ("n_out", ctypes.c_int),
                ("out_idx_dev", ctypes.c_void_p), ("out_const_dev", ctypes.c_void_p)]


class IntoptInfo(ctypes.Structure):
    _fields_ = [("m", ctypes.c_int), ("n", ctypes.c_int), ("nnz", ctypes.c_int), ("envelope", ctypes.c_int),
                ("dense_lower", ctypes.c_int), ("asm_terms", ctypes.c_longlong), ("chol_macs", ctypes.c_longlong),
                ("smem_bytes", ctypes.c_int), ("factor_in_smem", ctypes.c_int), ("threads", ctypes.c_int),
                ("ordering", ctypes.c_int), ("gather_in_smem", ctypes.c_int), ("ctas_per_sm", ctypes.c_int),
                ("scalar_envelope", ctypes.c_int), ("path", ctypes.c_int), ("asm_padded", ctypes.c_longlong),
                ("solve_macs", ctypes.c_longlong), ("tmem", ctypes.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["intopt_version", "intopt_last_error", "intopt_analyze", "intopt_problem_create",
           "intopt_problem_destroy", "intopt_problem_info", "intopt_workspace_bytes", "intopt_forward",
           "intopt_backward", "intopt_forward_values", "intopt_backward_values", "intopt_forward_mapped",
           "intopt_backward_mapped", "intopt_gather_convert",
           "intopt_dense_normal_solve", "intopt_measure_dmma_peak", "intopt_measure_fp64_peak"]
# only in the profiling build (csrc/build.sh profile -> lib/libintopt_b200_profile.so, include/intopt_b200_profile.h)
PROFILE_EXPORTS = ["intopt_debug_phase_cycles", "intopt_debug_microbench"]
PROFILE_LIB_PATH = os.path.join(_HERE, "lib", "libintopt_b200_profile.so")
FLAG_INSTANCE_VALUES = 1

_lib = None


def build(verbose=False, profile=False):
    """Compile libintopt_b200.so for sm_100a with nvcc (cross-compiles without a GPU).  `profile=True` builds the
    separate profiling library instead (phase-cycle probes + micro-benchmarks; never loaded by the product path)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    out = subprocess.run(["bash", script] + (["profile"] if profile else []), capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc build of libintopt_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout.strip())
    return LIB_PATH


def lib():
    """Load the shared library (once) and declare the prototypes of include/intopt_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libintopt_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or neuripsintopt_b200/csrc/build.sh. There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd, cll, csz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_longlong, ctypes.c_size_t
    L.intopt_version.restype = ci
    L.intopt_last_error.restype = ctypes.c_char_p
    L.intopt_analyze.argtypes = [ci, ci, vp, vp, vp, ctypes.POINTER(IntoptInfo), vp]
    L.intopt_problem_create.argtypes = [ctypes.POINTER(vp), ci, ci, vp, vp, vp, vp, ci]
    L.intopt_problem_destroy.argtypes = [vp]
    L.intopt_problem_info.argtypes = [vp, ctypes.POINTER(IntoptInfo)]
    L.intopt_workspace_bytes.argtypes = [vp, ci]
    L.intopt_workspace_bytes.restype = csz
    L.intopt_forward.argtypes = [vp, ci, vp, cll, vp, cll, ctypes.POINTER(IntoptOpts), vp, vp, vp, vp, vp, vp, csz, vp]
    L.intopt_backward.argtypes = [vp, ci, ci, cd, ci, vp, vp, cll, vp, cll, vp, vp, vp, vp, csz, vp]
    L.intopt_forward_values.argtypes = [vp, ci, vp, cll, vp, cll, vp, cll, ctypes.POINTER(IntoptOpts), vp, vp, vp, vp, vp,
                                        vp, csz, vp]
    L.intopt_backward_values.argtypes = [vp, ci, ci, cd, ci, vp, vp, cll, vp, cll, vp, cll, vp, vp, vp, vp, csz, vp]
    L.intopt_gather_convert.argtypes = [vp, ci, cll, ci, ci, vp, vp, vp, vp, ci, cll, vp]
    um = ctypes.POINTER(IntoptUserMap)
    L.intopt_forward_mapped.argtypes = [vp, ci, vp, cll, vp, cll, vp, cll, ctypes.POINTER(IntoptOpts), vp, um, vp, vp, vp, vp,
                                        vp, csz, vp]
    L.intopt_backward_mapped.argtypes = [vp, ci, ci, cd, ci, vp, vp, cll, vp, cll, vp, cll, um, vp, vp, vp, vp, csz, vp]
    L.intopt_measure_fp64_peak.argtypes = [ctypes.POINTER(cd), ci]
    L.intopt_measure_dmma_peak.argtypes = [ctypes.POINTER(cd), ci]
    L.intopt_dense_normal_solve.argtypes = [vp, ci, vp, cd, vp, vp, vp, csz, vp]
    for name in EXPORTS:
        getattr(L, name)  # AttributeError here means the library and the header disagree
    if hasattr(L, "intopt_debug_phase_cycles"):  # profiling build loaded through INTOPT_LIB_OVERRIDE
        L.intopt_debug_phase_cycles.argtypes = [ci, vp]
        L.intopt_debug_microbench.argtypes = [vp, ci, ci, ci, vp]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().intopt_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def analyze(csr):
    """Host-only structure analysis of a scipy CSR matrix -> (info dict, row permutation)."""
    import numpy as np
    L = lib()
    m, n = csr.shape
    rp = np.ascontiguousarray(csr.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(csr.indices, dtype=np.int32)
    val = np.ascontiguousarray(csr.data, dtype=np.float64)
    info = IntoptInfo()
    perm = np.zeros(m, dtype=np.int32)
    check(L.intopt_analyze(m, n, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, ctypes.byref(info),
                           perm.ctypes.data), "intopt_analyze")
    return info.as_dict(), perm


def measure_dmma_peak(repeats=5):
    """FP64 tensor-core (DMMA) throughput of the current device in GFLOP/s (roofline denominator of the dense path)."""
    out = ctypes.c_double(0.0)
    check(lib().intopt_measure_dmma_peak(ctypes.byref(out), repeats), "intopt_measure_dmma_peak")
    return out.value


def measure_fp64_peak(repeats=5):
    out = ctypes.c_double(0.0)
    check(lib().intopt_measure_fp64_peak(ctypes.byref(out), repeats), "intopt_measure_fp64_peak")
    return out.value


PHASES = ["setup", "residuals+tests", "assembly", "cholesky", "rhs/spmv", "solve2", "p/u loop", "corrector rhs",
          "solve1", "direction/ratio/update", "epilogue", "chol:panel", "chol:lead-update(w0)", "chol:diag(w0)",
          "chol:wait-others", "-", "resid:n-loop(A^T y)", "resid:spmv(A x)", "rhs:spmv(q)", "rhs:predictor loop",
          "rhs:spmv(v)", "corr:n-loop", "corr:spmv(v)", "dir:n-loop+min"]


def _need_profile_lib():
    if not hasattr(lib(), "intopt_debug_phase_cycles"):
        raise RuntimeError("profiling entry points live in libintopt_b200_profile.so: build it with "
                           "`bash neuripsintopt_b200/csrc/build.sh profile` and set INTOPT_LIB_OVERRIDE=%s" % PROFILE_LIB_PATH)


def phase_cycles(enable=True, read=False):
    """Debug aid: toggle per-phase cycle accounting of the forward kernel; optionally read + reset the buckets."""
    _need_profile_lib()
    import numpy as np
    out = np.zeros(32, dtype=np.uint64)
    check(lib().intopt_debug_phase_cycles(1 if enable else 0, out.ctypes.data if read else None),
          "intopt_debug_phase_cycles")
    return dict(zip(PHASES, out[:len(PHASES)].tolist())) if read else None


MICRO = ["barrier", "A.v (spmv_rows)", "A^T.y", "A^T.[y1 y2]", "assembly", "assembly+cholesky", "solve 2 rhs",
         "solve 1 rhs", "n-loop smem<-2 global", "cta_sum<5>", "n-loop global rmw", "256 dependent DFMA", "256 dependent LDS",
         "256 dependent SHFL64+DADD", "256 dependent DADD", "64 dependent L2 loads", "64 dependent rsqrt+add",
         "64 dependent div+add", "spmv: headers only", "spmv: +run_total", "spmv: +scan only", "spmv: full copy", "10 barriers",
         "9x(warp0 LDS+16 DFMA+STS, barrier)+barrier", "chol w/o trailing updates", "chol w/o diag factor",
         "chol w/o panels", "chol w/o fuse", "chol skeleton only", "solve2 w/o pivots", "solve2 w/o row updates",
         "solve2 skeleton only", "2*nsteps x (record load, barrier)", "2*nsteps x (record load, branches, barrier)"]


MICRO += ["-"] * (35 - len(MICRO))
MICRO[34] = "assembly: scan + zero fill only"


def microbench(handle, which, reps=20, grid=1):
    """Debug aid: median/max SM cycles per repetition of one building block on `grid` persistent CTAs."""
    import numpy as np
    _need_profile_lib()
    out = np.zeros(grid, dtype=np.uint64)
    check(lib().intopt_debug_microbench(handle, which, reps, grid, out.ctypes.data), "intopt_debug_microbench")
    return float(np.median(out)), float(out.max())
