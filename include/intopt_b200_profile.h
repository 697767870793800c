/*
 * intopt_b200 — profiling aids.  NOT part of the product library: these entry points exist only in
 * libintopt_b200_profile.so (neuripsintopt_b200/csrc/build.sh profile, every translation unit compiled with
 * -DINTOPT_PROFILE, which also switches the per-phase cycle probes of the solver kernels on).  The diagnostics
 * under tests/gpu_phases*.py and tests/gpu_micro.py load that library through INTOPT_LIB_OVERRIDE.
 */
#ifndef INTOPT_B200_PROFILE_H
#define INTOPT_B200_PROFILE_H

#include "intopt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Debug aid for profiles/: enable (1) / disable (0) per-phase SM-cycle accounting of the forward kernel;
 * if out32_host is non-NULL the 32 cycle buckets are copied out (after a device sync) and reset.
 * Buckets: 0 setup, 1 residuals+tests, 2 assembly, 3 Cholesky, 4 rhs/spmv, 5 solve(2 rhs), 6 p/u loop,
 * 7 corrector rhs, 8 solve(1 rhs), 9 direction/ratio/update loops, 10 epilogue, 11-14 Cholesky sub-phases,
 * 16-23 sub-phases of 1, 4, 7 and 9 (names in neuripsintopt_b200/_capi.py:PHASES). */
int intopt_debug_phase_cycles(int enable, unsigned long long* out32_host);

/* Debug aid for profiles/: SM cycles per repetition of one building block of the solver (which: 0 barrier only,
 * 1 A.v, 2 A^T.y, 3 A^T.[y1 y2], 4 assembly, 5 assembly+Cholesky, 6 solve (2 rhs), 7 solve (1 rhs), 8/10 n-loops,
 * 9 CTA reduction) on `grid` persistent CTAs; cycles_host[grid].  Synchronises the device. */
int intopt_debug_microbench(const IntoptProblem* p, int which, int reps, int grid, unsigned long long* cycles_host);

#ifdef __cplusplus
}
#endif
#endif /* INTOPT_B200_PROFILE_H */
