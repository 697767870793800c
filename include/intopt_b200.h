/*
 * intopt_b200 — C ABI of the B200-native differentiable interior-point LP layer.
 *
 * This is the drop-in boundary for the hot path of JayMan91/NeurIPSIntopt:
 *   forward  = IPOfunc_cls.forward's solver loop, Interior/ip_model_whole.py:1385-1447
 *              (_initialization :647, _get_delta :761, _get_step :870, _do_step :861,
 *              _indicators :880, termination tests :1408-1423)
 *   backward = IPOfunc_cls.backward, Interior/ip_model_whole.py:1502-1560 (mode 0, "whole") and
 *              Interior/ip_model_alter.py:1516-1624 (mode 1, "alter")
 * operating on the standard form  min c.x  s.t.  A x = b, x >= 0  that the reference builds on
 * the host (_clean_inputs/_presolve/_get_Abc, :67,:918,:333) — that part stays on the host
 * (neuripsintopt_b200/problem.py).
 *
 * The reference is pure Python and has no FFI; the Python binding a maintainer would add is
 * the ctypes stub shown in INTEGRATION.md (neuripsintopt_b200/_capi.py is that stub).
 *
 * Conventions: plain C types only.  Pointers named *_dev are device pointers owned by the
 * caller (PyTorch allocations) and only borrowed for the call; *_host are host pointers.
 * All work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 * the device except intopt_problem_create/destroy.  Return value: 0 = ok, <0 = error
 * (see intopt_last_error()).  Per-instance outcomes go to status_dev[B]:
 *   0 ok/stopped by the go test, 1 max_iter reached (:1414), 2 infeasibility/unboundedness
 *   indicators fired (:1409-1413), 3 factorisation breakdown (:855-857,:1419), 4 non-finite c.
 * There is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef INTOPT_B200_H
#define INTOPT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INTOPT_B200_VERSION 5

/* intopt_problem_create flag: the VALUES of A's non-zeros differ per instance (same sparsity pattern), e.g. the
 * knapsack driver whose equality row is the minibatch's cost vector (experiments/Building Knapsack/optimizer.py:396-405).
 * val_host then only fixes the pattern (pass any non-zero numbers); every forward/backward call must go through the
 * *_values entry points with a_vals_dev. */
#define INTOPT_FLAG_INSTANCE_VALUES 1

typedef struct IntoptProblem IntoptProblem; /* opaque, immutable after create, one per device */

/* Solver options = the IPOfunc keyword arguments that reach the loop (ip_model_whole.py:1287-1290). */
typedef struct IntoptOpts {
  double alpha0;      /* step damping, default 0.9995 (:1287, used at :1402) */
  double tol;         /* termination tolerance, default 1e-8 (:1288, :1408-1410) */
  double thr;         /* early stop: loop ends when mu <= thr (:1408); thr=None -> 0 (:1319) */
  double damping_fwd; /* diagonal damping of A.diag(x/t).A^T: 1e-6 in whole (:774), `damping` in alter */
  int max_iter;       /* up to max_iter+1 iterations run (:1414) */
  int plain;          /* 0: Mehrotra predictor-corrector (pc=True, every reference driver); 1: pc=False — one direction
                         per iteration with centring gamma = beta * mean(x.t) (:1392-1394, :846-847) */
  double beta;        /* only read when plain = 1; default 0.1 (:1287) */
} IntoptOpts;

/* Static facts about a problem after structure analysis (host side, no GPU needed to read). */
typedef struct IntoptInfo {
  int m, n, nnz;
  int envelope;           /* doubles in the 8x8 block-skyline storage of chol(A D A^T) after row ordering */
  int dense_lower;        /* m(m+1)/2, for comparison */
  long long asm_terms;    /* S = sum over columns of nnz_k(nnz_k+1)/2: MACs to assemble A D A^T */
  long long chol_macs;    /* MACs executed by the block-skyline Cholesky */
  int smem_bytes;         /* dynamic shared memory the solver kernel requests per CTA */
  int factor_in_smem;     /* 1 if the factor lives in shared memory, 0 if it spills to the workspace */
  int threads;            /* CTA size (env INTOPT_THREADS, default 256) */
  int ordering;           /* 0 natural, 1 reverse Cuthill-McKee, 2 dense rows last */
  int gather_in_smem;     /* 1 if the n-vector gather buffer lives in shared memory */
  int ctas_per_sm;        /* resident CTAs per SM (set by intopt_problem_create, 0 from intopt_analyze) */
  int scalar_envelope;    /* doubles a scalar skyline would need (for comparison) */
  int path;               /* 0: persistent-CTA path (one CTA owns one LP, factor in shared memory);
                             1: lock-step dense path (batch-wide DMMA kernels, factor tiled through HBM) */
  long long asm_padded;   /* gather-list slots incl. padding of the warp-interleaved layout */
  long long solve_macs;   /* MACs of one blocked L L^T solve */
  int tmem;               /* 1 if the kernels keep their element-wise n-vectors in tensor memory (tcgen05.ld/st) instead of
                             the L2-resident workspace: n <= 7 * threads (and n >= threads on 128-thread CTAs); CTAs of 128 (four per SM), 256 (two) or 512 threads (one); env INTOPT_TMEM != 0 (set by
                             intopt_problem_create, 0 from intopt_analyze) */
} IntoptInfo;

int intopt_version(void);
const char* intopt_last_error(void);

/* Host-only structure analysis of a standard-form A (CSR, m x n).  Fills `info` and, if non-NULL,
 * row_perm_host[m] (internal row i = original row row_perm_host[i]).  Needs no GPU.
 * Returns 0, -1..-4 for malformed input (NULL / non-monotone rowptr / column out of range / m too large for the
 * 16-bit row indices), -5 when a SPARSE problem is too large for the gather-list assembly (asm_terms > 2^28);
 * intopt_problem_create returns the same codes.  Dense problems (nnz >= m n / 4) with m >= 128 take the lock-step dense
 * path (info.path = 1, m <= 4096) and are not subject to that limit. */
int intopt_analyze(int m, int n, const int* rowptr_host, const int* colidx_host, const double* val_host,
                   IntoptInfo* info, int* row_perm_host);

/* Build the device-resident problem on the current CUDA device: CSR/CSC copies of A, row ordering,
 * envelope layout and the gather lists that assemble A.diag(d).A^T.
 * b_host may be NULL when every call passes b_batch_dev (per-instance right-hand sides,
 * shortespath/shortespath.py:407-413).  flags: reserved, pass 0. */
int intopt_problem_create(IntoptProblem** out, int m, int n, const int* rowptr_host, const int* colidx_host,
                          const double* val_host, const double* b_host, int flags);
int intopt_problem_destroy(IntoptProblem* p);
int intopt_problem_info(const IntoptProblem* p, IntoptInfo* info);

/* Bytes of device workspace a forward/backward call over B instances needs. */
size_t intopt_workspace_bytes(const IntoptProblem* p, int B);

/* Forward: B independent LPs sharing A.  Replaces the loop at ip_model_whole.py:1385-1447.
 *   c_dev        B x n standard-form costs, instance stride c_stride doubles (0 = one shared c)
 *   b_batch_dev  B x m right-hand sides, instance stride b_stride doubles (0 = one shared b), or NULL
 *                (use the problem's b)
 *   init_dev     optional warm start, B x (2n+m+2) in the `state` layout (init_val, :655-660), or NULL
 *   xhat_dev     out, B x n:  x / tau  (:1441)
 *   state_dev    out, B x (2n+m+2): x[n], y[m], t[n], tau, kappa — what forward saves for backward (:1463)
 *   status_dev, iters_dev  out, int32[B]; status 0 ok, 1 max_iter reached, 2 infeasible / unbounded by the indicators
 *                (:1409-1413), 3 factorisation breakdown that two retries with a raised diagonal did not cure (the
 *                ladder of :798-822; the iterate is left unchanged, :855-857), 4 non-finite cost vector
 */
int intopt_forward(const IntoptProblem* p, int B, const double* c_dev, long long c_stride,
                   const double* b_batch_dev, long long b_stride, const IntoptOpts* opts,
                   const double* init_dev,
                   double* xhat_dev, double* state_dev, int* status_dev, int* iters_dev,
                   void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same with per-instance values of A: a_vals_dev is B x nnz (instance stride a_stride doubles), the values of A's
 * non-zeros in the order of the CSR arrays given to intopt_problem_create.  NULL = the problem's own values. */
int intopt_forward_values(const IntoptProblem* p, int B, const double* c_dev, long long c_stride,
                          const double* b_batch_dev, long long b_stride, const double* a_vals_dev, long long a_stride,
                          const IntoptOpts* opts, const double* init_dev,
                          double* xhat_dev, double* state_dev, int* status_dev, int* iters_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* Backward: dL/dc at the saved iterate.  mode 0 = ip_model_whole.py:1502-1560 (HSD embedding, uses the
 * un-normalised iterate), mode 1 = ip_model_alter.py:1516-1624 (log barrier, uses x/tau, t/tau).
 *   grad_x_dev   B x nvar  upstream gradient w.r.t. the leading nvar standard-form variables
 *   grad_c_dev   out, B x nvar  (:1555-1556: only J[:nvar,:nvar] is used)
 *   status_dev   out, int32[B]: 0 ok, 3 factorisation breakdown
 */
int intopt_backward(const IntoptProblem* p, int B, int mode, double damping_bwd, int nvar,
                    const double* state_dev, const double* c_dev, long long c_stride,
                    const double* b_batch_dev, long long b_stride, const double* grad_x_dev,
                    double* grad_c_dev,
                    int* status_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

int intopt_backward_values(const IntoptProblem* p, int B, int mode, double damping_bwd, int nvar,
                           const double* state_dev, const double* c_dev, long long c_stride,
                           const double* b_batch_dev, long long b_stride, const double* a_vals_dev, long long a_stride,
                           const double* grad_x_dev, double* grad_c_dev,
                           int* status_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Optional mapping between the caller's own layout / precision and the standard form, applied INSIDE the solver kernels
 * (no separate pack / unpack launches).  Any of the two halves may be off (in_dev / out_dev NULL).
 *   input : standard-form entry j = in_scale[j] * in[inst][in_idx[j]]  (in_idx[j] < 0: 0; in_scale NULL: 1).  Forward: the
 *           user's cost vectors (length of in_idx = n); backward: the upstream gradient (length nvar).  The packed rows are
 *           written to c_dev / grad_x_dev, which are then outputs (the packed c is what the backward call wants as c_dev).
 *   output: user entry i (< n_out) = std[out_idx[i]], or out_const[i] where out_idx[i] < 0 (out_const NULL: 0).  Forward:
 *           x with the presolve-fixed variables re-inserted (_postprocess, ip_model_whole.py:1241-1285; xhat_dev may then
 *           be NULL); backward: c_grad[var_index] = dc, zeros elsewhere (:1559-1560); grad_c_dev is then NOT written (it must still
 *           be a valid pointer). */
typedef struct IntoptUserMap {
  const void* in_dev;
  int in_is_f32;
  long long in_stride;      /* elements between instances, 0 = one shared row */
  const int* in_idx_dev;
  const double* in_scale_dev;
  void* out_dev;
  int out_is_f32;
  long long out_stride;
  int n_out;
  const int* out_idx_dev;
  const double* out_const_dev;
} IntoptUserMap;

int intopt_forward_mapped(const IntoptProblem* p, int B, double* c_dev, long long c_stride,
                          const double* b_batch_dev, long long b_stride, const double* a_vals_dev, long long a_stride,
                          const IntoptOpts* opts, const double* init_dev, const IntoptUserMap* map,
                          double* xhat_dev, double* state_dev, int* status_dev, int* iters_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream);
int intopt_backward_mapped(const IntoptProblem* p, int B, int mode, double damping_bwd, int nvar,
                           const double* state_dev, const double* c_dev, long long c_stride,
                           const double* b_batch_dev, long long b_stride, const double* a_vals_dev, long long a_stride,
                           const IntoptUserMap* map, double* grad_x_dev, double* grad_c_dev,
                           int* status_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Host-glue helper (one launch on `stream`): dst[b][j] = idx[j] >= 0 ? scale[j] * src[b][idx[j]] : constv[j] for
 * b < B, j < n_out, with float <-> double conversion (src_is_f32 / dst_is_f32; idx NULL = identity, scale NULL = 1,
 * constv NULL = 0; strides in elements, src_stride 0 = one shared row).  Maps the user's cost vectors to the standard
 * form (ip_model_whole.py:1380 _get_Abc's c), the solution back (_postprocess :1241-1285, fixed variables
 * re-inserted) and the gradient both ways (:1518, :1559-1560) without a chain of cast / fill / index kernels. */
int intopt_gather_convert(const void* src_dev, int src_is_f32, long long src_stride, int B, int n_out,
                          const int* idx_dev, const double* scale_dev, const double* constv_dev, void* dst_dev,
                          int dst_is_f32, long long dst_stride, void* stream);

/* Lock-step dense path only (IntoptInfo.path == 1): z_b = (A diag(d_b) A^T + delta I)^-1 rhs_b for B instances — the
 * normal-matrix assembly (:772-775), tiled Cholesky (:742) and blocked substitution (:745) kernels as a stand-alone
 * primitive.  d_dev: B x n, rhs_dev: B x m (in: right-hand sides, out: solutions), fail_dev: int32[B] (non-positive
 * pivot) or NULL; workspace as for intopt_forward. */
int intopt_dense_normal_solve(const IntoptProblem* p, int B, const double* d_dev, double delta, double* rhs_dev,
                              int* fail_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Measured FP64 tensor-core (DMMA m8n8k4) throughput of the current device (GFLOP/s), register-resident chains:
 * the roofline denominator of the dense path.  Synchronises the device. */
int intopt_measure_dmma_peak(double* gflops_out, int repeats);

/* Measured FP64 FMA throughput of the current device (GFLOP/s) with a register-resident DFMA kernel;
 * the roofline denominator MEASURED_PEAKS.json lacks.  Synchronises the device. */
int intopt_measure_fp64_peak(double* gflops_out, int repeats);

#ifdef __cplusplus
}
#endif
#endif /* INTOPT_B200_H */
