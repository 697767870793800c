// C ABI (include/intopt_b200.h) over the sm_100a solver kernels.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/intopt_b200.h"
#include "intopt_solver.cuh"
#include "intopt_structure.hpp"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU_TRY(expr)                                                                             \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(-100 - (int)e_, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
  } while (0)

constexpr int kThreads = 256;
constexpr int kMaxDynSmem = 232448;  // 227 KB opt-in per CTA on sm_100

size_t smem_vec_bytes(int m) { return (size_t)(intopt::S_NVEC * m + 1 + 8 * 32) * sizeof(double); }

void fill_info(const intopt::HostStructure& S, IntoptInfo* info) {
  info->m = S.m;
  info->n = S.n;
  info->nnz = S.nnz;
  info->envelope = S.sky_size;
  info->dense_lower = (int)((long long)S.m * (S.m + 1) / 2);
  info->asm_terms = S.asm_terms;
  info->chol_macs = S.chol_macs;
  size_t with_factor = smem_vec_bytes(S.m) + (size_t)S.sky_size * sizeof(double);
  info->factor_in_smem = with_factor <= (size_t)kMaxDynSmem ? 1 : 0;
  info->smem_bytes = (int)(info->factor_in_smem ? with_factor : smem_vec_bytes(S.m));
  info->threads = kThreads;
  info->ordering = S.ordering;
}

template <class T>
cudaError_t upload(const std::vector<T>& h, T** d) {
  *d = nullptr;
  if (h.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc((void**)d, h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

struct IntoptProblem {
  intopt::ProblemDev dev{};
  IntoptInfo info{};
  int device = 0;
  std::vector<void*> allocs;
};

__global__ void __launch_bounds__(kThreads) intopt_forward_kernel(intopt::ProblemDev P, intopt::FwdParams a) {
  extern __shared__ double smem[];
  CtaCtx cx;
  cx.tid = threadIdx.x;
  cx.nthreads = blockDim.x;
  cx.lane = threadIdx.x & 31;
  cx.warp = threadIdx.x >> 5;
  cx.nwarps = blockDim.x >> 5;
  cx.red = smem + (a.factor_in_smem ? P.sky_size : 0) + intopt::S_NVEC * P.m + 1;
  for (int inst = blockIdx.x; inst < a.B; inst += gridDim.x) {
    intopt::forward_instance(cx, P, a, inst, smem);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) intopt_backward_kernel(intopt::ProblemDev P, intopt::BwdParams a) {
  extern __shared__ double smem[];
  CtaCtx cx;
  cx.tid = threadIdx.x;
  cx.nthreads = blockDim.x;
  cx.lane = threadIdx.x & 31;
  cx.warp = threadIdx.x >> 5;
  cx.nwarps = blockDim.x >> 5;
  cx.red = smem + (a.factor_in_smem ? P.sky_size : 0) + intopt::S_NVEC * P.m + 1;
  for (int inst = blockIdx.x; inst < a.B; inst += gridDim.x) {
    intopt::backward_instance(cx, P, a, inst, smem);
    __syncthreads();
  }
}

// Register-resident DFMA chains: the measured FP64 roofline denominator.
__global__ void __launch_bounds__(256) intopt_fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c);
      a1 = fma(a1, m, c);
      a2 = fma(a2, m, c);
      a3 = fma(a3, m, c);
      a4 = fma(a4, m, c);
      a5 = fma(a5, m, c);
      a6 = fma(a6, m, c);
      a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" {

int intopt_version(void) { return INTOPT_B200_VERSION; }
const char* intopt_last_error(void) { return g_err.c_str(); }

int intopt_analyze(int m, int n, const int* rowptr, const int* colidx, const double* val, IntoptInfo* info,
                   int* row_perm) {
  if (!info) return fail(-1, "info is NULL");
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rowptr, colidx, val, S);
  if (rc) return fail(rc, "malformed CSR input");
  fill_info(S, info);
  if (row_perm) std::memcpy(row_perm, S.perm.data(), sizeof(int) * m);
  return 0;
}

int intopt_problem_create(IntoptProblem** out, int m, int n, const int* rowptr, const int* colidx,
                          const double* val, const double* b, int flags) {
  (void)flags;
  if (!out) return fail(-1, "out is NULL");
  *out = nullptr;
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rowptr, colidx, val, S);
  if (rc) return fail(rc, "malformed CSR input");
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) return fail(-10, "no CUDA device: intopt_b200 has no CPU fallback");
  IntoptProblem* p = new (std::nothrow) IntoptProblem();
  if (!p) return fail(-2, "out of host memory");
  CU_TRY(cudaGetDevice(&p->device));
  fill_info(S, &p->info);
  intopt::ProblemDev& D = p->dev;
  D.m = m;
  D.n = n;
  D.nnz = S.nnz;
  D.sky_size = S.sky_size;
  D.n_asm = (int)S.asm_pos.size();
  D.asm_unit = S.asm_unit ? 1 : 0;
  std::vector<double> bperm;
  if (b) {
    bperm.resize(m);
    for (int i = 0; i < m; ++i) bperm[i] = b[S.perm[i]];
  }
  std::vector<double> wdummy;
  int *d_i = nullptr;
  double* d_d = nullptr;
  unsigned char* d_u = nullptr;
#define UP_I(vec, field)                    \
  CU_TRY(upload(vec, &d_i));                \
  D.field = d_i;                            \
  p->allocs.push_back(d_i);
#define UP_D(vec, field)                    \
  CU_TRY(upload(vec, &d_d));                \
  D.field = d_d;                            \
  p->allocs.push_back(d_d);
  UP_I(S.csr_ptr, csr_ptr)
  UP_I(S.csr_col, csr_col)
  UP_D(S.csr_val, csr_val)
  UP_I(S.csc_ptr, csc_ptr)
  UP_I(S.csc_row, csc_row)
  UP_D(S.csc_val, csc_val)
  UP_D(bperm, b)
  UP_I(S.perm, perm)
  UP_I(S.sky_first, sky_first)
  UP_I(S.sky_off, sky_off)
  UP_I(S.asm_ptr, asm_ptr)
  UP_I(S.asm_pos, asm_pos)
  UP_I(S.asm_k, asm_k)
  if (!S.asm_unit) {
    UP_D(S.asm_w, asm_w)
  } else {
    D.asm_w = nullptr;
  }
  CU_TRY(upload(S.asm_diag, &d_u));
  D.asm_diag = d_u;
  p->allocs.push_back(d_u);
#undef UP_I
#undef UP_D
  CU_TRY(cudaFuncSetAttribute(intopt_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  CU_TRY(cudaFuncSetAttribute(intopt_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  CU_TRY(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

int intopt_problem_destroy(IntoptProblem* p) {
  if (!p) return 0;
  for (void* a : p->allocs)
    if (a) cudaFree(a);
  delete p;
  return 0;
}

int intopt_problem_info(const IntoptProblem* p, IntoptInfo* info) {
  if (!p || !info) return fail(-1, "NULL argument");
  *info = p->info;
  return 0;
}

size_t intopt_workspace_bytes(const IntoptProblem* p, int B) {
  if (!p || B <= 0) return 0;
  size_t per = (size_t)intopt::W_NVEC * p->info.n + (p->info.factor_in_smem ? 0 : (size_t)p->info.envelope);
  return per * sizeof(double) * (size_t)B;
}

int intopt_forward(const IntoptProblem* p, int B, const double* c, long long c_stride, const double* b_batch,
                   long long b_stride, const IntoptOpts* o, const double* init, double* xhat, double* state, int* status, int* iters,
                   void* ws, size_t ws_bytes, void* stream) {
  if (!p || !c || !o || !xhat || !state || !status || !iters || !ws) return fail(-1, "NULL argument");
  if (B <= 0) return 0;
  if (!p->dev.b && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
  if (ws_bytes < intopt_workspace_bytes(p, B)) return fail(-4, "workspace too small");
  intopt::FwdParams a{};
  a.B = B;
  a.c = c;
  a.c_stride = c_stride;
  a.b_batch = b_batch;
  a.b_stride = b_stride;
  a.init = init;
  a.xhat = xhat;
  a.state = state;
  a.status = status;
  a.iters = iters;
  a.ws = (double*)ws;
  a.factor_in_smem = p->info.factor_in_smem;
  a.ws_stride = (long long)intopt::W_NVEC * p->info.n + (a.factor_in_smem ? 0 : p->info.envelope);
  a.alpha0 = o->alpha0;
  a.tol = o->tol;
  a.thr = o->thr;
  a.damping = o->damping_fwd;
  a.max_iter = o->max_iter;
  intopt_forward_kernel<<<B, kThreads, p->info.smem_bytes, (cudaStream_t)stream>>>(p->dev, a);
  CU_TRY(cudaGetLastError());
  return 0;
}

int intopt_backward(const IntoptProblem* p, int B, int mode, double damping, int nvar, const double* state,
                    const double* c, long long c_stride, const double* b_batch, long long b_stride,
                    const double* gx, double* gc,
                    int* status, void* ws, size_t ws_bytes, void* stream) {
  if (!p || !c || !state || !gx || !gc || !status || !ws) return fail(-1, "NULL argument");
  if (B <= 0) return 0;
  if (mode != 0 && mode != 1) return fail(-5, "mode must be 0 (whole) or 1 (alter)");
  if (nvar < 0 || nvar > p->info.n) return fail(-6, "nvar out of range");
  if (mode == 0 && !p->dev.b && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
  if (ws_bytes < intopt_workspace_bytes(p, B)) return fail(-4, "workspace too small");
  intopt::BwdParams a{};
  a.B = B;
  a.mode = mode;
  a.nvar = nvar;
  a.damping = damping;
  a.state = state;
  a.c = c;
  a.c_stride = c_stride;
  a.b_batch = b_batch;
  a.b_stride = b_stride;
  a.gx = gx;
  a.gc = gc;
  a.status = status;
  a.ws = (double*)ws;
  a.factor_in_smem = p->info.factor_in_smem;
  a.ws_stride = (long long)intopt::W_NVEC * p->info.n + (a.factor_in_smem ? 0 : p->info.envelope);
  intopt_backward_kernel<<<B, kThreads, p->info.smem_bytes, (cudaStream_t)stream>>>(p->dev, a);
  CU_TRY(cudaGetLastError());
  return 0;
}

int intopt_measure_fp64_peak(double* gflops_out, int repeats) {
  if (!gflops_out) return fail(-1, "NULL argument");
  int dev = 0, sms = 0;
  CU_TRY(cudaGetDevice(&dev));
  CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  CU_TRY(cudaMalloc((void**)&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < (repeats > 0 ? repeats : 5) + 2; ++r) {
    CU_TRY(cudaEventRecord(e0));
    intopt_fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + r);
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    double gf = flops / (ms * 1e-3) * 1e-9;
    if (r >= 2 && gf > best) best = gf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *gflops_out = best;
  return 0;
}

}  // extern "C"
