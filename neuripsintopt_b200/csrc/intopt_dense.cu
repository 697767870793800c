// Lock-step dense path: host side (problem data, workspace layout, the iteration schedule).  Device code and the
// description of the design: intopt_dense.cuh.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda.h>  // CUtensorMap and its enums only: cuTensorMapEncodeTiled is resolved through the runtime (no -lcuda)

#include "intopt_dense.cuh"
#include "intopt_internal.hpp"

struct IntoptDense {
  int m = 0, n = 0, m64 = 0, n64 = 0, sms = 0;
  double* A = nullptr;       // m64 x n64, zero padded
  double* At = nullptr;      // n64 x m64
  double* rowsum = nullptr;  // A 1
  double rowsq_max = 0.0;    // max_i sum_k A_ik^2 (bound of the diagonal of A diag(d) A^T per unit of max d)
  double* b = nullptr;       // m or null
  // TMA descriptors of the padded A: box = 16 columns (128 bytes) x 128 / 64 rows, 128-byte swizzle
  dn::TmaDesc tmap128{}, tmap64{};
  bool use_tma = false;
  bool use_ladder = true;  // INTOPT_DENSE_LADDER=0: a breakdown ends the instance at once (A/B measurements)
  std::vector<void*> allocs;
  // completion signalling of the iteration loop (one forward call at a time per handle)
  mutable std::mutex mu;
  int* h_nactive = nullptr;  // pinned
  cudaEvent_t ev = nullptr;
  size_t ws_budget = 0;
  // Forward as ONE CUDA graph: init kernel -> while-node whose body is an interior-point iteration -> epilogue.  The
  // condition is set on the device (dense_loop_cond_kernel), so a forward call enqueues one graph launch and never
  // waits on the host.  Graphs are cached per (workspace, chunk size, damping, pc, right-hand-side pointer): everything
  // else a call varies travels through the argument block in device memory.
  struct FwdGraph {
    const void* ws;
    int Bc, plain;
    double damping;
    const double* b_batch;
    long long b_stride;
    cudaGraphExec_t exec;
  };
  mutable std::vector<FwdGraph> graphs;
  cudaStream_t cap_stream = nullptr;
  bool use_graph = true;
};

namespace {

#define DN_TRY(expr)                                                                          \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      if (err) *err = std::string(#expr) + ": " + cudaGetErrorString(e_);                     \
      return -100 - (int)e_;                                                                  \
    }                                                                                         \
  } while (0)

constexpr size_t kArgsOffset = 256;
static_assert(sizeof(dn::DenseFwdArgs) <= 1024, "argument block of the forward graph");
int up64(int v) { return (v + 63) & ~63; }
int up128(int v) { return (v + 127) & ~127; }

// Workspace of a chunk of Bc instances (doubles unless noted).
struct Layout {
  int Bc, Bp;
  size_t X, T, Y, AX, ATY, R2, D, W, P, DX, DXDT, G, RHS, ATQV, SC, M, Linv, ints, dretry, total;  // offsets in bytes
};
size_t per_instance_bytes(int m64, int n64) {
  return ((size_t)12 * n64 + 4 * (size_t)n64 + 5 * (size_t)m64 + dn::SC_N + (size_t)m64 * m64 + (size_t)m64 * dn::NB) * 8 + 16;
}
Layout make_layout(const IntoptDense* d, int Bc) {
  Layout L{};
  L.Bc = Bc;
  L.Bp = up128(Bc);
  const size_t nv = (size_t)L.Bp * d->n64 * 8, mv = (size_t)L.Bp * d->m64 * 8;
  size_t o = kArgsOffset + 1024;  // [0,256): n_active counter; [256, 256 + 1 KB): the argument block of the forward graph
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += (bytes + 255) & ~(size_t)255;
    return at;
  };
  L.X = take(nv);
  L.T = take(nv);
  L.Y = take(mv);
  L.AX = take(mv);
  L.ATY = take(nv);
  L.R2 = take(nv);
  L.D = take(nv);
  L.W = take(nv);
  L.P = take(nv);
  L.DX = take(nv);
  L.DXDT = take(nv);
  L.G = take(2 * nv);
  L.RHS = take(2 * mv);
  L.ATQV = take(2 * nv);
  L.SC = take((size_t)L.Bp * dn::SC_N * 8);
  L.M = take((size_t)Bc * d->m64 * d->m64 * 8);
  L.Linv = take((size_t)Bc * d->m64 * dn::NB * 8);
  L.ints = take((size_t)L.Bp * 3 * sizeof(int));
  L.dretry = take((size_t)L.Bp * 8);
  L.total = o;
  return L;
}
int chunk_size(const IntoptDense* d, int B) {
  const size_t per = per_instance_bytes(d->m64, d->n64) + 2 * 128 * 20 * (size_t)d->n64 / 128;  // (+ row padding slack)
  const size_t cap = std::max<size_t>(1, d->ws_budget / per);
  return (int)std::min<size_t>((size_t)B, cap);
}
dn::DenseState bind_state(const IntoptDense* d, const Layout& L, char* ws) {
  dn::DenseState s{};
  s.B = L.Bc;
  s.Bp = L.Bp;
  s.m = d->m;
  s.n = d->n;
  s.m64 = d->m64;
  s.n64 = d->n64;
  auto at = [&](size_t off) { return reinterpret_cast<double*>(ws + off); };
  s.X = at(L.X);
  s.T = at(L.T);
  s.Y = at(L.Y);
  s.AX = at(L.AX);
  s.ATY = at(L.ATY);
  s.R2 = at(L.R2);
  s.D = at(L.D);
  s.W = at(L.W);
  s.P = at(L.P);
  s.DX = at(L.DX);
  s.DXDT = at(L.DXDT);
  s.G = at(L.G);
  s.RHS = at(L.RHS);
  s.ATQV = at(L.ATQV);
  s.SC = at(L.SC);
  s.M = at(L.M);
  s.Linv = at(L.Linv);
  s.active = reinterpret_cast<int*>(ws + L.ints);
  s.fail = s.active + L.Bp;
  s.retry = s.fail + L.Bp;
  s.delta_retry = at(L.dretry);
  s.n_active = reinterpret_cast<int*>(ws);
  s.rowsum = d->rowsum;
  return s;
}

constexpr int kGemmSmem = dn::GemmProd::SMEM_DOUBLES * 8;
constexpr int kSyrk128Smem = dn::SyrkCfg<128, 512>::G::SMEM_DOUBLES * 8;
constexpr int kSyrk64Smem = dn::SyrkCfg<64, 128>::G::SMEM_DOUBLES * 8;
constexpr int kCholSmem = dn::CHOL_SMEM_DOUBLES * 8;
constexpr int kSyrkTma128Smem = dn::SyrkTma<128, 512>::SMEM_BYTES;
constexpr int kSyrkTma64Smem = dn::SyrkTma<64, 128>::SMEM_BYTES;

// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: tiles of `box_rows` rows x 16 doubles of a row-major
// (rows x cols) fp64 matrix, 128-byte swizzle.  Returns false if the driver entry point is unavailable.
bool make_tile_map(dn::TmaDesc* out, const double* base, int rows, int cols, int box_rows) {
  static_assert(sizeof(CUtensorMap) == sizeof(dn::TmaDesc), "CUtensorMap is 128 bytes");
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 8};
  const cuuint32_t box[2] = {(cuuint32_t)dn::KT, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = ((EncodeFn)fn)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim,
                                    gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
int solve_smem(int nr, int m64) { return (nr * m64 + nr * dn::NB + 16 * nr * dn::NB) * 8; }

std::once_flag g_attr_once;
cudaError_t g_attr_err = cudaSuccess;
void set_attrs() {
  auto set = [&](const void* k, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess && g_attr_err == cudaSuccess) g_attr_err = e;
  };
  set((const void*)dn::dense_gemm_kernel, kGemmSmem);
  set((const void*)dn::dense_syrk_kernel<128, 512>, kSyrk128Smem);
  set((const void*)dn::dense_syrk_kernel<64, 128>, kSyrk64Smem);
  set((const void*)dn::dense_syrk_tma_kernel<128, 512>, kSyrkTma128Smem);
  set((const void*)dn::dense_syrk_tma_kernel<64, 128>, kSyrkTma64Smem);
  set((const void*)dn::dense_chol_col_kernel, kCholSmem);
  set((const void*)dn::dense_chol_trsm_kernel, kCholSmem);
  set((const void*)dn::dense_solve_kernel<1>, solve_smem(1, 4096));
  set((const void*)dn::dense_solve_kernel<2>, solve_smem(2, 4096));
}

// Out(rows x N) = In(rows x K) W(N x K)^T
void gemm(const double* In, long long ld_in, int rows, const double* W, long long ld_w, int N, int K, double* Out,
          long long ld_out, cudaStream_t st) {
  dim3 grid(N / 64, rows / 128);
  dn::dense_gemm_kernel<<<grid, 256, kGemmSmem, st>>>(In, ld_in, W, ld_w, K, Out, ld_out);
}

// M_b = A diag(D_b) A^T + delta I for the instances selected by `mask`, then its tiled Cholesky factorisation (in place)
// with the inverses of the diagonal blocks.  `delta_arr` (optional): per-instance damping.
void assemble_and_factor_masked(const IntoptDense* d, const dn::DenseState& s, int Bc, double delta, const int* mask,
                                const double* delta_arr, cudaStream_t st) {
  static const int tile_env = [] {
    const char* v = std::getenv("INTOPT_DENSE_TILE");
    return v && *v ? std::atoi(v) : 0;
  }();
  // 64-wide tiles by default: four 128-thread CTAs per SM hide the TMA / DMMA latencies better than one 512-thread CTA
  // with 128-wide tiles and waste less on diagonal tiles (measured: m = 256 0.38 -> 0.45 of the DMMA peak, m = 1024
  // 0.63 -> 0.70, m = 2048 0.70 -> 0.76); INTOPT_DENSE_TILE=128 selects the wide variant.
  if (d->use_tma && d->m64 % 128 == 0 && tile_env == 128) {
    const int T = d->m64 / 128;
    dn::dense_syrk_tma_kernel<128, 512><<<dim3(T * (T + 1) / 2, Bc), 512, kSyrkTma128Smem, st>>>(
        d->tmap128, d->n64, d->m, d->m64, s.D, d->n64, delta, s.M, mask, delta_arr);
  } else if (d->use_tma) {
    const int T = d->m64 / 64;
    dn::dense_syrk_tma_kernel<64, 128><<<dim3(T * (T + 1) / 2, Bc), 128, kSyrkTma64Smem, st>>>(
        d->tmap64, d->n64, d->m, d->m64, s.D, d->n64, delta, s.M, mask, delta_arr);
  } else if (d->m64 % 128 == 0) {
    const int T = d->m64 / 128;
    dn::dense_syrk_kernel<128, 512><<<dim3(T * (T + 1) / 2, Bc), 512, kSyrk128Smem, st>>>(d->A, d->n64, d->m, d->m64, s.D,
                                                                                      d->n64, delta, s.M, mask, delta_arr);
  } else {
    const int T = d->m64 / 64;
    dn::dense_syrk_kernel<64, 128><<<dim3(T * (T + 1) / 2, Bc), 128, kSyrk64Smem, st>>>(d->A, d->n64, d->m, d->m64, s.D,
                                                                                    d->n64, delta, s.M, mask, delta_arr);
  }
  const int nb = d->m64 / dn::NB;
  for (int k = 0; k < nb; ++k) {
    dn::dense_chol_col_kernel<<<dim3(nb - k, Bc), 256, kCholSmem, st>>>(s.M, d->m64, k, s.Linv, s.fail, mask);
    if (k + 1 < nb)
      dn::dense_chol_trsm_kernel<<<dim3(nb - k - 1, Bc), 256, kCholSmem, st>>>(s.M, d->m64, k, s.Linv, mask);
  }
}

// Assembly + factorisation of the active instances, then the two retry rungs of the failure ladder (:798-822) for the
// instances whose factorisation broke down (`ladder`: the forward; the backward reports a breakdown as the reference does).
void assemble_and_factor(const IntoptDense* d, const dn::DenseState& s, int Bc, double delta, cudaStream_t st,
                         bool ladder = false) {
  assemble_and_factor_masked(d, s, Bc, delta, s.active, nullptr, st);
  if (!ladder || !d->use_ladder) return;
  for (int rung = 0; rung < 2; ++rung) {
    dn::dense_retry_prep_kernel<<<Bc, 256, 0, st>>>(s, delta, d->rowsq_max, rung, nullptr);
    assemble_and_factor_masked(d, s, Bc, delta, s.retry, s.delta_retry, st);
  }
}

template <int NR>
void solve(const IntoptDense* d, const dn::DenseState& s, int Bc, dn::SolveArgs a, cudaStream_t st) {
  a.M = s.M;
  a.Linv = s.Linv;
  a.m = d->m;
  a.m64 = d->m64;
  a.ld = d->m64;
  a.active = s.active;
  a.skip = s.fail;
  dn::dense_solve_kernel<NR><<<Bc, 256, solve_smem(NR, d->m64), st>>>(a);
}

__global__ void pad_rows_kernel(const double* src, long long src_stride, int len, double* dst, long long dst_stride,
                                int dst_len, double fill) {
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < dst_len; k += blockDim.x)
    dst[(long long)b * dst_stride + k] = k < len ? src[(long long)b * src_stride + k] : fill;
}
__global__ void unpad_rows_kernel(const double* src, long long src_stride, int len, double* dst, long long dst_stride) {
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < len; k += blockDim.x) dst[(long long)b * dst_stride + k] = src[(long long)b * src_stride + k];
}
__global__ void set_flags_kernel(int* active, int* fail, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    active[i] = 1;
    fail[i] = 0;
  }
}

}  // namespace

namespace intopt_dense {

bool wanted(int m, int n, long long nnz, int flags) {
  const char* v = std::getenv("INTOPT_PATH");
  if (flags & INTOPT_FLAG_INSTANCE_VALUES) return false;  // per-instance values of A: persistent-CTA path only
  if (m > 4096 || n > 16384) return false;
  if (v && !std::strcmp(v, "dense")) return true;
  if (v && !std::strcmp(v, "cta")) return false;
  // dense problems whose factor is too large to share an SM's shared memory between resident LPs
  // (measured crossover against the persistent-CTA path: m = 128 wins at every batch size from 64 up - 341 k vs 196 k LP/s at
  // batch 1024 - m = 96 only from batch ~1000)
  return m >= 128 && (double)nnz >= 0.25 * (double)m * (double)n;
}

int create(IntoptDense** out, int m, int n, const int* rowptr, const int* colidx, const double* val, const double* b,
           std::string* err) {
  *out = nullptr;
  IntoptDense* d = new IntoptDense();
  d->m = m;
  d->n = n;
  d->m64 = up64(m);
  d->n64 = up64(n);
  std::vector<double> A((size_t)d->m64 * d->n64, 0.0), At((size_t)d->n64 * d->m64, 0.0), rs(m, 0.0);
  for (int i = 0; i < m; ++i) {
    double s = 0.0;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      A[(size_t)i * d->n64 + colidx[e]] = val[e];
      At[(size_t)colidx[e] * d->m64 + i] = val[e];
      s += val[e];
    }
    rs[i] = s;
    double sq = 0.0;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) sq += val[e] * val[e];
    d->rowsq_max = std::max(d->rowsq_max, sq);
  }
  auto upl = [&](const std::vector<double>& h, double** dev) -> cudaError_t {
    cudaError_t e = cudaMalloc((void**)dev, h.size() * sizeof(double));
    if (e != cudaSuccess) return e;
    d->allocs.push_back(*dev);
    return cudaMemcpy(*dev, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
  };
  cudaError_t e = upl(A, &d->A);
  if (e == cudaSuccess) e = upl(At, &d->At);
  if (e == cudaSuccess) e = upl(rs, &d->rowsum);
  if (e == cudaSuccess && b) {
    std::vector<double> bv(b, b + m);
    e = upl(bv, &d->b);
  }
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&d->h_nactive, 64, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking);
  {
    const char* v = std::getenv("INTOPT_DENSE_GRAPH");
    d->use_graph = !(v && v[0] == '0');
    const char* l = std::getenv("INTOPT_DENSE_LADDER");
    d->use_ladder = !(l && l[0] == '0');
  }
  int dev = 0;
  if (e == cudaSuccess) e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) {
    std::call_once(g_attr_once, set_attrs);
    e = g_attr_err;
  }
  if (e != cudaSuccess) {
    if (err) *err = std::string("dense path set-up: ") + cudaGetErrorString(e);
    destroy(d);
    return -100 - (int)e;
  }
  {  // TMA staging of the A tiles (INTOPT_DENSE_TMA=0: the cp.async variant)
    const char* v = std::getenv("INTOPT_DENSE_TMA");
    d->use_tma = !(v && v[0] == '0') && make_tile_map(&d->tmap128, d->A, d->m64, d->n64, 128) &&
                 make_tile_map(&d->tmap64, d->A, d->m64, d->n64, 64);
  }
  const char* gb = std::getenv("INTOPT_DENSE_WS_GB");
  d->ws_budget = (size_t)((gb && *gb ? std::atof(gb) : 32.0) * (double)(1ull << 30));
  *out = d;
  return 0;
}

void destroy(IntoptDense* d) {
  if (!d) return;
  for (void* a : d->allocs) cudaFree(a);
  if (d->h_nactive) cudaFreeHost(d->h_nactive);
  if (d->ev) cudaEventDestroy(d->ev);
  for (auto& g : d->graphs) cudaGraphExecDestroy(g.exec);
  if (d->cap_stream) cudaStreamDestroy(d->cap_stream);
  delete d;
}

size_t workspace_bytes(const IntoptDense* d, int B) {
  if (B <= 0) return 0;
  return make_layout(d, chunk_size(d, B)).total;
}

// One interior-point iteration of a chunk, enqueued on `st` (the host loop and the graph capture share it).
// `adev`: the argument block in device memory.
static void enqueue_after_factor(const IntoptDense* d, const dn::DenseState& s, const dn::DenseFwdArgs* adev, int Bc, int Bp,
                                 bool plain, const double* b_rhs, long long b_rhs_stride, cudaStream_t st);
static void enqueue_iteration(const IntoptDense* d, const dn::DenseState& s, const dn::DenseFwdArgs* adev, int Bc, int Bp,
                              double damping, bool plain, const double* b_rhs, long long b_rhs_stride, cudaStream_t st) {
  gemm(s.G, d->n64, 2 * Bp, d->A, d->n64, d->m64, d->n64, s.RHS, d->m64, st);  // [A(D c) ; A(D w)]
  assemble_and_factor(d, s, Bc, damping, st, true);
  enqueue_after_factor(d, s, adev, Bc, Bp, plain, b_rhs, b_rhs_stride, st);
}
static void enqueue_after_factor(const IntoptDense* d, const dn::DenseState& s, const dn::DenseFwdArgs* adev, int Bc, int Bp,
                                 bool plain, const double* b_rhs, long long b_rhs_stride, cudaStream_t st) {
  const int m64 = d->m64, n64 = d->n64;
  {
    dn::SolveArgs sa{};
    sa.rhs[0] = s.RHS;
    sa.rhs[1] = s.RHS + (size_t)Bp * m64;
    sa.add[0] = b_rhs;  // q: r = b + A(D c)   (:666 with (c,b))
    sa.add_stride[0] = b_rhs_stride;
    sa.add[1] = s.AX;  // v: r = eta r1 + A(D w)                                              (:785,666)
    sa.add_stride[1] = m64;
    sa.cadd[1] = s.SC + dn::SC_ETA;
    sa.coef_stride = dn::SC_N;
    solve<2>(d, s, Bc, sa, st);
  }
  gemm(s.RHS, m64, 2 * Bp, d->At, m64, n64, m64, s.ATQV, n64, st);  // [A^T q ; A^T v]
  dn::dense_predict_kernel<<<Bc, 256, 0, st>>>(adev);
  if (!plain) {
    gemm(s.G, n64, Bp, d->A, n64, m64, n64, s.RHS + (size_t)Bp * m64, m64, st);  // A(D w')
    {
      dn::SolveArgs sa{};
      sa.rhs[0] = s.RHS + (size_t)Bp * m64;
      sa.add[0] = s.AX;
      sa.add_stride[0] = m64;
      sa.cadd[0] = s.SC + dn::SC_ETA;
      sa.coef_stride = dn::SC_N;
      solve<1>(d, s, Bc, sa, st);
    }
    gemm(s.RHS + (size_t)Bp * m64, m64, Bp, d->At, m64, n64, m64, s.ATQV + (size_t)Bp * n64, n64, st);  // A^T v'
    dn::dense_step_kernel<<<Bc, 256, 0, st>>>(adev);
  }
}
// loop top: A x, A^T y, residuals + termination tests (counts the instances that enter the iteration)
static cudaError_t enqueue_loop_top(const IntoptDense* d, const dn::DenseState& s, const dn::DenseFwdArgs* adev, int Bc, int Bp,
                                    cudaStream_t st, bool zero_count) {
  const int m64 = d->m64, n64 = d->n64;
  cudaError_t e = zero_count ? cudaMemsetAsync(s.n_active, 0, sizeof(int), st) : cudaSuccess;
  gemm(s.X, n64, Bp, d->A, n64, m64, n64, s.AX, m64, st);   // A x
  gemm(s.Y, m64, Bp, d->At, m64, n64, m64, s.ATY, n64, st);  // A^T y
  dn::dense_resid_kernel<<<Bc, 256, 0, st>>>(adev);
  return e;
}

// Build (or find) the forward graph of a chunk: init -> while (some instance active) { iteration } -> epilogue.
static cudaGraphExec_t forward_graph(const IntoptDense* d, const dn::DenseState& s, const dn::DenseFwdArgs* adev, void* ws, int Bc,
                                     int Bp, double damping, bool plain, const double* b_batch, long long b_stride,
                                     const double* b_rhs, long long b_rhs_stride) {
  for (const auto& g : d->graphs)
    if (g.ws == ws && g.Bc == Bc && g.plain == (int)plain && std::memcmp(&g.damping, &damping, sizeof(double)) == 0 &&
        g.b_batch == b_batch && g.b_stride == b_stride)
      return g.exec;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool ok = cudaGraphCreate(&graph, 0) == cudaSuccess;
  cudaGraphConditionalHandle handle{};
  ok = ok && cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
  cudaGraphNode_t n_init = nullptr, n_loop = nullptr, n_epi = nullptr;
  void* kargs[1] = {(void*)&adev};
  if (ok) {
    cudaKernelNodeParams kp{};
    kp.func = (void*)dn::dense_init_kernel;
    kp.gridDim = dim3(Bc);
    kp.blockDim = dim3(256);
    kp.kernelParams = kargs;
    ok = cudaGraphAddKernelNode(&n_init, graph, nullptr, 0, &kp) == cudaSuccess;
  }
  cudaGraph_t body = nullptr;
  if (ok) {
    cudaGraphNodeParams np{};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    ok = cudaGraphAddNode(&n_loop, graph, &n_init, 1, &np) == cudaSuccess;
    if (ok) body = np.conditional.phGraph_out[0];
  }
  // the leaf of a graph that was filled by capturing one stream (a chain): the node nothing depends on
  auto leaf_of = [](cudaGraph_t g, cudaGraphNode_t* leaf) {
    size_t n = 0;
    if (cudaGraphGetNodes(g, nullptr, &n) != cudaSuccess || n == 0) return false;
    std::vector<cudaGraphNode_t> nodes(n);
    if (cudaGraphGetNodes(g, nodes.data(), &n) != cudaSuccess) return false;
    for (cudaGraphNode_t nd : nodes) {
      size_t nd_out = 0;
      if (cudaGraphNodeGetDependentNodes(nd, nullptr, &nd_out) != cudaSuccess) return false;
      if (nd_out == 0) {
        *leaf = nd;
        return true;
      }
    }
    return false;
  };
  auto capture_into = [&](cudaGraph_t g, const cudaGraphNode_t* deps, size_t ndeps, auto&& enqueue) {
    if (cudaStreamBeginCaptureToGraph(d->cap_stream, g, deps, nullptr, ndeps, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
      return false;
    enqueue();
    cudaGraph_t got = nullptr;
    return cudaStreamEndCapture(d->cap_stream, &got) == cudaSuccess;
  };
  // IF node (condition set by dense_if_cond_kernel) appended to the chain in `g`; returns its body graph
  auto add_if = [&](cudaGraph_t g, cudaGraphConditionalHandle h, cudaGraphNode_t* node_out, cudaGraph_t* body_out) {
    cudaGraphNode_t leaf = nullptr;
    if (!leaf_of(g, &leaf)) return false;
    cudaGraphNodeParams np{};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = h;
    np.conditional.type = cudaGraphCondTypeIf;
    np.conditional.size = 1;
    if (cudaGraphAddNode(node_out, g, &leaf, 1, &np) != cudaSuccess) return false;
    *body_out = np.conditional.phGraph_out[0];
    return true;
  };
  int* n_retry = s.n_active + 1;  // (second word of the counter block at the head of the workspace)
  const bool ladder = d->use_ladder;
  cudaGraphConditionalHandle h_r0{}, h_r1{};
  if (ok && ladder)
    ok = cudaGraphConditionalHandleCreate(&h_r0, graph, 0, cudaGraphCondAssignDefault) == cudaSuccess &&
         cudaGraphConditionalHandleCreate(&h_r1, graph, 0, cudaGraphCondAssignDefault) == cudaSuccess;
  if (ok) {
    // segment 1: loop top, loop condition, right-hand-side products, assembly + factorisation, first ladder test
    ok = capture_into(body, nullptr, 0, [&] {
      enqueue_loop_top(d, s, adev, Bc, Bp, d->cap_stream, false);  // (the condition kernel re-zeroes the count)
      dn::dense_loop_cond_kernel<<<1, 32, 0, d->cap_stream>>>(handle, s.n_active);
      gemm(s.G, d->n64, 2 * Bp, d->A, d->n64, d->m64, d->n64, s.RHS, d->m64, d->cap_stream);
      assemble_and_factor_masked(d, s, Bc, damping, s.active, nullptr, d->cap_stream);
      if (ladder) {
        dn::dense_retry_prep_kernel<<<Bc, 256, 0, d->cap_stream>>>(s, damping, d->rowsq_max, 0, n_retry);
        dn::dense_if_cond_kernel<<<1, 32, 0, d->cap_stream>>>(h_r0, n_retry);
      }
    });
  }
  cudaGraphNode_t tail = nullptr;  // what the rest of the iteration depends on
  if (ok && ladder) {
    // the retry rungs only exist as bodies of IF nodes: an iteration without a breakdown pays two tiny kernels for them
    cudaGraphNode_t if0 = nullptr, if1 = nullptr;
    cudaGraph_t body0 = nullptr, body1 = nullptr;
    ok = add_if(body, h_r0, &if0, &body0);
    ok = ok && capture_into(body0, nullptr, 0, [&] {
      assemble_and_factor_masked(d, s, Bc, damping, s.retry, s.delta_retry, d->cap_stream);
      dn::dense_retry_prep_kernel<<<Bc, 256, 0, d->cap_stream>>>(s, damping, d->rowsq_max, 1, n_retry);
      dn::dense_if_cond_kernel<<<1, 32, 0, d->cap_stream>>>(h_r1, n_retry);
    });
    ok = ok && add_if(body0, h_r1, &if1, &body1);
    ok = ok && capture_into(body1, nullptr, 0, [&] {
      assemble_and_factor_masked(d, s, Bc, damping, s.retry, s.delta_retry, d->cap_stream);
    });
    tail = if0;
  } else if (ok) {
    ok = leaf_of(body, &tail);
  }
  if (ok) {
    ok = capture_into(body, &tail, 1, [&] {
      enqueue_after_factor(d, s, adev, Bc, Bp, plain, b_rhs, b_rhs_stride, d->cap_stream);
    });
  }
  if (ok) {
    cudaKernelNodeParams kp{};
    kp.func = (void*)dn::dense_epilogue_kernel;
    kp.gridDim = dim3(Bc);
    kp.blockDim = dim3(256);
    kp.kernelParams = kargs;
    ok = cudaGraphAddKernelNode(&n_epi, graph, &n_loop, 1, &kp) == cudaSuccess;
  }
  ok = ok && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    cudaGetLastError();  // clear: the caller falls back to the host-driven loop
    return nullptr;
  }
  if (d->graphs.size() >= 16) {  // bounded cache
    cudaGraphExecDestroy(d->graphs.front().exec);
    d->graphs.erase(d->graphs.begin());
  }
  d->graphs.push_back({ws, Bc, (int)plain, damping, b_batch, b_stride, exec});
  return exec;
}

int forward(const IntoptDense* d, int B, double* c, long long c_stride, const double* b_batch, long long b_stride,
            const IntoptOpts* o, const double* init, const intopt::UserMap& um, double* xhat, double* state, int* status,
            int* iters, void* ws, size_t ws_bytes, cudaStream_t st, std::string* err) {
  const int Bc0 = chunk_size(d, B);
  const Layout L0 = make_layout(d, Bc0);
  if (ws_bytes < L0.total) {
    if (err) *err = "workspace too small";
    return -4;
  }
  std::lock_guard<std::mutex> lock(d->mu);
  for (int b0 = 0; b0 < B; b0 += Bc0) {
    const int Bc = std::min(Bc0, B - b0);
    const Layout L = make_layout(d, Bc);
    dn::DenseFwdArgs a{};
    a.s = bind_state(d, L, (char*)ws);
    a.inst0 = b0;
    a.c = c;
    a.c_stride = c_stride;
    a.b_batch = b_batch;
    a.b_stride = b_stride;
    a.b_problem = d->b;
    a.init = init;
    a.um = um;
    a.xhat = xhat;
    a.state = state;
    a.status = status;
    a.iters = iters;
    a.alpha0 = o->alpha0;
    a.tol = o->tol;
    a.thr = o->thr;
    a.damping = o->damping_fwd;
    a.beta = o->beta;
    a.max_iter = o->max_iter;
    a.plain = o->plain != 0;
    const dn::DenseState& s = a.s;
    const int Bp = L.Bp;
    const dn::DenseFwdArgs* adev = reinterpret_cast<const dn::DenseFwdArgs*>((char*)ws + kArgsOffset);
    DN_TRY(cudaMemcpyAsync((void*)adev, &a, sizeof(a), cudaMemcpyHostToDevice, st));
    const double* b_rhs = b_batch ? b_batch + (long long)b0 * b_stride : d->b;
    const long long b_rhs_stride = b_batch ? b_stride : 0;
    if (d->use_graph) {
      cudaGraphExec_t exec = forward_graph(d, s, adev, ws, Bc, Bp, a.damping, a.plain != 0, b_rhs, b_rhs_stride, b_rhs, b_rhs_stride);
      if (exec) {
        DN_TRY(cudaGraphLaunch(exec, st));
        continue;
      }
    }
    // host-driven loop (INTOPT_DENSE_GRAPH=0, or no conditional-node support): one event wait per iteration
    dn::dense_init_kernel<<<Bc, 256, 0, st>>>(adev);
    for (long long it = 0;; ++it) {
      DN_TRY(enqueue_loop_top(d, s, adev, Bc, Bp, st, true));
      DN_TRY(cudaMemcpyAsync(d->h_nactive, s.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      DN_TRY(cudaEventRecord(d->ev, st));
      enqueue_iteration(d, s, adev, Bc, Bp, a.damping, a.plain != 0, b_rhs, b_rhs_stride, st);
      DN_TRY(cudaEventSynchronize(d->ev));  // how many instances entered this iteration (they all left: stop)
      if (*d->h_nactive == 0) break;
      if (it > (long long)o->max_iter + 2) break;
    }
    dn::dense_epilogue_kernel<<<Bc, 256, 0, st>>>(adev);
    DN_TRY(cudaGetLastError());
  }
  return 0;
}

int backward(const IntoptDense* d, int B, int mode, double damping, int nvar, const double* state, const double* c,
             long long c_stride, const double* b_batch, long long b_stride, const intopt::UserMap& um, double* gx,
             double* gc, int* status, void* ws, size_t ws_bytes, cudaStream_t st, std::string* err) {
  const int Bc0 = chunk_size(d, B);
  const Layout L0 = make_layout(d, Bc0);
  if (ws_bytes < L0.total) {
    if (err) *err = "workspace too small";
    return -4;
  }
  for (int b0 = 0; b0 < B; b0 += Bc0) {
    const int Bc = std::min(Bc0, B - b0);
    const Layout L = make_layout(d, Bc);
    dn::DenseBwdArgs a{};
    a.s = bind_state(d, L, (char*)ws);
    a.inst0 = b0;
    a.mode = mode;
    a.nvar = nvar;
    a.damping = damping;
    a.state = state;
    a.c = c;
    a.c_stride = c_stride;
    a.b_batch = b_batch;
    a.b_stride = b_stride;
    a.b_problem = d->b;
    a.gx = gx;
    a.um = um;
    a.gc = gc;
    a.status = status;
    const dn::DenseState& s = a.s;
    const int Bp = L.Bp, m64 = d->m64, n64 = d->n64;
    dn::dense_bwd_prep_kernel<<<Bc, 256, 0, st>>>(a);
    const bool whole = mode == 0;
    if (whole)
      gemm(s.G, n64, 2 * Bp, d->A, n64, m64, n64, s.RHS, m64, st);  // [A(X2 c) ; A(X2 g)]
    else
      gemm(s.G + (size_t)Bp * n64, n64, Bp, d->A, n64, m64, n64, s.RHS + (size_t)Bp * m64, m64, st);  // A(X2 g)
    assemble_and_factor(d, s, Bc, damping, st);  // :1525-1527,1535
    if (whole) {
      dn::SolveArgs sa{};
      sa.rhs[0] = s.RHS;  // q: mu b + A(X2 c)  (:1530)
      sa.rhs[1] = s.RHS + (size_t)Bp * m64;  // v: tau A(X2 g) (:1529)
      sa.add[0] = b_batch ? b_batch + (long long)b0 * b_stride : d->b;
      sa.add_stride[0] = b_batch ? b_stride : 0;
      sa.cadd[0] = s.SC + dn::SB_MU;
      sa.csrc[1] = s.SC + dn::SB_TAU;
      sa.coef_stride = dn::SC_N;
      solve<2>(d, s, Bc, sa, st);
      gemm(s.RHS, m64, 2 * Bp, d->At, m64, n64, m64, s.ATQV, n64, st);
    } else {
      dn::SolveArgs sa{};
      sa.rhs[0] = s.RHS + (size_t)Bp * m64;
      sa.coef_stride = dn::SC_N;
      solve<1>(d, s, Bc, sa, st);
      gemm(s.RHS + (size_t)Bp * m64, m64, Bp, d->At, m64, n64, m64, s.ATQV + (size_t)Bp * n64, n64, st);
    }
    dn::dense_bwd_final_kernel<<<Bc, 256, 0, st>>>(a);
    DN_TRY(cudaGetLastError());
  }
  return 0;
}

// z_b = (A diag(d_b) A^T + delta I)^-1 rhs_b for B instances: the assembly, factorisation and solve kernels of the
// dense path as a stand-alone primitive (rhs: B x m in/out, dvec: B x n, fail: int32[B] out or null).
int normal_solve(const IntoptDense* d, int B, const double* dvec, double delta, double* rhs, int* fail, void* ws,
                 size_t ws_bytes, cudaStream_t st, std::string* err) {
  const int Bc0 = chunk_size(d, B);
  if (ws_bytes < make_layout(d, Bc0).total) {
    if (err) *err = "workspace too small";
    return -4;
  }
  for (int b0 = 0; b0 < B; b0 += Bc0) {
    const int Bc = std::min(Bc0, B - b0);
    const Layout L = make_layout(d, Bc);
    const dn::DenseState s = bind_state(d, L, (char*)ws);
    pad_rows_kernel<<<Bc, 256, 0, st>>>(dvec + (long long)b0 * d->n, d->n, d->n, s.D, d->n64, d->n64, 0.0);
    pad_rows_kernel<<<Bc, 256, 0, st>>>(rhs + (long long)b0 * d->m, d->m, d->m, s.RHS, d->m64, d->m64, 0.0);
    set_flags_kernel<<<(L.Bp + 255) / 256, 256, 0, st>>>(s.active, s.fail, L.Bp);
    assemble_and_factor(d, s, Bc, delta, st);
    dn::SolveArgs sa{};
    sa.rhs[0] = s.RHS;
    sa.coef_stride = dn::SC_N;
    sa.M = s.M;
    sa.Linv = s.Linv;
    sa.m = d->m;
    sa.m64 = d->m64;
    sa.ld = d->m64;
    sa.active = s.active;
    sa.skip = nullptr;
    dn::dense_solve_kernel<1><<<Bc, 256, solve_smem(1, d->m64), st>>>(sa);
    unpad_rows_kernel<<<Bc, 256, 0, st>>>(s.RHS, d->m64, d->m, rhs + (long long)b0 * d->m, d->m);
    if (fail) DN_TRY(cudaMemcpyAsync(fail + b0, s.fail, sizeof(int) * Bc, cudaMemcpyDeviceToDevice, st));
    DN_TRY(cudaGetLastError());
  }
  return 0;
}

int measure_dmma_peak(double* gflops_out, int repeats) {
  std::string* err = nullptr;
  int dev = 0, sms = 0;
  DN_TRY(cudaGetDevice(&dev));
  DN_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 2048;
  double* out = nullptr;
  DN_TRY(cudaMalloc((void**)&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  DN_TRY(cudaEventCreate(&e0));
  DN_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < (repeats > 0 ? repeats : 5) + 2; ++r) {
    DN_TRY(cudaEventRecord(e0));
    dn::dense_dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + r);
    DN_TRY(cudaEventRecord(e1));
    DN_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    DN_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 512.0 * 8.0 * iters * (double)blocks * (threads / 32);  // 2*8*8*4 per DMMA per warp
    const double gf = flops / (ms * 1e-3) * 1e-9;
    if (r >= 2 && gf > best) best = gf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *gflops_out = best;
  return 0;
}

}  // namespace intopt_dense
