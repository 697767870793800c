// C ABI (include/intopt_b200.h) over the sm_100a solver kernels.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/intopt_b200.h"
#include "intopt_bind.hpp"

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

constexpr int kMaxThreads = 512;
constexpr int kMaxDynSmem = 232448 - 256;  // 227 KB opt-in per CTA on sm_100, minus the kernels' static smem
constexpr size_t kCounterBytes = 256;

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

// Decide what lives in shared memory.  Preference: factor + gather buffer; then factor only; then neither.
constexpr int kMaxMetaInts = 4096;  // schedule arrays up to 16 KB are staged in shared memory

void plan_smem(const intopt::HostStructure& S, int* factor_in_smem, int* gather_in_smem, int* smem_bytes) {
  const int mi = intopt::meta_ints(S);
  const size_t fixed = intopt::smem_fixed_doubles(S) + (mi <= kMaxMetaInts ? (size_t)(mi + 1) / 2 : 0);
  const size_t fac = (size_t)S.bsize, gat = 2 * (size_t)intopt::gather_len(S.n);
  *factor_in_smem = 1;
  *gather_in_smem = 1;
  if ((fixed + fac + gat) * 8 > (size_t)kMaxDynSmem) {
    *gather_in_smem = 0;
    if ((fixed + fac) * 8 > (size_t)kMaxDynSmem) {
      *factor_in_smem = 0;
      if ((fixed + gat) * 8 <= (size_t)kMaxDynSmem) *gather_in_smem = 1;
    }
  }
  *smem_bytes = (int)((fixed + (*factor_in_smem ? fac : 0) + (*gather_in_smem ? gat : 0)) * 8);
}

void fill_info(const intopt::HostStructure& S, IntoptInfo* info) {
  info->m = S.m;
  info->n = S.n;
  info->nnz = S.nnz;
  info->envelope = S.bsize;
  info->dense_lower = (int)((long long)S.m * (S.m + 1) / 2);
  info->asm_terms = S.asm_terms;
  info->chol_macs = S.chol_macs;
  int g = 0;
  plan_smem(S, &info->factor_in_smem, &g, &info->smem_bytes);
  info->gather_in_smem = g;
  // CTA size: INTOPT_THREADS overrides.  Default 256 (two LPs resident per SM for the experiment shapes); small
  // problems take 128 threads so that four LPs share an SM; problems whose shared memory admits a single resident
  // CTA take 512 threads (the only way to put more warps on that SM).
  int thr = env_int("INTOPT_THREADS", 0);
  if (thr <= 0) {
    thr = 256;
    if (S.m <= 64 && info->smem_bytes <= (kMaxDynSmem + 1024) / 4 - 1024) thr = 128;
    if (info->smem_bytes > (kMaxDynSmem + 1024) / 2 - 1024) thr = 512;
  }
  info->threads = thr;
  info->ordering = S.ordering;
  info->scalar_envelope = S.sky_size;
  info->asm_padded = S.asm_padded;
  info->solve_macs = S.solve_macs;
}

template <class T>
cudaError_t upload(const std::vector<T>& h, const T** d, std::vector<void*>& allocs) {
  *d = nullptr;
  if (h.empty()) return cudaSuccess;
  T* p = nullptr;
  cudaError_t e = cudaMalloc((void**)&p, h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  allocs.push_back(p);
  *d = p;
  return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

// Cold-start factor cache (see FwdParams::cold_factor): one entry per forward damping value seen.
struct ColdEntry {
  double delta;
  double* factor;  // bsize doubles
  int* ok;         // device flag: 1 if the factorisation succeeded
  cudaEvent_t ready;
  bool done;
};

struct IntoptProblem {
  intopt::ProblemDev dev{};
  IntoptInfo info{};
  const double* b_dev = nullptr;
  int device = 0;
  int sms = 0;
  int slots = 0;  // persistent CTAs per launch
  std::vector<void*> allocs;
  // lazily filled, guarded by `mu`; the problem data itself is immutable after create
  mutable std::mutex mu;
  mutable std::vector<ColdEntry> cold;
  mutable std::vector<void*> cold_allocs;
};

__device__ __forceinline__ CtaCtx make_ctx(double* red) {
  CtaCtx cx;
  cx.tid = threadIdx.x;
  cx.nthreads = blockDim.x;
  cx.lane = threadIdx.x & 31;
  cx.warp = threadIdx.x >> 5;
  cx.nwarps = blockDim.x >> 5;
  cx.red = red;
  cx.bar_id = 0;
  cx.bar_n = blockDim.x;
  return cx;
}

// Persistent CTAs: each pulls the next LP instance from an atomic counter until the batch is exhausted,
// so finished LPs free their SM slot without any host involvement.
// PLAIN: pc=False upstream (one direction per iteration) — its own instantiation, so that the predictor-corrector kernels
// are compiled exactly as if the option did not exist.
template <int TMAX, bool ALLSMEM, bool PLAIN = false>
__global__ void __launch_bounds__(TMAX, TMAX <= 128 ? 4 : TMAX <= 256 ? 2 : 1) intopt_forward_kernel(const __grid_constant__ intopt::ProblemDev P,
                                                               const __grid_constant__ intopt::FwdParams a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_inst;
  double* wsi = a.ws + (long long)blockIdx.x * a.ws_stride;
  const intopt::SmemLayout S =
      intopt::carve<ALLSMEM>(P, smem, wsi + intopt::ws_nvec_doubles(P.n), a.factor_in_smem, a.gather_in_smem);
  const CtaCtx cx = make_ctx(S.red);
  intopt::Meta M = P.meta_ints <= kMaxMetaInts ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
  if (a.a_vals) intopt::instance_value_streams(P, M, a.vals_ws + (long long)blockIdx.x * a.vals_ws_stride);
  const bool cold = a.cold_factor != nullptr && a.init == nullptr && *a.cold_ok != 0;  // launch-uniform
  if (cold && a.factor_in_smem && threadIdx.x == 0) intopt::mbar_init(S.mbar, 1);
  unsigned served = 0;  // instances this CTA has run = completed phases of the prefetch mbarrier
  for (;;) {
    if (threadIdx.x == 0) s_inst = atomicAdd(a.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= a.B) break;
    intopt::forward_instance_t<PLAIN>(cx, P, M, a, inst, wsi, S, cold, served & 1u);
    ++served;
  }
}

// One CTA computes the cold-start factor chol(A A^T + delta I) with the instance code path (same CTA size).
template <int TMAX, bool ALLSMEM>
__global__ void __launch_bounds__(TMAX, TMAX <= 128 ? 4 : TMAX <= 256 ? 2 : 1) intopt_coldfactor_kernel(const __grid_constant__ intopt::ProblemDev P,
                                                                  double delta, double* ws, int factor_in_smem,
                                                                  int gather_in_smem, double* out, int* ok) {
  extern __shared__ __align__(16) double smem[];
  const intopt::SmemLayout S = intopt::carve<ALLSMEM>(P, smem, ws + intopt::ws_nvec_doubles(P.n), factor_in_smem, gather_in_smem);
  const CtaCtx cx = make_ctx(S.red);
  const intopt::Meta M = P.meta_ints <= kMaxMetaInts ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
  intopt::cold_factor_build(cx, P, M, delta, ws, S, out, ok);
}

template <int TMAX, bool ALLSMEM>
__global__ void __launch_bounds__(TMAX, TMAX <= 128 ? 4 : TMAX <= 256 ? 2 : 1) intopt_backward_kernel(const __grid_constant__ intopt::ProblemDev P,
                                                                const __grid_constant__ intopt::BwdParams a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_inst;
  double* wsi = a.ws + (long long)blockIdx.x * a.ws_stride;
  const intopt::SmemLayout S =
      intopt::carve<ALLSMEM>(P, smem, wsi + intopt::ws_nvec_doubles(P.n), a.factor_in_smem, a.gather_in_smem);
  const CtaCtx cx = make_ctx(S.red);
  intopt::Meta M = P.meta_ints <= kMaxMetaInts ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
  if (a.a_vals) intopt::instance_value_streams(P, M, a.vals_ws + (long long)blockIdx.x * a.vals_ws_stride);
  for (;;) {
    if (threadIdx.x == 0) s_inst = atomicAdd(a.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= a.B) break;
    intopt::backward_instance(cx, P, M, a, inst, wsi, S);
  }
}

// experimental copy of intopt::spmv_rows (run-length path) with switchable stages, for the micro-benchmark only
__device__ __forceinline__ void spmv_rows_dbg(const CtaCtx& cx, const intopt::ProblemDev& P, const double* src, double* out,
                                              int flags) {
  int slot = cx.warp;
  uint32_t hn = 0u;
  int rown = 0xffff;
  if (slot < P.rs_slots) {
    hn = P.rs_hdr[slot * 32 + cx.lane];
    rown = P.rs_row[slot * 32 + cx.lane];
  }
  for (; slot < P.rs_slots; slot += cx.nwarps) {
    const uint32_t h = hn;
    const int row = rown;
    const int nxt = slot + cx.nwarps;
    if (nxt < P.rs_slots) {
      hn = P.rs_hdr[nxt * 32 + cx.lane];
      rown = P.rs_row[nxt * 32 + cx.lane];
    }
    double s = (flags & 1) ? intopt::run_total16(src + intopt::rs_start(h), intopt::rs_len(h), intopt::rs_rot(h)) : src[h & 0xffffu];
    if (flags & 2) {
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, s, off);
        const int r = __shfl_up_sync(0xffffffffu, row, off);
        if (cx.lane >= off && r == row) s += v;
      }
    }
    const int rnext = __shfl_down_sync(0xffffffffu, row, 1);
    if (row != 0xffff && (cx.lane == 31 || rnext != row)) out[row] = s;
  }
}

// Debug aid for profiles/: runs ONE building block of the solver `reps` times per CTA and reports the cycles per
// repetition seen by thread 0 (barrier to barrier).  Not used by the product path.
template <int TMAX, bool ALLSMEM>
__global__ void __launch_bounds__(TMAX, TMAX <= 128 ? 4 : TMAX <= 256 ? 2 : 1) intopt_microbench_kernel(const __grid_constant__ intopt::ProblemDev P,
                                                                  int which, int reps, double* ws, long long ws_stride,
                                                                  int factor_in_smem, int gather_in_smem,
                                                                  unsigned long long* out) {
  extern __shared__ __align__(16) double smem[];
  double* wsi = ws + (long long)blockIdx.x * ws_stride;
  const intopt::SmemLayout S = intopt::carve<ALLSMEM>(P, smem, wsi + intopt::ws_nvec_doubles(P.n), factor_in_smem, gather_in_smem);
  const CtaCtx cx = make_ctx(S.red);
  const intopt::Meta M = P.meta_ints <= kMaxMetaInts ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
  const int n = P.n;
  double* D = wsi + (long long)intopt::W_D * n;
  double* c = wsi + (long long)intopt::W_P * n;
  CTA_FOR(k, n) {
    D[k] = 1.0 + 1e-3 * (k % 7);
    c[k] = 0.5 + (k % 5);
    intopt::gstore(S.G, k, 1.0 + 0.01 * (k % 3));
  }
  CTA_FOR(i, P.mv) {
    S.q[i] = i < P.m ? 1.0 + 0.1 * (i % 4) : 0.0;
    S.v[i] = i < P.m ? 2.0 - 0.1 * (i % 3) : 0.0;
    S.y[i] = 0.0;
    S.r1[i] = 0.0;
    S.bs[i] = 0.0;
  }
  if (cx.tid == 0) S.G[intopt::gpad(n)] = 0.0;
  __syncthreads();
  if (which >= 5 && which <= 7) {  // a valid factor for the factor/solve blocks
    intopt::assemble_normal_matrix(cx, P, M, D, S.G, S.scan, 1e-6, S.L);
    if (which != 5) intopt::block_cholesky(cx, P, M, S.L, S.dscr, S.flag);
    __syncthreads();
  }
  const long long t0 = clock64();
  double sink = 0.0;
  for (int r = 0; r < reps; ++r) {
    switch (which) {
      case 0: break;
      case 1: intopt::spmv_rows(cx, P, M, S.G, S.q); break;
      case 2: intopt::at_mult<1>(cx, P, M, S.q, nullptr, S.G, nullptr); break;
      case 3: intopt::at_mult<2>(cx, P, M, S.q, S.v, S.G, S.G + intopt::gather_len(n)); break;
      case 4: intopt::assemble_normal_matrix(cx, P, M, D, S.G, S.scan, 1e-6, S.L); break;
      case 5:
        intopt::assemble_normal_matrix(cx, P, M, D, S.G, S.scan, 1e-6, S.L);
        intopt::block_cholesky(cx, P, M, S.L, S.dscr, S.flag);
        break;
      case 6: {
        CTA_FOR(i, P.mv) {
          S.q[i] = i < P.m ? 1.0 : 0.0;
          S.v[i] = i < P.m ? 2.0 : 0.0;
        }
        __syncthreads();
        double* const rhs2[2] = {S.q, S.v};
        intopt::block_solve<2>(cx, P, M, S.L, rhs2, S.ybuf);
      } break;
      case 7: {
        CTA_FOR(i, P.mv) S.v[i] = i < P.m ? 2.0 : 0.0;
        __syncthreads();
        double* const rhs1[1] = {S.v};
        intopt::block_solve<1>(cx, P, M, S.L, rhs1, S.ybuf);
      } break;
      case 8: CTA_FOR(k, n) intopt::gstore(S.G, k, D[k] * c[k]); break;
      case 9: {
        double a5[5] = {1.0, 2.0, 3.0, 4.0, 5.0};
        intopt::cta_sum<5>(cx, a5);
        sink += a5[0];
      } break;
      case 10: CTA_FOR(k, n) D[k] = D[k] * 1.0000001 + c[k] * 1e-9; break;
      case 18: spmv_rows_dbg(cx, P, S.G, S.q, 0); break;
      case 19: spmv_rows_dbg(cx, P, S.G, S.q, 1); break;
      case 20: spmv_rows_dbg(cx, P, S.G, S.q, 2); break;
      case 21: spmv_rows_dbg(cx, P, S.G, S.q, 3); break;
      case 34: intopt::assemble_normal_matrix(cx, P, M, D, S.G, S.scan, 1e-6, S.L, 1); break;  // prefix scan + zero fill only
      case 11: {  // dependent DFMA chain, 256 long
        double z = S.G[cx.tid & 7];
#pragma unroll 16
        for (int i = 0; i < 256; ++i) z = fma(z, 0.999999, 1e-9);
        sink += z;
      } break;
      case 12: {  // dependent shared-memory load chain, 256 long (pointer chasing through S.meta-like ints)
        int idx = cx.tid & 31;
        const double* gp = S.G;
#pragma unroll 16
        for (int i = 0; i < 256; ++i) idx = (int)gp[idx] + (idx & 1);
        sink += idx;
      } break;
      case 13: {  // dependent 64-bit shuffle chain, 256 long
        double z = S.G[cx.tid & 7];
#pragma unroll 16
        for (int i = 0; i < 256; ++i) z = __shfl_xor_sync(0xffffffffu, z, 1) + 1.0;
        sink += z;
      } break;
      case 14: {  // dependent DADD chain, 256 long
        double z = S.G[cx.tid & 7];
#pragma unroll 16
        for (int i = 0; i < 256; ++i) z = z + 1.25;
        sink += z;
      } break;
      case 15: {  // dependent global (L2) load chain, 64 long
        int idx = cx.tid;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) idx = (int)D[idx] + (idx & 1) + cx.tid - 1;
        sink += idx;
      } break;
      case 16: {  // rsqrt chain 64 long
        double z = S.G[cx.tid & 7] + 2.0;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) z = rsqrt(z) + 1.5;
        sink += z;
      } break;
      case 17: {  // division chain 64 long
        double z = S.G[cx.tid & 7] + 2.0;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) z = 3.0 / z + 1.5;
        sink += z;
      } break;
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (cx.tid == 0) out[blockIdx.x] = (unsigned long long)((t1 - t0) / (reps > 0 ? reps : 1)) + (sink == 12345.0 ? 1 : 0);
}

// Host-glue helper: dst[b, j] = idx[j] >= 0 ? scale[j] * src[b, idx[j]] : constv[j]   (scale NULL = 1, constv NULL = 0),
// converting between float and double on the way.  One launch replaces the cast + zero-fill + index / index_put
// chains that map the user's cost / gradient vectors to the standard form and the solution back.
template <class TS, class TD>
__global__ void __launch_bounds__(256) intopt_gather_convert_kernel(const TS* __restrict__ src, long long src_stride, int B,
                                                                    int n_out, const int* __restrict__ idx,
                                                                    const double* __restrict__ scale,
                                                                    const double* __restrict__ constv, TD* __restrict__ dst,
                                                                    long long dst_stride) {
  const long long total = (long long)B * n_out;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / n_out), j = (int)(e - (long long)b * n_out);
    const int k = idx ? idx[j] : j;
    double v;
    if (k >= 0) {
      v = (double)src[(long long)b * src_stride + k];
      if (scale) v *= scale[j];
    } else {
      v = constv ? constv[j] : 0.0;
    }
    dst[(long long)b * dst_stride + j] = (TD)v;
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

__device__ unsigned long long g_phase_cycles[32];

// Kernel variants: CTA size cap (128: 4 CTAs/SM x 128 registers, 192: 2 x 168, 256: 2 x 128, 512: 1 x 128) x whether every
// buffer is in shared memory.
#ifdef INTOPT_DEV_BUILD  // quick developer builds (INTOPT_NVCC_EXTRA=-DINTOPT_DEV_BUILD): default CTA size only
#define INTOPT_VARIANTS(F) F(128, true) F(128, false) F(256, true) F(256, false) F(512, true) F(512, false)
#else
#define INTOPT_VARIANTS(F)                                                                                   \
  F(128, true) F(128, false) F(192, true) F(192, false) F(256, true) F(256, false) F(512, true) F(512, false)
#endif
static std::vector<const void*> all_kernels() {
  std::vector<const void*> v;
#define ADD_(T, A)                                          \
  v.push_back((const void*)intopt_forward_kernel<T, A>);    \
  v.push_back((const void*)intopt_forward_kernel<T, A, true>); \
  v.push_back((const void*)intopt_backward_kernel<T, A>);   \
  v.push_back((const void*)intopt_coldfactor_kernel<T, A>); \
  v.push_back((const void*)intopt_microbench_kernel<T, A>);
  INTOPT_VARIANTS(ADD_)
#undef ADD_
  return v;
}
static const void* pick_kernel(bool forward, int threads, bool allsmem, bool plain = false) {
#define PICK_(T, A)                                                                                        \
  if (threads <= T && allsmem == A)                                                                        \
    return forward ? (plain ? (const void*)intopt_forward_kernel<T, A, true> : (const void*)intopt_forward_kernel<T, A>) \
                   : (const void*)intopt_backward_kernel<T, A>;
  INTOPT_VARIANTS(PICK_)
#undef PICK_
  return nullptr;
}
static const void* pick_cold_kernel(int threads, bool allsmem) {
#define PICKC_(T, A) \
  if (threads <= T && allsmem == A) return (const void*)intopt_coldfactor_kernel<T, A>;
  INTOPT_VARIANTS(PICKC_)
#undef PICKC_
  return nullptr;
}

namespace {
bool g_profile_phases = false;
size_t vals_ws_doubles(const IntoptProblem* p) {
  return (size_t)p->dev.nnz + (size_t)p->dev.n_csc_pad + (size_t)p->dev.n_asm_pad;
}
size_t ws_per_slot_doubles(const IntoptProblem* p) {
  size_t per = (size_t)intopt::ws_nvec_doubles(p->info.n);
  if (!p->info.factor_in_smem) per += (size_t)p->info.envelope;
  if (!p->info.gather_in_smem) per += 2 * (size_t)intopt::gather_len(p->info.n);
  per = (per + 1) & ~(size_t)1;
  if (p->dev.inst_values) per += vals_ws_doubles(p);  // this instance's copies of the value streams of A
  return (per + 1) & ~(size_t)1;
}

// Returns the cold-start factor for `delta` (computing it on `stream` the first time) or nulls if unavailable.
int cold_factor_for(const IntoptProblem* p, double delta, cudaStream_t stream, const double** factor, const int** ok) {
  *factor = nullptr;
  *ok = nullptr;
  if (env_int("INTOPT_NO_COLD", 0)) return 0;
  std::lock_guard<std::mutex> lock(p->mu);
  for (ColdEntry& e : p->cold) {
    if (std::memcmp(&e.delta, &delta, sizeof(double)) != 0) continue;
    if (!e.done) {
      if (cudaEventQuery(e.ready) == cudaSuccess)
        e.done = true;
      else
        CU_TRY(cudaStreamWaitEvent(stream, e.ready, 0));
    }
    *factor = e.factor;
    *ok = e.ok;
    return 0;
  }
  if (p->cold.size() >= 8) return 0;  // unusual: many different dampings on one problem -> plain path
  ColdEntry e{};
  e.delta = delta;
  e.done = false;
  double* ws = nullptr;
  CU_TRY(cudaMalloc((void**)&e.factor, sizeof(double) * (size_t)p->info.envelope));
  p->cold_allocs.push_back(e.factor);
  CU_TRY(cudaMalloc((void**)&e.ok, sizeof(int)));
  p->cold_allocs.push_back(e.ok);
  CU_TRY(cudaMalloc((void**)&ws, sizeof(double) * ws_per_slot_doubles(p)));
  p->cold_allocs.push_back(ws);
  CU_TRY(cudaMemsetAsync(e.ok, 0, sizeof(int), stream));
  CU_TRY(cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming));
  {
    int fis = p->info.factor_in_smem, gis = p->info.gather_in_smem;
    void* args[7] = {(void*)&p->dev, (void*)&delta, (void*)&ws, (void*)&fis, (void*)&gis, (void*)&e.factor, (void*)&e.ok};
    CU_TRY(cudaLaunchKernel(pick_cold_kernel(p->info.threads, fis && gis), dim3(1), dim3(p->info.threads), args,
                            (size_t)p->info.smem_bytes, stream));
  }
  CU_TRY(cudaEventRecord(e.ready, stream));
  p->cold.push_back(e);
  *factor = e.factor;
  *ok = e.ok;
  return 0;
}
}  // namespace

extern "C" {

int intopt_version(void) { return INTOPT_B200_VERSION; }
const char* intopt_last_error(void) { return g_err.c_str(); }

int intopt_analyze(int m, int n, const int* rowptr, const int* colidx, const double* val, IntoptInfo* info,
                   int* row_perm) {
  if (!info) return fail(-1, "info is NULL");
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rowptr, colidx, val, S);
  if (rc == -5) return fail(rc, "problem too large for the gather-list assembly (sum over columns of nnz*(nnz+1)/2 > 2^28)");
  if (rc) return fail(rc, "malformed CSR input");
  fill_info(S, info);
  if (row_perm) std::memcpy(row_perm, S.perm.data(), sizeof(int) * m);
  return 0;
}

int intopt_problem_create(IntoptProblem** out, int m, int n, const int* rowptr, const int* colidx,
                          const double* val, const double* b, int flags) {
  if (!out) return fail(-1, "out is NULL");
  *out = nullptr;
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rowptr, colidx, val, S, -1, (flags & INTOPT_FLAG_INSTANCE_VALUES) != 0);
  if (rc == -5) return fail(rc, "problem too large for the gather-list assembly (sum over columns of nnz*(nnz+1)/2 > 2^28)");
  if (rc) return fail(rc, "malformed CSR input");
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev <= 0) return fail(-10, "no CUDA device: intopt_b200 has no CPU fallback");
  IntoptProblem* p = new (std::nothrow) IntoptProblem();
  if (!p) return fail(-2, "out of host memory");
  cudaError_t err = cudaSuccess;
  auto up = [&](const auto& v) {
    using T = typename std::decay<decltype(v)>::type::value_type;
    const T* d = nullptr;
    cudaError_t e = upload(v, &d, p->allocs);
    if (e != cudaSuccess) err = e;
    return d;
  };
  fill_info(S, &p->info);
  intopt::bind_problem(S, p->dev, up);
  if (b) {
    std::vector<double> bperm(m);
    for (int i = 0; i < m; ++i) bperm[i] = b[S.perm[i]];
    p->b_dev = up(bperm);
  }
  if (err != cudaSuccess) {
    intopt_problem_destroy(p);
    return fail(-100 - (int)err, std::string("device upload: ") + cudaGetErrorString(err));
  }
  CU_TRY(cudaGetDevice(&p->device));
  CU_TRY(cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, p->device));
  if (p->info.threads < 32 || p->info.threads > kMaxThreads || p->info.threads % 32) {
    intopt_problem_destroy(p);
    return fail(-12, "INTOPT_THREADS must be a multiple of 32 in [32, 512]");
  }
  for (const void* k : all_kernels())
    CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  int occ = 0;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &occ, pick_kernel(true, p->info.threads, p->info.factor_in_smem && p->info.gather_in_smem), p->info.threads,
      (size_t)p->info.smem_bytes));
  if (occ < 1) {
    intopt_problem_destroy(p);
    return fail(-11, "solver kernel does not fit on an SM (shared memory / registers)");
  }
  occ = std::min(occ, env_int("INTOPT_MAX_CTAS_PER_SM", 8));
  p->info.ctas_per_sm = occ;
  p->slots = p->sms * occ;
  CU_TRY(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

int intopt_problem_destroy(IntoptProblem* p) {
  if (!p) return 0;
  for (void* a : p->allocs)
    if (a) cudaFree(a);
  for (void* a : p->cold_allocs)
    if (a) cudaFree(a);
  for (ColdEntry& e : p->cold) cudaEventDestroy(e.ready);
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
  const size_t slots = (size_t)std::min(B, p->slots);
  return kCounterBytes + ws_per_slot_doubles(p) * sizeof(double) * slots;
}

int intopt_forward(const IntoptProblem* p, int B, const double* c, long long c_stride, const double* b_batch,
                   long long b_stride, const IntoptOpts* o, const double* init, double* xhat, double* state,
                   int* status, int* iters, void* ws, size_t ws_bytes, void* stream) {
  return intopt_forward_values(p, B, c, c_stride, b_batch, b_stride, nullptr, 0, o, init, xhat, state, status, iters, ws,
                               ws_bytes, stream);
}

int intopt_forward_values(const IntoptProblem* p, int B, const double* c, long long c_stride, const double* b_batch,
                          long long b_stride, const double* a_vals, long long a_stride, const IntoptOpts* o,
                          const double* init, double* xhat, double* state, int* status, int* iters, void* ws,
                          size_t ws_bytes, void* stream) {
  return intopt_forward_mapped(p, B, const_cast<double*>(c), c_stride, b_batch, b_stride, a_vals, a_stride, o, init, nullptr,
                               xhat, state, status, iters, ws, ws_bytes, stream);
}

static void fill_user_map(intopt::UserMap& u, const IntoptUserMap* m) {
  u = intopt::UserMap{};
  if (!m) return;
  u.in = m->in_dev;
  u.in_f32 = m->in_is_f32;
  u.in_stride = m->in_stride;
  u.in_idx = m->in_idx_dev;
  u.in_scale = m->in_scale_dev;
  u.out = m->out_dev;
  u.out_f32 = m->out_is_f32;
  u.out_stride = m->out_stride;
  u.n_out = m->n_out;
  u.out_idx = m->out_idx_dev;
  u.out_const = m->out_const_dev;
}

int intopt_forward_mapped(const IntoptProblem* p, int B, double* c, long long c_stride, const double* b_batch,
                          long long b_stride, const double* a_vals, long long a_stride, const IntoptOpts* o,
                          const double* init, const IntoptUserMap* um, double* xhat, double* state, int* status,
                          int* iters, void* ws, size_t ws_bytes, void* stream) {
  if (!p || !c || !o || !state || !status || !iters || !ws) return fail(-1, "NULL argument");
  if (!xhat && !(um && um->out_dev)) return fail(-1, "neither xhat_dev nor a mapped output was given");
  if (um && um->in_dev && (!um->in_idx_dev || c_stride < p->info.n)) return fail(-8, "mapped input needs in_idx_dev and one writable c row per instance");
  if (um && um->out_dev && !um->out_idx_dev) return fail(-8, "mapped output needs out_idx_dev");
  if (a_vals && !p->dev.inst_values) return fail(-7, "per-instance values need a problem created with INTOPT_FLAG_INSTANCE_VALUES");
  if (!a_vals && p->dev.inst_values) return fail(-7, "this problem was created for per-instance values: pass a_vals_dev");
  if (B <= 0) return 0;
  if (!p->b_dev && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
  if (ws_bytes < intopt_workspace_bytes(p, B)) return fail(-4, "workspace too small");
  intopt::FwdParams a{};
  a.B = B;
  a.c = c;
  a.c_stride = c_stride;
  a.b_batch = b_batch;
  a.b_stride = b_stride;
  a.b_problem = p->b_dev;
  a.init = init;
  fill_user_map(a.um, um);
  a.a_vals = a_vals;
  a.a_stride = a_stride;
  a.xhat = xhat;
  a.state = state;
  a.status = status;
  a.iters = iters;
  a.counter = (int*)ws;
  a.ws = (double*)((char*)ws + kCounterBytes);
  a.factor_in_smem = p->info.factor_in_smem;
  a.gather_in_smem = p->info.gather_in_smem;
  a.ws_stride = (long long)ws_per_slot_doubles(p);
  // the value streams sit at the end of every slot
  a.vals_ws = a.ws + (a.ws_stride - (long long)((vals_ws_doubles(p) + 1) & ~(size_t)1));
  a.vals_ws_stride = a.ws_stride;
  a.alpha0 = o->alpha0;
  a.tol = o->tol;
  a.thr = o->thr;
  a.damping = o->damping_fwd;
  a.max_iter = o->max_iter;
  a.plain = o->plain != 0;
  a.beta = o->beta;
  a.phase_cycles = nullptr;
  a.debug = env_int("INTOPT_DEBUG", 0);
  if (!init && !a_vals) {  // (the cold-start factor is a constant of the problem only if A is)
    int rc = cold_factor_for(p, a.damping, (cudaStream_t)stream, &a.cold_factor, &a.cold_ok);
    if (rc) return rc;
  }
  if (g_profile_phases) {
    void* sym = nullptr;
    CU_TRY(cudaGetSymbolAddress(&sym, g_phase_cycles));
    a.phase_cycles = (unsigned long long*)sym;
  }
  const int grid = std::min(B, p->slots);
  CU_TRY(cudaMemsetAsync(ws, 0, kCounterBytes, (cudaStream_t)stream));
  {
    void* args[2] = {(void*)&p->dev, (void*)&a};
    CU_TRY(cudaLaunchKernel(pick_kernel(true, p->info.threads, a.factor_in_smem && a.gather_in_smem, a.plain != 0), dim3(grid),
                            dim3(p->info.threads), args, (size_t)p->info.smem_bytes, (cudaStream_t)stream));
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

int intopt_backward(const IntoptProblem* p, int B, int mode, double damping, int nvar, const double* state,
                    const double* c, long long c_stride, const double* b_batch, long long b_stride,
                    const double* gx, double* gc, int* status, void* ws, size_t ws_bytes, void* stream) {
  return intopt_backward_values(p, B, mode, damping, nvar, state, c, c_stride, b_batch, b_stride, nullptr, 0, gx, gc, status,
                                ws, ws_bytes, stream);
}

int intopt_backward_values(const IntoptProblem* p, int B, int mode, double damping, int nvar, const double* state,
                           const double* c, long long c_stride, const double* b_batch, long long b_stride,
                           const double* a_vals, long long a_stride, const double* gx, double* gc, int* status, void* ws,
                           size_t ws_bytes, void* stream) {
  return intopt_backward_mapped(p, B, mode, damping, nvar, state, c, c_stride, b_batch, b_stride, a_vals, a_stride, nullptr,
                                const_cast<double*>(gx), gc, status, ws, ws_bytes, stream);
}

int intopt_backward_mapped(const IntoptProblem* p, int B, int mode, double damping, int nvar, const double* state,
                           const double* c, long long c_stride, const double* b_batch, long long b_stride,
                           const double* a_vals, long long a_stride, const IntoptUserMap* um, double* gx, double* gc,
                           int* status, void* ws, size_t ws_bytes, void* stream) {
  if (!p || !c || !state || !gx || !gc || !status || !ws) return fail(-1, "NULL argument");
  if (um && ((um->in_dev && !um->in_idx_dev) || (um->out_dev && !um->out_idx_dev))) return fail(-8, "mapped input / output needs its index array");
  if ((a_vals != nullptr) != (p->dev.inst_values != 0)) return fail(-7, "a_vals_dev must be given exactly for problems created with INTOPT_FLAG_INSTANCE_VALUES");
  if (B <= 0) return 0;
  if (mode != 0 && mode != 1) return fail(-5, "mode must be 0 (whole) or 1 (alter)");
  if (nvar < 0 || nvar > p->info.n) return fail(-6, "nvar out of range");
  if (mode == 0 && !p->b_dev && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
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
  a.b_problem = p->b_dev;
  a.gx = gx;
  fill_user_map(a.um, um);
  a.a_vals = a_vals;
  a.a_stride = a_stride;
  a.gc = gc;
  a.status = status;
  a.counter = (int*)ws;
  a.ws = (double*)((char*)ws + kCounterBytes);
  a.factor_in_smem = p->info.factor_in_smem;
  a.gather_in_smem = p->info.gather_in_smem;
  a.ws_stride = (long long)ws_per_slot_doubles(p);
  // the value streams sit at the end of every slot
  a.vals_ws = a.ws + (a.ws_stride - (long long)((vals_ws_doubles(p) + 1) & ~(size_t)1));
  a.vals_ws_stride = a.ws_stride;
  const int grid = std::min(B, p->slots);
  CU_TRY(cudaMemsetAsync(ws, 0, kCounterBytes, (cudaStream_t)stream));
  {
    void* args[2] = {(void*)&p->dev, (void*)&a};
    CU_TRY(cudaLaunchKernel(pick_kernel(false, p->info.threads, a.factor_in_smem && a.gather_in_smem), dim3(grid),
                            dim3(p->info.threads), args, (size_t)p->info.smem_bytes, (cudaStream_t)stream));
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

int intopt_debug_phase_cycles(int enable, unsigned long long* out32_host) {
  // debug aid: enable/disable per-phase cycle accounting of the forward kernel and read + reset the buckets
  g_profile_phases = enable != 0;
  unsigned long long zero[32] = {0};
  if (out32_host) {
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpyFromSymbol(out32_host, g_phase_cycles, sizeof(zero)));
  }
  CU_TRY(cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)));
  return 0;
}

int intopt_debug_microbench(const IntoptProblem* p, int which, int reps, int grid, unsigned long long* cycles_host) {
  // debug aid: cycles per repetition of one solver building block, per CTA (see intopt_microbench_kernel)
  if (!p || !cycles_host || grid <= 0) return fail(-1, "bad argument");
  grid = std::min(grid, p->slots);
  double* ws = nullptr;
  unsigned long long* out = nullptr;
  const size_t per = ws_per_slot_doubles(p);
  CU_TRY(cudaMalloc((void**)&ws, sizeof(double) * per * (size_t)grid));
  CU_TRY(cudaMalloc((void**)&out, sizeof(unsigned long long) * (size_t)grid));
  int fis = p->info.factor_in_smem, gis = p->info.gather_in_smem;
  long long stride = (long long)per;
  const void* k = nullptr;
#define PICKM_(T, A) \
  if (!k && p->info.threads <= T && (fis && gis) == A) k = (const void*)intopt_microbench_kernel<T, A>;
  INTOPT_VARIANTS(PICKM_)
#undef PICKM_
  void* args[8] = {(void*)&p->dev, (void*)&which, (void*)&reps, (void*)&ws, (void*)&stride, (void*)&fis, (void*)&gis, (void*)&out};
  CU_TRY(cudaLaunchKernel(k, dim3(grid), dim3(p->info.threads), args, (size_t)p->info.smem_bytes, 0));
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(cycles_host, out, sizeof(unsigned long long) * (size_t)grid, cudaMemcpyDeviceToHost));
  cudaFree(ws);
  cudaFree(out);
  return 0;
}

int intopt_gather_convert(const void* src, int src_is_f32, long long src_stride, int B, int n_out, const int* idx,
                          const double* scale, const double* constv, void* dst, int dst_is_f32, long long dst_stride,
                          void* stream) {
  if (!src || !dst || B < 0 || n_out < 0) return fail(-1, "bad argument");
  const long long total = (long long)B * n_out;
  if (total == 0) return 0;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_is_f32 && dst_is_f32)
    intopt_gather_convert_kernel<float, float><<<blocks, threads, 0, st>>>((const float*)src, src_stride, B, n_out, idx, scale, constv, (float*)dst, dst_stride);
  else if (src_is_f32)
    intopt_gather_convert_kernel<float, double><<<blocks, threads, 0, st>>>((const float*)src, src_stride, B, n_out, idx, scale, constv, (double*)dst, dst_stride);
  else if (dst_is_f32)
    intopt_gather_convert_kernel<double, float><<<blocks, threads, 0, st>>>((const double*)src, src_stride, B, n_out, idx, scale, constv, (float*)dst, dst_stride);
  else
    intopt_gather_convert_kernel<double, double><<<blocks, threads, 0, st>>>((const double*)src, src_stride, B, n_out, idx, scale, constv, (double*)dst, dst_stride);
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
