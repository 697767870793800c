// Kernel templates of the persistent-CTA solver path (one CTA owns one LP for the whole IPM loop) and the
// table of instantiations.  Each intopt_inst_*.cu compiles one family of instantiations (so the families build
// in parallel) and exports pickers; intopt_capi.cu launches through them.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "intopt_solver.cuh"

constexpr int kMaxMetaInts = 4096;  // schedule arrays up to 16 KB are staged in shared memory
// (The ALLSMEM kernels are launched only when the schedule is staged: the choice is then known at compile time, every
// schedule pointer provably points into shared memory and the lookups compile to LDS instead of generic loads - they sit
// on the dependent chain of every substitution / factorisation step.)

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
  intopt::Meta M = (ALLSMEM || P.meta_ints <= kMaxMetaInts) ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
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

// Tensor memory of a CTA: NV_COLS columns for each of its groups of four warps - 128 columns for a 128-thread CTA (four
// CTAs per SM), 256 for a 256-thread CTA (two), all 512 for a 512-thread CTA (one): CTAs x columns never exceeds the SM's 512
// because registers (128 per thread) admit at most 512 threads of these kernels per SM.  Warp 0 allocates and frees; the address
// travels through shared memory.
template <int COLS>
__device__ __forceinline__ uint32_t tmem_acquire(const CtaCtx& cx, uint32_t* slot) {
  if (cx.warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(slot)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
template <int COLS>
__device__ __forceinline__ void tmem_release(const CtaCtx& cx, uint32_t tmem) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cx.warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(COLS) : "memory");
}

// The same forward with the element-wise n-vectors in tensor memory (see NVTmem): 256-thread CTAs, two per SM, each
// owning 256 of the SM's 512 columns (two groups of four warps x 128 columns), one 512-thread CTA per SM owning all of
// them, or four 128-thread CTAs with 128 columns each.  Picked at launch when every thread's share of an n-vector fits (n <= 7 * threads).
template <int TMAX, bool ALLSMEM>
__global__ void __launch_bounds__(TMAX, TMAX <= 128 ? 4 : TMAX <= 256 ? 2 : 1) intopt_forward_tmem_kernel(const __grid_constant__ intopt::ProblemDev P,
                                                                     const __grid_constant__ intopt::FwdParams a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_inst;
  __shared__ uint32_t s_tmem;
  double* wsi = a.ws + (long long)blockIdx.x * a.ws_stride;
  const intopt::SmemLayout S =
      intopt::carve<ALLSMEM>(P, smem, wsi + intopt::ws_nvec_doubles(P.n), a.factor_in_smem, a.gather_in_smem);
  const CtaCtx cx = make_ctx(S.red);
  const uint32_t tmem = tmem_acquire<intopt::NV_COLS * (TMAX / 128)>(cx, &s_tmem);
  intopt::Meta M = (ALLSMEM || P.meta_ints <= kMaxMetaInts) ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
  if (a.a_vals) intopt::instance_value_streams(P, M, a.vals_ws + (long long)blockIdx.x * a.vals_ws_stride);
  const bool cold = a.cold_factor != nullptr && a.init == nullptr && *a.cold_ok != 0;  // launch-uniform
  if (cold && a.factor_in_smem && threadIdx.x == 0) intopt::mbar_init(S.mbar, 1);
  unsigned served = 0;
  for (;;) {
    if (threadIdx.x == 0) s_inst = atomicAdd(a.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= a.B) break;
    intopt::forward_instance_tmem(cx, P, M, a, inst, wsi, S, tmem, cold, served & 1u);
    ++served;
  }
  tmem_release<intopt::NV_COLS * (TMAX / 128)>(cx, tmem);
}

// One CTA computes the cold-start factor chol(A A^T + delta I) with the instance code path (same CTA size).
template <int TMAX, bool ALLSMEM>
__global__ void __launch_bounds__(TMAX, TMAX <= 128 ? 4 : TMAX <= 256 ? 2 : 1) intopt_coldfactor_kernel(const __grid_constant__ intopt::ProblemDev P,
                                                                  double delta, double* ws, int factor_in_smem,
                                                                  int gather_in_smem, double* out, int* ok) {
  extern __shared__ __align__(16) double smem[];
  const intopt::SmemLayout S = intopt::carve<ALLSMEM>(P, smem, ws + intopt::ws_nvec_doubles(P.n), factor_in_smem, gather_in_smem);
  const CtaCtx cx = make_ctx(S.red);
  const intopt::Meta M = (ALLSMEM || P.meta_ints <= kMaxMetaInts) ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
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
  intopt::Meta M = (ALLSMEM || P.meta_ints <= kMaxMetaInts) ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
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


template <int TMAX, bool ALLSMEM>
__global__ void __launch_bounds__(TMAX, TMAX <= 128 ? 4 : TMAX <= 256 ? 2 : 1) intopt_backward_tmem_kernel(const __grid_constant__ intopt::ProblemDev P,
                                                                      const __grid_constant__ intopt::BwdParams a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_inst;
  __shared__ uint32_t s_tmem;
  double* wsi = a.ws + (long long)blockIdx.x * a.ws_stride;
  const intopt::SmemLayout S =
      intopt::carve<ALLSMEM>(P, smem, wsi + intopt::ws_nvec_doubles(P.n), a.factor_in_smem, a.gather_in_smem);
  const CtaCtx cx = make_ctx(S.red);
  const uint32_t tmem = tmem_acquire<intopt::NV_COLS * (TMAX / 128)>(cx, &s_tmem);
  intopt::Meta M = (ALLSMEM || P.meta_ints <= kMaxMetaInts) ? intopt::stage_meta(cx, P, S.meta) : intopt::global_meta(P);
  if (a.a_vals) intopt::instance_value_streams(P, M, a.vals_ws + (long long)blockIdx.x * a.vals_ws_stride);
  for (;;) {
    if (threadIdx.x == 0) s_inst = atomicAdd(a.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= a.B) break;
    intopt::backward_instance_tmem(cx, P, M, a, inst, wsi, S, tmem);
  }
  tmem_release<intopt::NV_COLS * (TMAX / 128)>(cx, tmem);
}

// Kernel variants: CTA size cap (128: 4 CTAs/SM x 128 registers, 256: 2 x 128, 512: 1 x 128; 192: 2 x 168 only in
// profiling builds) x whether every buffer is in shared memory.
#ifdef INTOPT_PROFILE
#define INTOPT_VARIANTS(F)                                                                                   \
  F(128, true) F(128, false) F(192, true) F(192, false) F(256, true) F(256, false) F(512, true) F(512, false)
#else
#define INTOPT_VARIANTS(F) F(128, true) F(128, false) F(256, true) F(256, false) F(512, true) F(512, false)
#endif

namespace intopt_k {
const void* pick_forward(int threads, bool allsmem, bool plain);
const void* pick_forward_tmem(int threads, bool allsmem);
const void* pick_backward_tmem(int threads, bool allsmem);
const void* pick_backward(int threads, bool allsmem);
const void* pick_cold(int threads, bool allsmem);
void list_forward(std::vector<const void*>& v);
void list_forward_plain(std::vector<const void*>& v);
void list_tmem(std::vector<const void*>& v);
void list_backward(std::vector<const void*>& v);
}  // namespace intopt_k
