// Profiling aids (per-phase cycle buckets, building-block micro-benchmarks).  Compiled ONLY into
// libintopt_b200_profile.so (csrc/build.sh profile: every translation unit with -DINTOPT_PROFILE); the product
// library contains none of this.  Declarations: include/intopt_b200_profile.h.
#ifndef INTOPT_PROFILE
#error "intopt_profile.cu belongs to the profiling build (-DINTOPT_PROFILE)"
#endif
#include <algorithm>
#include <cstdlib>
#include <string>

#include "../../include/intopt_b200_profile.h"
#include "intopt_internal.hpp"
#include "intopt_kernels.cuh"

namespace {
constexpr int kMaxDynSmemP = 232448 - 256;
#define CU_TRYP(expr)                          \
  do {                                         \
    cudaError_t e_ = (expr);                   \
    if (e_ != cudaSuccess) return -100 - (int)e_; \
  } while (0)
bool g_profile_phases = false;
}  // namespace

__device__ unsigned long long g_phase_cycles[32];

unsigned long long* intopt_profile_phase_buckets() {
  if (!g_profile_phases) return nullptr;
  void* sym = nullptr;
  if (cudaGetSymbolAddress(&sym, g_phase_cycles) != cudaSuccess) return nullptr;
  return (unsigned long long*)sym;
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


extern "C" {

int intopt_debug_phase_cycles(int enable, unsigned long long* out32_host) {
  // debug aid: enable/disable per-phase cycle accounting of the forward kernel and read + reset the buckets
  g_profile_phases = enable != 0;
  unsigned long long zero[32] = {0};
  if (out32_host) {
    CU_TRYP(cudaDeviceSynchronize());
    CU_TRYP(cudaMemcpyFromSymbol(out32_host, g_phase_cycles, sizeof(zero)));
  }
  CU_TRYP(cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)));
  return 0;
}

int intopt_debug_microbench(const IntoptProblem* p, int which, int reps, int grid, unsigned long long* cycles_host) {
  // debug aid: cycles per repetition of one solver building block, per CTA (see intopt_microbench_kernel)
  if (!p || !cycles_host || grid <= 0) return -1;
  grid = std::min(grid, p->slots);
  double* ws = nullptr;
  unsigned long long* out = nullptr;
  const size_t per = ws_per_slot_doubles(p);
  CU_TRYP(cudaMalloc((void**)&ws, sizeof(double) * per * (size_t)grid));
  CU_TRYP(cudaMalloc((void**)&out, sizeof(unsigned long long) * (size_t)grid));
  int fis = p->info.factor_in_smem, gis = p->info.gather_in_smem;
  long long stride = (long long)per;
  const void* k = nullptr;
#define PICKM_(T, A) \
  if (!k && p->info.threads <= T && (fis && gis) == A) k = (const void*)intopt_microbench_kernel<T, A>;
  INTOPT_VARIANTS(PICKM_)
  CU_TRYP(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmemP));
#undef PICKM_
  void* args[8] = {(void*)&p->dev, (void*)&which, (void*)&reps, (void*)&ws, (void*)&stride, (void*)&fis, (void*)&gis, (void*)&out};
  CU_TRYP(cudaLaunchKernel(k, dim3(grid), dim3(p->info.threads), args, (size_t)p->info.smem_bytes, 0));
  CU_TRYP(cudaDeviceSynchronize());
  CU_TRYP(cudaMemcpy(cycles_host, out, sizeof(unsigned long long) * (size_t)grid, cudaMemcpyDeviceToHost));
  cudaFree(ws);
  cudaFree(out);
  return 0;
}


}  // extern "C"
