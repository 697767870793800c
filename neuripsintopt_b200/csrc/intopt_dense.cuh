// Lock-step dense path: device code.
//
// For problems whose normal matrix M = A diag(d) A^T is dense and too large for one CTA's shared memory (dense A,
// m > ~180) the whole batch advances one interior-point iteration at a time through batch-wide kernels:
//   * every product with A runs for all instances at once as an FP64 tensor-core GEMM (DMMA m8n8k4) against the one
//     shared copy of A (vectors are the rows of a B x n matrix);
//   * M_b = A diag(d_b) A^T is a batched, scaled SYRK on DMMA tiles (lower tiles only), A tiles staged through a
//     cp.async pipeline and re-used from L2 by every instance;
//   * the factorisation is a tiled left-looking Cholesky (64-wide block columns, DMMA updates, the 64x64 diagonal
//     factor + inverse in shared memory), the solves are blocked substitutions streaming L once per sweep;
//   * the element-wise phases (residuals, Mehrotra centring, ratio test, step, termination tests) are one CTA per
//     instance with fixed-order reductions, masked by a per-instance `active` flag, so finished LPs drop out.
// Arithmetic follows ip_model_whole.py:761-917,1385-1447 (forward) and :1502-1560 / ip_model_alter.py:1516-1624
// (backward) exactly as intopt_solver.cuh does for the persistent-CTA path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "intopt_solver.cuh"

namespace dn {

constexpr int NB = 64;      // block-column width of the tiled Cholesky
constexpr int KT = 16;      // k-extent of one pipeline stage
constexpr int LDT = KT + 4; // row stride of a staged tile in doubles: == 4 (mod 16) makes the DMMA fragment reads conflict-free
constexpr int STAGES = 3;
constexpr int LDD = NB + 1; // row stride of the 64x64 diagonal block in shared memory

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col): lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2(l%4)], C[l/4][2(l%4)+1]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------
// CTA-level GEMM main loop, "NT" form: acc(TM x TN) += P(TM x K) * Q(TN x K)^T with both operands k-contiguous
// in global memory (leading dimensions ldp / ldq, multiples of 2 doubles; K a multiple of KT).  SCALE multiplies
// column k of P by dsc[k] on the fragment (the diag(d) of the normal matrix).  `compute` = false lets a warp
// take part in the staging only (warps above the diagonal of a diagonal SYRK tile).
// ---------------------------------------------------------------------------------------------------------
template <int TM, int TN, int WM, int WN, int NTHR, bool SCALE>
struct GemmNT {
  static constexpr int MT = WM / 8, NT = WN / 8;
  static constexpr int WARPS_N = TN / WN;
  static constexpr int STAGE_DOUBLES = (TM + TN) * LDT + (SCALE ? KT : 0);
  static constexpr int SMEM_DOUBLES = STAGES * STAGE_DOUBLES;
  static_assert((TM / WM) * (TN / WN) * 32 == NTHR, "warp grid must cover the tile");

  __device__ static void zero(double (&acc)[MT][NT][2]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  __device__ static void run(const double* __restrict__ P, long long ldp, const double* __restrict__ Q, long long ldq,
                             int K, const double* __restrict__ dsc, double* smem, double (&acc)[MT][NT][2],
                             bool compute = true) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp / WARPS_N) * WM, wn = (warp % WARPS_N) * WN;
    const int nk = K / KT;
    auto load_stage = [&](int kt, int s) {
      double* Ps = smem + s * STAGE_DOUBLES;
      double* Qs = Ps + TM * LDT;
      const int k0 = kt * KT;
      for (int idx = tid; idx < (TM + TN) * 8; idx += NTHR) {
        if (idx < TM * 8) {
          const int r = idx >> 3, c = idx & 7;
          cp_async16(Ps + r * LDT + 2 * c, P + (long long)r * ldp + k0 + 2 * c);
        } else {
          const int j = idx - TM * 8, r = j >> 3, c = j & 7;
          cp_async16(Qs + r * LDT + 2 * c, Q + (long long)r * ldq + k0 + 2 * c);
        }
      }
      if (SCALE && tid < 8) cp_async16(Qs + TN * LDT + 2 * tid, dsc + k0 + 2 * tid);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nk) load_stage(s, s);
      cp_commit();
    }
    const int kk = lane & 3, rr = lane >> 2;
    for (int kt = 0; kt < nk; ++kt) {
      cp_wait<STAGES - 2>();
      __syncthreads();
      {
        const int nxt = kt + STAGES - 1;
        if (nxt < nk) load_stage(nxt, nxt % STAGES);
        cp_commit();
      }
      if (compute) {
        const double* Ps = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double* Qs = Ps + TM * LDT;
        const double* Ds = Qs + TN * LDT;
#pragma unroll
        for (int k4 = 0; k4 < KT / 4; ++k4) {
          double a[MT], b[NT];
#pragma unroll
          for (int i = 0; i < MT; ++i) a[i] = Ps[(wm + 8 * i + rr) * LDT + 4 * k4 + kk];
#pragma unroll
          for (int j = 0; j < NT; ++j) b[j] = Qs[(wn + 8 * j + rr) * LDT + 4 * k4 + kk];
          if (SCALE) {
            const double d = Ds[4 * k4 + kk];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] *= d;
          }
#pragma unroll
          for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
      }
    }
    cp_wait<0>();
    __syncthreads();  // every warp is done with the staged tiles: the caller may reuse `smem`
  }

  // (row, col) of accumulator element e of tile (i, j) inside the CTA tile
  __device__ static int row_of(int i) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return (warp / WARPS_N) * WM + 8 * i + (lane >> 2);
  }
  __device__ static int col_of(int j) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return (warp % WARPS_N) * WN + 8 * j + 2 * (lane & 3);
  }
};

// index of the row of the lower triangle that holds flat pair index p (pairs (a,b), b <= a, row-major)
__device__ __forceinline__ int tri_row(int p) {
  int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while ((a + 1) * (a + 2) / 2 <= p) ++a;
  while (a * (a + 1) / 2 > p) --a;
  return a;
}

// ---------------------------------------------------------------------------------------------------------
// Out(rows x N) = In(rows x K) * W(N x K)^T: every product with A (W = A, K = n) or A^T (W = A^T, K = m) for the whole
// batch.  rows, N multiples of the tile, K a multiple of KT; padded rows / columns hold zeros or harmless numbers.
// grid (N / 64, rows / 128), 256 threads.
// ---------------------------------------------------------------------------------------------------------
using GemmProd = GemmNT<128, 64, 32, 32, 256, false>;
__global__ void __launch_bounds__(256) dense_gemm_kernel(const double* __restrict__ In, long long ld_in,
                                                         const double* __restrict__ W, long long ld_w, int K,
                                                         double* __restrict__ Out, long long ld_out) {
  extern __shared__ __align__(16) double smem[];
  const int tn = blockIdx.x, tm = blockIdx.y;
  double acc[GemmProd::MT][GemmProd::NT][2];
  GemmProd::zero(acc);
  GemmProd::run(In + (long long)tm * 128 * ld_in, ld_in, W + (long long)tn * 64 * ld_w, ld_w, K, nullptr, smem, acc);
#pragma unroll
  for (int i = 0; i < GemmProd::MT; ++i)
#pragma unroll
    for (int j = 0; j < GemmProd::NT; ++j) {
      const int r = tm * 128 + GemmProd::row_of(i), c = tn * 64 + GemmProd::col_of(j);
      double2 v;
      v.x = acc[i][j][0];
      v.y = acc[i][j][1];
      *reinterpret_cast<double2*>(Out + (long long)r * ld_out + c) = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// M_b = A diag(d_b) A^T + delta I (lower tiles; rows >= m get a unit diagonal so that the padded matrix stays
// positive definite).  ip_model_whole.py:772-775 / :1525-1535.  grid (T(T+1)/2 tiles, B), T = m64 / TS.
// ---------------------------------------------------------------------------------------------------------
template <int TS, int NTHR>
struct SyrkCfg {
  using G = GemmNT<TS, TS, 32, 32, NTHR, true>;
};
template <int TS, int NTHR>
__global__ void __launch_bounds__(NTHR) dense_syrk_kernel(const double* __restrict__ A, int n64, int m, int m64,
                                                          const double* __restrict__ D, long long ld_d, double delta,
                                                          double* __restrict__ Mout, const int* __restrict__ active,
                                                          const double* __restrict__ delta_arr) {
  using G = typename SyrkCfg<TS, NTHR>::G;
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  if (delta_arr) delta = delta_arr[b];  // (retry rungs of the failure ladder: per-instance damping)
  const int ti = tri_row(blockIdx.x), tj = blockIdx.x - ti * (ti + 1) / 2;
  const int warp = threadIdx.x >> 5;
  const int wi = warp / G::WARPS_N, wj = warp % G::WARPS_N;
  const bool compute = !(ti == tj && wj > wi);  // warp tiles strictly above the diagonal are never read
  double acc[G::MT][G::NT][2];
  G::zero(acc);
  G::run(A + (long long)ti * TS * n64, n64, A + (long long)tj * TS * n64, n64, n64, D + (long long)b * ld_d, smem, acc,
         compute);
  if (!compute) return;
  double* Mb = Mout + (long long)b * m64 * m64;
#pragma unroll
  for (int i = 0; i < G::MT; ++i)
#pragma unroll
    for (int j = 0; j < G::NT; ++j) {
      const int r = ti * TS + G::row_of(i), c = tj * TS + G::col_of(j);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = c + e;
        if (cc > r) continue;
        double v = acc[i][j][e];
        if (cc == r) v = r < m ? v + delta : 1.0;
        Mb[(long long)r * m64 + cc] = v;
      }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same assembly with the A tiles staged by TMA: one elected thread issues 2-D tensor copies
// (cp.async.bulk.tensor, the tile box {16 columns x TS rows} described by a CUtensorMap of the padded A, 128-byte
// swizzle) plus a bulk copy of the 16 weights of the stage; the copies complete on an mbarrier per pipeline stage.  No
// thread spends issue slots or registers on addresses, and the 128-byte swizzle (16-byte chunk c of row r lands at chunk
// c ^ (r & 7)) makes the DMMA fragment reads conflict free without padding.  Every instance of the batch pulls the same
// A tiles: after the first CTAs they come out of L2.
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) TmaDesc {
  unsigned long long opaque[16];  // a CUtensorMap (128 bytes), kept opaque here so that this header needs no <cuda.h>
};
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const TmaDesc* desc, int x, int y, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int TS, int NTHR>
struct SyrkTma {
  static constexpr int WM = 32, WN = 32, MT = 4, NT = 4;
  static constexpr int WARPS_N = TS / WN;
  static constexpr int TILE_BYTES = TS * KT * 8;                    // one operand tile of a stage (TS rows x 128 bytes)
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES + 1024;         // two operand tiles + the weights (padded to 1 KB)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 64;  // (+ alignment slack + the mbarriers)
  static_assert((TS / WM) * (TS / WN) * 32 == NTHR, "warp grid must cover the tile");
};

template <int TS, int NTHR>
__global__ void __launch_bounds__(NTHR) dense_syrk_tma_kernel(const __grid_constant__ TmaDesc tmapA, int n64, int m, int m64,
                                                              const double* __restrict__ D, long long ld_d, double delta,
                                                              double* __restrict__ Mout, const int* __restrict__ active,
                                                              const double* __restrict__ delta_arr) {
  using C = SyrkTma<TS, NTHR>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  if (delta_arr) delta = delta_arr[b];  // (retry rungs of the failure ladder: per-instance damping)
  // 1 KB aligned stage buffers (the 128-byte swizzle pattern repeats every 8 rows = 1 KB)
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + STAGES * C::STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ti = tri_row(blockIdx.x), tj = blockIdx.x - ti * (ti + 1) / 2;
  const int wi = warp / C::WARPS_N, wj = warp % C::WARPS_N;
  const int wm = wi * C::WM, wn = wj * C::WN;
  const bool compute = !(ti == tj && wj > wi);
  const double* dsc = D + (long long)b * ld_d;
  const int nk = n64 / KT;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int kt) {  // one thread: the stage's two tiles and its weights, all on one mbarrier
    const int s = kt % STAGES;
    unsigned char* st = base + s * C::STAGE_BYTES;
    mbar_expect_tx(bars + s, 2 * C::TILE_BYTES + KT * 8);
    tma_load_2d(st, &tmapA, kt * KT, ti * TS, bars + s);
    tma_load_2d(st + C::TILE_BYTES, &tmapA, kt * KT, tj * TS, bars + s);
    bulk_load_1d(st + 2 * C::TILE_BYTES, dsc + kt * KT, KT * 8, bars + s);
  };
  if (tid == 0) {
    for (int s = 0; s < STAGES && s < nk; ++s) issue(s);
  }
  double acc[C::MT][C::NT][2];
#pragma unroll
  for (int i = 0; i < C::MT; ++i)
#pragma unroll
    for (int j = 0; j < C::NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int kk = lane & 3, rr = lane >> 2;
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt % STAGES;
    mbar_wait(bars + s, (unsigned)((kt / STAGES) & 1));
    if (compute) {
      const unsigned char* Pb = base + s * C::STAGE_BYTES;
      const unsigned char* Qb = Pb + C::TILE_BYTES;
      const double* Ds = reinterpret_cast<const double*>(Qb + C::TILE_BYTES);
#pragma unroll
      for (int k4 = 0; k4 < KT / 4; ++k4) {
        // element (row r, column k) of a swizzled tile: byte r * 128 + (((k >> 1) ^ (r & 7)) << 4) + (k & 1) * 8.
        // The 8 rows of a DMMA fragment are taken 2 apart (tile t of a warp = rows 16 (t / 2) + (t % 2) + 2 rr): the
        // 16 lanes of a half warp then see r & 7 = {0,2,4,6} + t % 2, i.e. eight different 16-byte chunks, and the
        // 8-byte fragment loads are conflict free (consecutive rows would put two lanes on every chunk).
        const int k = 4 * k4 + kk;
        double a[C::MT], bq[C::NT];
#pragma unroll
        for (int i = 0; i < C::MT; ++i) {
          const int r = wm + 16 * (i >> 1) + (i & 1) + 2 * rr;
          a[i] = *reinterpret_cast<const double*>(Pb + r * 128 + (((k >> 1) ^ (r & 7)) << 4) + (k & 1) * 8);
        }
#pragma unroll
        for (int j = 0; j < C::NT; ++j) {
          const int r = wn + 16 * (j >> 1) + (j & 1) + 2 * rr;
          bq[j] = *reinterpret_cast<const double*>(Qb + r * 128 + (((k >> 1) ^ (r & 7)) << 4) + (k & 1) * 8);
        }
        const double d = Ds[k];
#pragma unroll
        for (int i = 0; i < C::MT; ++i) a[i] *= d;
#pragma unroll
        for (int i = 0; i < C::MT; ++i)
#pragma unroll
          for (int j = 0; j < C::NT; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], bq[j]);
      }
    }
    __syncthreads();  // every warp is done with stage s: it may be refilled
    if (tid == 0 && kt + STAGES < nk) issue(kt + STAGES);
  }
  if (!compute) return;
  double* Mb = Mout + (long long)b * m64 * m64;
#pragma unroll
  for (int i = 0; i < C::MT; ++i)
#pragma unroll
    for (int j = 0; j < C::NT; ++j) {
      // (same row interleave as the fragment loads: accumulator (rr, 2 kk + e) of tile (i, j))
      const int r = ti * TS + wm + 16 * (i >> 1) + (i & 1) + 2 * rr;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = tj * TS + wn + 16 * (j >> 1) + (j & 1) + 2 * (2 * kk + e);
        if (cc > r) continue;
        double v = acc[i][j][e];
        if (cc == r) v = r < m ? v + delta : 1.0;
        Mb[(long long)r * m64 + cc] = v;
      }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 64x64 diagonal block: Cholesky in shared memory, then the explicit inverse of the factor (what the panel
// TRSM and the substitutions multiply by).  S (64 x 64, row stride LDD, lower part valid) is overwritten by L;
// X (same shape) receives L^-1 (lower, zeros above).  256 threads.  tmp: 64 * LDD doubles, rinv: 64 doubles.
// A non-positive / non-finite pivot sets *fail and continues with pivot 1 (the reference's cho_factor
// failure, ip_model_whole.py:742,755-758).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void diag8_factor(double* S, double* rinv, int* fail) {
  // one warp; lane (l & 7) owns row l & 7 of the 8x8 block at S (row stride LDD); lanes 8..31 mirror lanes 0..7
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, i = lane & 7;
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = S[i * LDD + k];
  bool bad = false;
  double rv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double piv = __shfl_sync(full, a[j], j);
    if (!(piv > 0.0) || !(piv < 1e300)) {
      bad = true;
      piv = 1.0;
    }
    const double r = rsqrt(piv);
    rv[j] = r;
    a[j] = (i == j) ? piv * r : a[j] * r;
#pragma unroll
    for (int k = j + 1; k < 8; ++k) {
      const double lkj = __shfl_sync(full, a[j], k);
      a[k] = fma(-a[j], lkj, a[k]);
    }
  }
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k <= i) S[i * LDD + k] = a[k];
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) rinv[k] = rv[k];
    if (bad) *fail = 1;
  }
}

__device__ void chol64_factor_invert(double* S, double* X, double* tmp, double* rinv, int* fail) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int kb = 0; kb < 8; ++kb) {
    const int k0 = 8 * kb;
    if (warp == 0) diag8_factor(S + k0 * LDD + k0, rinv + k0, fail);
    __syncthreads();
    const int nrows = NB - (k0 + 8);
    if (tid < nrows) {  // panel: row r of L_(.,kb) by substitution with the reciprocal pivots
      const int r = k0 + 8 + tid;
      double c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = S[r * LDD + k0 + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double xj = c[j] * rinv[k0 + j];
        c[j] = xj;
#pragma unroll
        for (int k = j + 1; k < 8; ++k) c[k] = fma(-xj, S[(k0 + k) * LDD + k0 + j], c[k]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) S[r * LDD + k0 + j] = c[j];
    }
    __syncthreads();
    for (int idx = tid; idx < nrows * nrows; idx += blockDim.x) {  // trailing update (lower part)
      const int i = idx / nrows, j = idx - i * nrows;
      if (j > i) continue;
      const double* pi = S + (k0 + 8 + i) * LDD + k0;
      const double* pj = S + (k0 + 8 + j) * LDD + k0;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s = fma(pi[k], pj[k], s);
      S[(k0 + 8 + i) * LDD + k0 + 8 + j] -= s;
    }
    __syncthreads();
  }
  // ---- inverse: diagonal 8x8 blocks first (warp w -> block w, lane c < 8 -> column c), zeros above the diagonal
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx >> 6, c = idx & 63;
    X[r * LDD + c] = 0.0;
  }
  __syncthreads();
  if (lane < 8) {
    const int k0 = 8 * warp, c = lane;
    double x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < r) s = fma(-S[(k0 + r) * LDD + k0 + k], (k >= c) ? x[k] : 0.0, s);
      x[r] = (r >= c) ? s * rinv[k0 + r] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r >= c) X[(k0 + r) * LDD + k0 + c] = x[r];
  }
  __syncthreads();
  // ---- off-diagonal blocks by block diagonals: X_IJ = -X_II * sum_{K=J}^{I-1} L_IK X_KJ
  for (int d = 1; d < 8; ++d) {
    const int nblk = 8 - d;
    for (int idx = tid; idx < nblk * 64; idx += blockDim.x) {
      const int J = idx >> 6, I = J + d, r = (idx >> 3) & 7, c = idx & 7;
      double s = 0.0;
      for (int k = 8 * J; k < 8 * I; ++k) s = fma(S[(8 * I + r) * LDD + k], X[k * LDD + 8 * J + c], s);
      tmp[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < nblk * 64; idx += blockDim.x) {
      const int J = idx >> 6, I = J + d, r = (idx >> 3) & 7, c = idx & 7;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k <= r) s = fma(X[(8 * I + r) * LDD + 8 * I + k], tmp[(J << 6) + (k << 3) + c], s);
      X[(8 * I + r) * LDD + 8 * J + c] = -s;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tiled left-looking Cholesky, step k (block column k of every active instance):
//   col kernel : T_ik = C_ik - sum_{j<k} L_ij L_kj^T for the row tiles i >= k (DMMA); the CTA of the diagonal
//                tile then factors T_kk and writes L_kk and L_kk^-1.                         grid (nb - k, B)
//   trsm kernel: L_ik = T_ik L_kk^-T for i > k (DMMA against the explicit inverse).            grid (nb - k - 1, B)
// M is B x m64 x m64 row-major (lower part), Linv is B x nb x 64 x 64.
// ---------------------------------------------------------------------------------------------------------
using GemmChol = GemmNT<64, 64, 32, 16, 256, false>;
// (the factor-and-invert step needs the block, its inverse and 7 x 64 doubles of scratch: 71 KB, three CTAs per SM - the
// diagonal-tile CTAs run a serial pivot chain, so residency is what hides it)
constexpr int CHOL_TMP_DOUBLES = 7 * 64;
constexpr int CHOL_SMEM_DOUBLES = (GemmChol::SMEM_DOUBLES > 2 * NB * LDD + CHOL_TMP_DOUBLES + NB + 8)
                                      ? GemmChol::SMEM_DOUBLES
                                      : 2 * NB * LDD + CHOL_TMP_DOUBLES + NB + 8;

__global__ void __launch_bounds__(256) dense_chol_col_kernel(double* __restrict__ M, int m64, int k,
                                                             double* __restrict__ Linv, int* __restrict__ fail,
                                                             const int* __restrict__ active) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const int i = k + blockIdx.x, nb = m64 / NB;
  double* Mb = M + (long long)b * m64 * m64;
  double acc[GemmChol::MT][GemmChol::NT][2];
  GemmChol::zero(acc);
  if (k > 0)
    GemmChol::run(Mb + (long long)i * NB * m64, m64, Mb + (long long)k * NB * m64, m64, NB * k, nullptr, smem, acc);
  double* Ct = Mb + (long long)i * NB * m64 + (long long)k * NB;
  if (i != k) {
#pragma unroll
    for (int a = 0; a < GemmChol::MT; ++a)
#pragma unroll
      for (int c = 0; c < GemmChol::NT; ++c) {
        double2* p = reinterpret_cast<double2*>(Ct + (long long)GemmChol::row_of(a) * m64 + GemmChol::col_of(c));
        double2 v = *p;
        v.x -= acc[a][c][0];
        v.y -= acc[a][c][1];
        *p = v;
      }
    return;
  }
  double* S = smem;
  double* X = S + NB * LDD;
  double* tmp = X + NB * LDD;
  double* rinv = tmp + CHOL_TMP_DOUBLES;
  int* sfail = reinterpret_cast<int*>(rinv + NB);
  if (threadIdx.x == 0) *sfail = 0;
#pragma unroll
  for (int a = 0; a < GemmChol::MT; ++a)
#pragma unroll
    for (int c = 0; c < GemmChol::NT; ++c) {
      const int r = GemmChol::row_of(a), cc = GemmChol::col_of(c);
      const double2 v = *reinterpret_cast<const double2*>(Ct + (long long)r * m64 + cc);
      S[r * LDD + cc] = v.x - acc[a][c][0];
      S[r * LDD + cc + 1] = v.y - acc[a][c][1];
    }
  __syncthreads();
  chol64_factor_invert(S, X, tmp, rinv, sfail);
  double* Lk = Linv + ((long long)b * nb + k) * NB * NB;
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
    const int r = idx >> 6, c = idx & 63;
    if (c <= r) Ct[(long long)r * m64 + c] = S[r * LDD + c];
    Lk[idx] = X[r * LDD + c];
  }
  if (threadIdx.x == 0 && *sfail) fail[b] = 1;
}

__global__ void __launch_bounds__(256) dense_chol_trsm_kernel(double* __restrict__ M, int m64, int k,
                                                              const double* __restrict__ Linv,
                                                              const int* __restrict__ active) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const int i = k + 1 + blockIdx.x, nb = m64 / NB;
  double* T = M + (long long)b * m64 * m64 + (long long)i * NB * m64 + (long long)k * NB;
  const double* Lk = Linv + ((long long)b * nb + k) * NB * NB;
  double acc[GemmChol::MT][GemmChol::NT][2];
  GemmChol::zero(acc);
  GemmChol::run(T, m64, Lk, NB, NB, nullptr, smem, acc);
#pragma unroll
  for (int a = 0; a < GemmChol::MT; ++a)
#pragma unroll
    for (int c = 0; c < GemmChol::NT; ++c) {
      double2 v;
      v.x = acc[a][c][0];
      v.y = acc[a][c][1];
      *reinterpret_cast<double2*>(T + (long long)GemmChol::row_of(a) * m64 + GemmChol::col_of(c)) = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Blocked substitution L L^T z = r for NR right-hand sides of one instance per CTA (256 threads); L is streamed from
// global memory exactly once per sweep, the diagonal blocks act through their explicit inverses.
// The right-hand sides are formed on the way in:  r_q = cadd[q] * add[q] + csrc[q] * rhs_q   (add may be null), which is
// r = r2 + A(D r1) of _sym_solve (ip_model_whole.py:663-681) / :1529-1530 of the backward.
// ---------------------------------------------------------------------------------------------------------
struct SolveArgs {
  const double* M;      // B x m64 x m64
  const double* Linv;   // B x nb x 64 x 64
  int m, m64;
  double* rhs[2];       // per right-hand side: B x ld (in: product term, out: solution)
  long long ld;
  const double* add[2]; // per right-hand side: added vector (instance stride add_stride[q]) or null
  long long add_stride[2];
  const double* cadd[2];  // per right-hand side: per-instance scalar multiplying `add` (stride coef_stride) or null (= 1)
  const double* csrc[2];  // ... multiplying the product term already in rhs, or null (= 1)
  int coef_stride;
  const int* active;
  const int* skip;  // optional per-instance flag: non-zero = leave the right-hand sides alone (failed factorisation)
};

template <int NR>
__global__ void __launch_bounds__(256) dense_solve_kernel(const SolveArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.x;
  if (a.active && !a.active[b]) return;
  if (a.skip && a.skip[b]) return;
  const int tid = threadIdx.x, m64 = a.m64, nb = m64 / NB;
  double* r = smem;                 // NR x m64
  double* yb = r + NR * m64;        // NR x 64: current pivot block
  double* red = yb + NR * NB;       // 16 x NR x 64 partial sums of the backward sweep
  const double* Mb = a.M + (long long)b * m64 * m64;
  const double* Lb = a.Linv + (long long)b * nb * NB * NB;
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    const double ca = a.cadd[q] ? a.cadd[q][(long long)b * a.coef_stride] : 1.0;
    const double cs = a.csrc[q] ? a.csrc[q][(long long)b * a.coef_stride] : 1.0;
    const double* add = a.add[q] ? a.add[q] + (long long)b * a.add_stride[q] : nullptr;
    const double* src = a.rhs[q] + (long long)b * a.ld;
    for (int i = tid; i < m64; i += blockDim.x)
      r[q * m64 + i] = i < a.m ? (add ? ca * add[i] : 0.0) + (a.csrc[q] ? cs * src[i] : src[i]) : 0.0;
  }
  __syncthreads();
  const int row4 = tid >> 2, part = tid & 3;
  // ---- forward: L y = r ------------------------------------------------------------------------------
  for (int k = 0; k < nb; ++k) {
    {  // y_k = Linv_kk r_k: 4 lanes per row, 16 interleaved columns each
      const double* Lr = Lb + (long long)k * NB * NB + row4 * NB;
      double s[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q) s[q] = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = 2 * part + 8 * j;
        const double2 l2 = *reinterpret_cast<const double2*>(Lr + c);
#pragma unroll
        for (int q = 0; q < NR; ++q) s[q] = fma(l2.y, r[q * m64 + k * NB + c + 1], fma(l2.x, r[q * m64 + k * NB + c], s[q]));
      }
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        s[q] += __shfl_xor_sync(0xffffffffu, s[q], 1);
        s[q] += __shfl_xor_sync(0xffffffffu, s[q], 2);
      }
      __syncthreads();  // every partial product has read r_k
      if (part == 0) {
#pragma unroll
        for (int q = 0; q < NR; ++q) {
          yb[q * NB + row4] = s[q];
          r[q * m64 + k * NB + row4] = s[q];
        }
      }
      __syncthreads();
    }
    for (int i0 = (k + 1) * NB; i0 < m64; i0 += NB) {  // r_i -= L_ik y_k
      const int i = i0 + row4;
      const double* Lr = Mb + (long long)i * m64 + k * NB;
      double s[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q) s[q] = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = 2 * part + 8 * j;
        const double2 l2 = *reinterpret_cast<const double2*>(Lr + c);
#pragma unroll
        for (int q = 0; q < NR; ++q) s[q] = fma(l2.y, yb[q * NB + c + 1], fma(l2.x, yb[q * NB + c], s[q]));
      }
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        s[q] += __shfl_xor_sync(0xffffffffu, s[q], 1);
        s[q] += __shfl_xor_sync(0xffffffffu, s[q], 2);
        if (part == 0) r[q * m64 + i] -= s[q];
      }
    }
    __syncthreads();
  }
  // ---- backward: L^T z = y ---------------------------------------------------------------------------
  const int c4 = tid & 15, rg = tid >> 4;  // 4 columns per thread, 16 row groups
  for (int k = nb - 1; k >= 0; --k) {
    double s[NR][4];
#pragma unroll
    for (int q = 0; q < NR; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[q][e] = 0.0;
    for (int i = (k + 1) * NB + rg; i < m64; i += 16) {  // sum_{i>k} L_ik^T z_i
      const double* Lr = Mb + (long long)i * m64 + k * NB + 4 * c4;
      const double2 l0 = *reinterpret_cast<const double2*>(Lr), l1 = *reinterpret_cast<const double2*>(Lr + 2);
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const double z = r[q * m64 + i];
        s[q][0] = fma(l0.x, z, s[q][0]);
        s[q][1] = fma(l0.y, z, s[q][1]);
        s[q][2] = fma(l1.x, z, s[q][2]);
        s[q][3] = fma(l1.y, z, s[q][3]);
      }
    }
#pragma unroll
    for (int q = 0; q < NR; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[(rg * NR + q) * NB + 4 * c4 + e] = s[q][e];
    __syncthreads();
    if (tid < NR * NB) {  // s_k = y_k - sum (fixed order)
      const int q = tid / NB, c = tid - q * NB;
      double t = 0.0;
      for (int g = 0; g < 16; ++g) t += red[(g * NR + q) * NB + c];
      yb[q * NB + c] = r[q * m64 + k * NB + c] - t;
    }
    __syncthreads();
    // z_k = Linv_kk^T s_k
#pragma unroll
    for (int q = 0; q < NR; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[q][e] = 0.0;
    for (int rr = rg; rr < NB; rr += 16) {
      const double* Lr = Lb + (long long)k * NB * NB + rr * NB + 4 * c4;
      const double2 l0 = *reinterpret_cast<const double2*>(Lr), l1 = *reinterpret_cast<const double2*>(Lr + 2);
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const double z = yb[q * NB + rr];
        s[q][0] = fma(l0.x, z, s[q][0]);
        s[q][1] = fma(l0.y, z, s[q][1]);
        s[q][2] = fma(l1.x, z, s[q][2]);
        s[q][3] = fma(l1.y, z, s[q][3]);
      }
    }
#pragma unroll
    for (int q = 0; q < NR; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[(rg * NR + q) * NB + 4 * c4 + e] = s[q][e];
    __syncthreads();
    if (tid < NR * NB) {
      const int q = tid / NB, c = tid - q * NB;
      double t = 0.0;
      for (int g = 0; g < 16; ++g) t += red[(g * NR + q) * NB + c];
      r[q * m64 + k * NB + c] = t;
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    double* dst = a.rhs[q] + (long long)b * a.ld;
    for (int i = tid; i < m64; i += blockDim.x) dst[i] = r[q * m64 + i];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Element-wise phases: one CTA (256 threads) per instance, fixed-order reductions.
// ---------------------------------------------------------------------------------------------------------
template <int N>
__device__ void block_sum(double (&v)[N], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[k * 32 + w];
    v[k] = s;
  }
  __syncthreads();
}
__device__ inline double block_min(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = red[0];
  for (int w = 1; w < nw; ++w) s = fmin(s, red[w]);
  __syncthreads();
  return s;
}

// per-instance scalars (doubles), SC_N per instance
enum { SC_TAU = 0, SC_KAPPA, SC_MU, SC_R3, SC_RP0, SC_RD0, SC_RG0, SC_CP, SC_BQ, SC_DTAU, SC_DKAPPA, SC_GAMMA, SC_ETA,
       SC_XT, SC_ONE, SC_N = 16 };

struct DenseState {
  int B, Bp, m, n, m64, n64;
  // batch matrices (Bp rows): n-vectors with row stride n64, m-vectors with row stride m64
  double *X, *T, *Y, *AX, *ATY, *R2, *D, *W, *P, *DX, *DXDT;
  double* G;     // 2 Bp x n64: gather inputs of the products with A ([D c ; D w], [X2 g ; X2 c])
  double* RHS;   // 2 Bp x m64: [q ; v]
  double* ATQV;  // 2 Bp x n64: [A^T q ; A^T v]
  double* SC;    // Bp x SC_N
  int *active, *fail, *n_active;
  int* retry;         // instances whose factorisation is being re-attempted (failure ladder)
  double* delta_retry;  // their damping for the attempt
  double* M;
  double* Linv;
  const double* rowsum;  // A 1 (m): the cold-start primal residual of _indicators (:884)
};

struct DenseFwdArgs {
  DenseState s;
  int inst0;  // first instance of this chunk in the caller's arrays (c, b_batch, init, maps, outputs)
  double* c;  // B x c_stride (written when the user map packs it)
  long long c_stride;
  const double* b_batch;
  long long b_stride;
  const double* b_problem;
  const double* init;
  intopt::UserMap um;
  double* xhat;
  double* state;
  int* status;
  int* iters;
  double alpha0, tol, thr, damping, beta;
  int max_iter, plain;
};

__device__ __forceinline__ double b_at(const DenseFwdArgs& a, int inst, int i) {  // inst: chunk-local
  return a.b_batch ? a.b_batch[(long long)(a.inst0 + inst) * a.b_stride + i] : a.b_problem[i];
}

// _initialization (:647-661) + cold-start residual norms (:884,907-910)
__global__ void __launch_bounds__(256) dense_init_kernel(const DenseFwdArgs* __restrict__ ap) {
  const DenseFwdArgs& a = *ap;
  __shared__ double red[4 * 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const DenseState& s = a.s;
  const int n = s.n, m = s.m;
  double* c = a.c + (long long)(a.inst0 + b) * a.c_stride;
  if (a.um.in) {
    for (int k = tid; k < n; k += blockDim.x) {
      const int src = a.um.in_idx[k];
      c[k] = src >= 0 ? (a.um.in_scale ? a.um.in_scale[k] : 1.0) * intopt::um_load(a.um, a.inst0 + b, src) : 0.0;
    }
    __syncthreads();
  }
  const long long slen = 2LL * n + m + 2;
  const double* s0 = a.init ? a.init + (long long)(a.inst0 + b) * slen : nullptr;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k = tid; k < s.n64; k += blockDim.x) {
    const long long o = (long long)b * s.n64 + k;
    s.X[o] = (s0 && k < n) ? s0[k] : 1.0;
    s.T[o] = (s0 && k < n) ? s0[n + m + k] : 1.0;
    s.P[o] = 0.0;
    s.DX[o] = 0.0;
    s.DXDT[o] = 0.0;
    s.W[o] = 0.0;
    s.R2[o] = 0.0;
    s.D[o] = k < n ? 1.0 : 0.0;
    s.G[o] = 0.0;
    s.G[(long long)(s.Bp + b) * s.n64 + k] = 0.0;
    if (k < n) {
      const double ck = c[k], e = ck - 1.0;
      acc[0] = fma(e, e, acc[0]);
      acc[1] += ck;
      if (!(fabs(ck) < 1e300)) acc[2] += 1.0;
    }
  }
  for (int i = tid; i < s.m64; i += blockDim.x) {
    s.Y[(long long)b * s.m64 + i] = (s0 && i < m) ? s0[n + i] : 0.0;
    if (i < m) {
      const double e = b_at(a, b, i) - s.rowsum[i];
      acc[3] = fma(e, e, acc[3]);
    }
  }
  block_sum<4>(acc, red);
  if (tid == 0) {
    double* sc = s.SC + (long long)b * SC_N;
    sc[SC_TAU] = s0 ? s0[2 * n + m] : 1.0;
    sc[SC_KAPPA] = s0 ? s0[2 * n + m + 1] : 1.0;
    sc[SC_RD0] = fmax(1.0, sqrt(acc[0]));
    sc[SC_RG0] = fmax(1.0, fabs(1.0 + acc[1]));
    sc[SC_RP0] = fmax(1.0, sqrt(acc[3]));
    sc[SC_ONE] = 1.0;
    sc[SC_DTAU] = sc[SC_DKAPPA] = sc[SC_GAMMA] = 0.0;
    const int nonfinite = acc[2] > 0.0;
    a.status[a.inst0 + b] = nonfinite ? intopt::ST_NONFINITE : intopt::ST_OK;  // the reference raises LinAlgError here (:1332-1341)
    a.iters[a.inst0 + b] = 0;
    s.active[b] = nonfinite ? 0 : 1;
    s.fail[b] = 0;
    if (b == 0) s.n_active[0] = s.n_active[1] = 0;  // (instance count of the loop test, retry count of the ladder)
  }
}

// Loop top: residuals (:765-770), indicators and termination tests (:880-917, :1408-1423), then the right-hand-side
// inputs of the (p,q) solve and of the predictor: G = [D c ; D w].  AX = A x and ATY = A^T y come from the GEMMs.
__global__ void __launch_bounds__(256) dense_resid_kernel(const DenseFwdArgs* __restrict__ ap) {
  const DenseFwdArgs& a = *ap;
  __shared__ double red[5 * 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const DenseState& s = a.s;
  if (!s.active[b]) return;
  const int n = s.n, m = s.m;
  const long long on = (long long)b * s.n64, om = (long long)b * s.m64;
  double* sc = s.SC + (long long)b * SC_N;
  const double tau = sc[SC_TAU], kappa = sc[SC_KAPPA];
  const double* c = a.c + (long long)(a.inst0 + b) * a.c_stride;
  const int iters = a.iters[a.inst0 + b];
  const bool solved = s.fail[b] == 0;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = tid; k < n; k += blockDim.x) {
    const double xk = s.X[on + k], tk = s.T[on + k], ck = c[k];
    const double r = -(s.ATY[on + k] + tk - ck * tau);  // :766
    s.R2[on + k] = r;
    s.D[on + k] = xk / tk;  // :772
    acc[0] = fma(xk, tk, acc[0]);
    acc[1] = fma(ck, xk, acc[1]);
    acc[2] = fma(r, r, acc[2]);
  }
  for (int i = tid; i < m; i += blockDim.x) {
    const double bi = b_at(a, b, i);
    const double r = -(s.AX[om + i] - bi * tau);  // :765
    s.AX[om + i] = r;                             // AX now holds r1
    acc[3] = fma(r, r, acc[3]);
    acc[4] = fma(bi, s.Y[om + i], acc[4]);
  }
  block_sum<5>(acc, red);
  const double xt_ = acc[0], cx_ = acc[1], nd2 = acc[2], np2 = acc[3], by_ = acc[4];
  const double mu = (xt_ + tau * kappa) / (double)(n + 1);  // :770
  if (iters > 0) {
    const double rho_p = sqrt(np2) / sc[SC_RP0];
    const double rho_d = sqrt(nd2) / sc[SC_RD0];
    const double rho_A = fabs(cx_ - by_) / (tau + fabs(by_));
    const double rho_g = fabs(kappa + cx_ - by_) / sc[SC_RG0];
    const bool go = (rho_p > a.tol || rho_d > a.tol || rho_A > a.tol) && (mu > a.thr);
    const bool inf1 = rho_p < a.tol && rho_d < a.tol && rho_g < a.tol && tau < a.tol * fmax(1.0, kappa);
    const bool inf2 = mu < a.tol && tau < a.tol * fmin(1.0, kappa);
    int st = -1;
    if (inf1 || inf2)
      st = intopt::ST_INFEAS;
    else if (a.max_iter < iters)
      st = intopt::ST_MAXITER;
    else if (!solved)
      st = intopt::ST_BREAKDOWN;
    else if (!go)
      st = intopt::ST_OK;
    if (st >= 0) {
      if (tid == 0) {
        a.status[a.inst0 + b] = st;
        s.active[b] = 0;
      }
      return;
    }
  }
  const double gamma0 = a.plain ? a.beta * (xt_ / (double)n) : 0.0;  // :1392-1394
  const double eta0 = 1.0 - gamma0, gm0 = gamma0 * mu;
  for (int k = tid; k < n; k += blockDim.x) {
    const double xk = s.X[on + k], tk = s.T[on + k], dk = s.D[on + k];
    const double rxt = a.plain ? -(xk * tk - gm0) : -(xk * tk);                          // :788
    const double wk = (a.plain ? eta0 * s.R2[on + k] : s.R2[on + k]) - (1.0 / xk) * rxt;  // :801
    s.W[on + k] = wk;
    s.G[on + k] = dk * c[k];
    s.G[(long long)(s.Bp + b) * s.n64 + k] = dk * wk;
  }
  if (tid == 0) {
    a.iters[a.inst0 + b] = iters + 1;
    sc[SC_MU] = mu;
    sc[SC_R3] = -(-cx_ + by_ - kappa);  // :767
    sc[SC_GAMMA] = gamma0;
    sc[SC_ETA] = eta0;
    sc[SC_XT] = xt_;
    s.fail[b] = 0;
    atomicAdd(s.n_active, 1);
  }
}

// ratio test pieces shared by the predictor and the corrector (:839-844, :870-878)
struct StepScalars {
  double dtau, dkappa, minr;
};

// After the (p,q) + predictor solve: p, u, the scalars of :837-842, the predictor's ratio test, gamma (:844), and the
// corrector's right-hand-side input G = D w' (:792-801).  With pc=False (plain) the first direction is the step.
__global__ void __launch_bounds__(256) dense_predict_kernel(const DenseFwdArgs* __restrict__ ap) {
  const DenseFwdArgs& a = *ap;
  __shared__ double red[5 * 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const DenseState& s = a.s;
  if (!s.active[b]) return;
  if (s.fail[b]) return;  // unsolved (:855-857): the iterate stays, the tests at the loop top end the instance
  const int n = s.n, m = s.m;
  const long long on = (long long)b * s.n64, om = (long long)b * s.m64;
  const long long on2 = (long long)(s.Bp + b) * s.n64, om2 = (long long)(s.Bp + b) * s.m64;
  double* sc = s.SC + (long long)b * SC_N;
  const double tau = sc[SC_TAU], kappa = sc[SC_KAPPA], mu = sc[SC_MU], r3 = sc[SC_R3];
  const double* c = a.c + (long long)(a.inst0 + b) * a.c_stride;
  const double* q = s.RHS + om;
  const double* v = s.RHS + om2;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = tid; k < n; k += blockDim.x) {
    const double dk = s.D[on + k], ck = c[k];
    const double pk = dk * (s.ATQV[on + k] - ck);            // :678 for (c,b)
    const double uk = dk * (s.ATQV[on2 + k] - s.W[on + k]);  // :678 for the predictor
    s.P[on + k] = pk;
    s.DX[on + k] = uk;
    acc[0] = fma(ck, pk, acc[0]);
    acc[2] = fma(ck, uk, acc[2]);
    if (!(fabs(pk) < 1e300)) acc[4] += 1.0;
  }
  for (int i = tid; i < m; i += blockDim.x) {
    const double bi = b_at(a, b, i);
    acc[1] = fma(bi, q[i], acc[1]);
    acc[3] = fma(bi, v[i], acc[3]);
    if (!(fabs(q[i]) < 1e300)) acc[4] += 1.0;
  }
  block_sum<5>(acc, red);
  const double cp_ = acc[0], bq_ = acc[1], cu_ = acc[2], bv_ = acc[3];
  if (acc[4] > 0.0) {  // NaN in p or q (:805-806) -> unsolved
    if (tid == 0) s.fail[b] = 1;
    return;
  }
  const double gamma = sc[SC_GAMMA], eta = sc[SC_ETA];
  const double gm = gamma * mu;
  const double rtk = -(tau * kappa - gm);                                                                    // :789
  const double dtau = ((eta * r3 + 1.0 / tau * rtk - (-cu_ + bv_)) / (1.0 / tau * kappa + (-cp_ + bq_)));  // :837-838
  const double dkappa = (rtk - kappa * dtau) / tau;                                                          // :842
  double mn = INFINITY;
  for (int k = tid; k < n; k += blockDim.x) {
    const double xk = s.X[on + k], tk = s.T[on + k];
    const double rxt = -(xk * tk - gm);
    const double dxk = s.DX[on + k] + s.P[on + k] * dtau;  // :839
    const double rx = 1.0 / xk;
    const double dtk = rx * (rxt - tk * dxk);  // :841
    s.DX[on + k] = dxk;
    s.DXDT[on + k] = dxk * dtk;
    if (dxk < 0.0) mn = fmin(mn, dxk * rx);
    if (dtk < 0.0) mn = fmin(mn, dtk * (s.D[on + k] * rx));
  }
  mn = block_min(mn, red);
  const double minr = mn < 0.0 ? -1.0 / mn : INFINITY;
  double alpha = fmin(1.0, minr);  // _get_step with alpha0 = 1 (:843)
  if (dtau < 0.0) alpha = fmin(alpha, tau / -dtau);
  if (dkappa < 0.0) alpha = fmin(alpha, kappa / -dkappa);
  const double oma = 1.0 - alpha;
  const double gamma1 = oma * oma * fmin(0.1, oma);  // :844
  if (a.plain) {
    // (:846-847) the first direction is the step
    double al = fmin(1.0, a.alpha0 * minr);
    if (dtau < 0.0) al = fmin(al, a.alpha0 * tau / -dtau);
    if (dkappa < 0.0) al = fmin(al, a.alpha0 * kappa / -dkappa);
    for (int k = tid; k < n; k += blockDim.x) {
      const double xk = s.X[on + k], tk = s.T[on + k];
      const double rxt = -(xk * tk - gm);
      const double dxk = s.DX[on + k];
      const double dtk = (1.0 / xk) * (rxt - tk * dxk);
      s.X[on + k] = xk + al * dxk;
      s.T[on + k] = tk + al * dtk;
    }
    for (int i = tid; i < m; i += blockDim.x) s.Y[om + i] += al * (v[i] + q[i] * dtau);  // :840
    if (tid == 0) {
      sc[SC_TAU] = tau + al * dtau;
      sc[SC_KAPPA] = kappa + al * dkappa;
    }
    return;
  }
  const double eta1 = 1.0 - gamma1, gm1 = gamma1 * mu;
  for (int k = tid; k < n; k += blockDim.x) {
    const double xk = s.X[on + k], tk = s.T[on + k];
    const double rxt = -(xk * tk - gm1) - s.DXDT[on + k];           // :788,793
    const double wk = eta1 * s.R2[on + k] - (1.0 / xk) * rxt;        // :801
    s.W[on + k] = wk;
    s.G[on + k] = s.D[on + k] * wk;
  }
  if (tid == 0) {
    sc[SC_CP] = cp_;
    sc[SC_BQ] = bq_;
    sc[SC_DTAU] = dtau;
    sc[SC_DKAPPA] = dkappa;
    sc[SC_GAMMA] = gamma1;
    sc[SC_ETA] = eta1;
  }
}

// After the corrector solve: direction (:837-842 with the Mehrotra terms :792-795), ratio test with alpha0 and the
// step (:1402-1403, :861-868).
__global__ void __launch_bounds__(256) dense_step_kernel(const DenseFwdArgs* __restrict__ ap) {
  const DenseFwdArgs& a = *ap;
  __shared__ double red[2 * 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const DenseState& s = a.s;
  if (!s.active[b] || s.fail[b] || a.plain) return;
  const int n = s.n, m = s.m;
  const long long on = (long long)b * s.n64, om = (long long)b * s.m64;
  double* sc = s.SC + (long long)b * SC_N;
  const double tau = sc[SC_TAU], kappa = sc[SC_KAPPA], mu = sc[SC_MU], r3 = sc[SC_R3];
  const double* c = a.c + (long long)(a.inst0 + b) * a.c_stride;
  const double* q = s.RHS + om;
  const double* v = s.RHS + (long long)(s.Bp + b) * s.m64;  // corrector solution
  const double* atv = s.ATQV + (long long)(s.Bp + b) * s.n64;
  double acc[2] = {0.0, 0.0};
  for (int k = tid; k < n; k += blockDim.x) {
    const double uk = s.D[on + k] * (atv[k] - s.W[on + k]);  // :678
    s.DX[on + k] = uk;
    acc[0] = fma(c[k], uk, acc[0]);
  }
  for (int i = tid; i < m; i += blockDim.x) acc[1] = fma(b_at(a, b, i), v[i], acc[1]);
  block_sum<2>(acc, red);
  const double cu_ = acc[0], bv_ = acc[1], cp_ = sc[SC_CP], bq_ = sc[SC_BQ];
  const double gamma = sc[SC_GAMMA], eta = sc[SC_ETA], gm = gamma * mu;
  double rtk = -(tau * kappa - gm);
  rtk -= sc[SC_DTAU] * sc[SC_DKAPPA];  // :795
  const double dtau = ((eta * r3 + 1.0 / tau * rtk - (-cu_ + bv_)) / (1.0 / tau * kappa + (-cp_ + bq_)));
  const double dkappa = (rtk - kappa * dtau) / tau;
  double mn = INFINITY;
  for (int k = tid; k < n; k += blockDim.x) {
    const double xk = s.X[on + k], tk = s.T[on + k];
    const double rxt = -(xk * tk - gm) - s.DXDT[on + k];
    const double dxk = s.DX[on + k] + s.P[on + k] * dtau;
    const double rx = 1.0 / xk;
    const double dtk = rx * (rxt - tk * dxk);
    s.DX[on + k] = dxk;
    if (dxk < 0.0) mn = fmin(mn, dxk * rx);
    if (dtk < 0.0) mn = fmin(mn, dtk * (s.D[on + k] * rx));
  }
  mn = block_min(mn, red);
  const double minr = mn < 0.0 ? -1.0 / mn : INFINITY;
  double al = fmin(1.0, a.alpha0 * minr);
  if (dtau < 0.0) al = fmin(al, a.alpha0 * tau / -dtau);
  if (dkappa < 0.0) al = fmin(al, a.alpha0 * kappa / -dkappa);
  for (int k = tid; k < n; k += blockDim.x) {
    const double xk = s.X[on + k], tk = s.T[on + k];
    const double rxt = -(xk * tk - gm) - s.DXDT[on + k];
    const double dxk = s.DX[on + k];
    const double dtk = (1.0 / xk) * (rxt - tk * dxk);
    s.X[on + k] = xk + al * dxk;
    s.T[on + k] = tk + al * dtk;
  }
  for (int i = tid; i < m; i += blockDim.x) s.Y[om + i] += al * (v[i] + q[i] * dtau);  // d_y = v + q d_tau (:840)
  if (tid == 0) {
    sc[SC_TAU] = tau + al * dtau;
    sc[SC_KAPPA] = kappa + al * dkappa;
  }
}

// Epilogue (:1441-1447): x_hat = x / tau, the saved iterate, x in the caller's layout.
__global__ void __launch_bounds__(256) dense_epilogue_kernel(const DenseFwdArgs* __restrict__ ap) {
  const DenseFwdArgs& a = *ap;
  const int b = blockIdx.x, tid = threadIdx.x;
  const DenseState& s = a.s;
  const int n = s.n, m = s.m;
  const long long on = (long long)b * s.n64, om = (long long)b * s.m64, slen = 2LL * n + m + 2;
  const double* sc = s.SC + (long long)b * SC_N;
  const double tau = sc[SC_TAU];
  double* st = a.state + (long long)(a.inst0 + b) * slen;
  for (int k = tid; k < n; k += blockDim.x) {
    const double xk = s.X[on + k];
    if (a.xhat) a.xhat[(long long)(a.inst0 + b) * n + k] = xk / tau;
    st[k] = xk;
    st[n + m + k] = s.T[on + k];
  }
  for (int i = tid; i < m; i += blockDim.x) st[n + i] = s.Y[om + i];
  if (tid == 0) {
    st[2 * n + m] = tau;
    st[2 * n + m + 1] = sc[SC_KAPPA];
    if (s.active[b]) a.status[a.inst0 + b] = intopt::ST_MAXITER;  // still running when the host loop gave up
  }
  if (a.um.out) {
    for (int i = tid; i < a.um.n_out; i += blockDim.x) {
      const int k = a.um.out_idx[i];
      intopt::um_store(a.um, a.inst0 + b, i, k >= 0 ? s.X[on + k] / tau : (a.um.out_const ? a.um.out_const[i] : 0.0));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward (ip_model_whole.py:1515-1560, ip_model_alter.py:1524-1624).
// ---------------------------------------------------------------------------------------------------------
struct DenseBwdArgs {
  DenseState s;
  int inst0;
  int mode, nvar;
  double damping;
  const double* state;
  const double* c;
  long long c_stride;
  const double* b_batch;
  long long b_stride;
  const double* b_problem;
  double* gx;  // B x nvar (written when the user map packs it)
  intopt::UserMap um;
  double* gc;
  int* status;
};
enum { SB_TAU = 0, SB_MU, SB_XG };

__global__ void __launch_bounds__(256) dense_bwd_prep_kernel(const DenseBwdArgs a) {
  __shared__ double red[2 * 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const DenseState& s = a.s;
  const int n = s.n, m = s.m, nvar = a.nvar;
  const long long on = (long long)b * s.n64, slen = 2LL * n + m + 2;
  const double* st = a.state + (long long)(a.inst0 + b) * slen;
  const double* c = a.c + (long long)(a.inst0 + b) * a.c_stride;
  double* g = a.gx + (long long)(a.inst0 + b) * nvar;
  if (a.um.in) {
    for (int k = tid; k < nvar; k += blockDim.x) g[k] = intopt::um_load(a.um, a.inst0 + b, a.um.in_idx[k]);
    __syncthreads();
  }
  const double tau = st[2 * n + m], kappa = st[2 * n + m + 1];
  const bool whole = a.mode == 0;
  double acc[2] = {0.0, 0.0};
  for (int k = tid; k < s.n64; k += blockDim.x) {
    double xx = 0.0, g1 = 0.0, g2 = 0.0;
    if (k < n) {
      const double xk = whole ? st[k] : st[k] / tau;
      const double tk = whole ? st[n + m + k] : st[n + m + k] / tau;
      xx = xk * xk;
      acc[0] = fma(xk, tk, acc[0]);
      const double gk = k < nvar ? g[k] : 0.0;
      acc[1] = fma(xk, gk, acc[1]);
      g1 = xx * gk;
      g2 = whole ? xx * c[k] : 0.0;
    }
    s.D[on + k] = xx;
    s.G[(long long)(s.Bp + b) * s.n64 + k] = g1;  // second half -> RHS "v" row
    s.G[on + k] = g2;                             // first half  -> RHS "q" row
  }
  block_sum<2>(acc, red);
  if (tid == 0) {
    double* sc = s.SC + (long long)b * SC_N;
    sc[SB_TAU] = tau;
    sc[SB_MU] = whole ? (acc[0] + tau * kappa) / (double)(n + 1) : acc[0] / (double)n;  // :1517 | alter :1525
    sc[SB_XG] = acc[1];
    s.active[b] = 1;
    s.fail[b] = 0;
  }
}

__global__ void __launch_bounds__(256) dense_bwd_final_kernel(const DenseBwdArgs a) {
  __shared__ double red[4 * 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const DenseState& s = a.s;
  const int n = s.n, m = s.m, nvar = a.nvar;
  const long long on = (long long)b * s.n64, om = (long long)b * s.m64;
  const long long on2 = (long long)(s.Bp + b) * s.n64, om2 = (long long)(s.Bp + b) * s.m64;
  const double* sc = s.SC + (long long)b * SC_N;
  const double tau = sc[SB_TAU], mu = sc[SB_MU], xg = sc[SB_XG];
  const double* c = a.c + (long long)(a.inst0 + b) * a.c_stride;
  const double* g = a.gx + (long long)(a.inst0 + b) * nvar;
  double* gc = a.gc + (long long)(a.inst0 + b) * nvar;
  const bool whole = a.mode == 0;
  if (s.fail[b]) {  // the reference raises LinAlgError (:1546-1547); here: zero gradient + status
    for (int k = tid; k < nvar; k += blockDim.x) gc[k] = 0.0;
    if (tid == 0) a.status[a.inst0 + b] = intopt::ST_BREAKDOWN;
  } else if (!whole) {
    for (int k = tid; k < nvar; k += blockDim.x) gc[k] = -s.D[on + k] * (g[k] - s.ATQV[on2 + k]) / mu;
    if (tid == 0) a.status[a.inst0 + b] = intopt::ST_OK;
  } else {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const double* q = s.RHS + om;
    const double* v = s.RHS + om2;
    for (int k = tid; k < n; k += blockDim.x) {
      const double gk = k < nvar ? g[k] : 0.0, x2 = s.D[on + k];
      const double ugk = x2 * (s.ATQV[on2 + k] - tau * gk) / mu;  // :1549 applied to g
      const double pk = x2 * (s.ATQV[on + k] - c[k]) / mu;        // :1548
      s.DX[on + k] = ugk;
      s.P[on + k] = pk;
      acc[0] = fma(c[k], ugk, acc[0]);
      acc[1] = fma(c[k], pk, acc[1]);
    }
    for (int i = tid; i < m; i += blockDim.x) {
      const double bi = a.b_batch ? a.b_batch[(long long)(a.inst0 + b) * a.b_stride + i] : (a.b_problem ? a.b_problem[i] : 0.0);
      acc[2] = fma(bi, v[i], acc[2]);
      acc[3] = fma(bi, q[i], acc[3]);
    }
    block_sum<4>(acc, red);
    const double sfac = (xg + acc[0] - acc[2]) / (acc[3] - acc[1] + mu / (tau * tau));  // :1552-1553
    for (int k = tid; k < nvar; k += blockDim.x) gc[k] = s.DX[on + k] + s.P[on + k] * sfac;  // :1554-1556
    if (tid == 0) a.status[a.inst0 + b] = intopt::ST_OK;
  }
  if (a.um.out) {
    __syncthreads();
    for (int i = tid; i < a.um.n_out; i += blockDim.x) {
      const int k = a.um.out_idx[i];
      intopt::um_store(a.um, a.inst0 + b, i, k >= 0 ? gc[k] : 0.0);
    }
  }
}

// The failure ladder of _get_delta (ip_model_whole.py:798-822) on the lock-step path: an active instance whose Cholesky hit
// a non-positive pivot gets two more attempts with the diagonal raised by eps * dmax, eps = 1e-12 then 1e-8, before it is
// reported unsolved.  dmax is bounded by max_k d_k * max_i sum_k A_ik^2 (the largest diagonal entry of A diag(d) A^T can
// be no larger).  This kernel marks the instances of attempt `rung` (0 / 1) and sets their damping; the assembly and the
// factorisation then run again on exactly those (everyone else leaves at the mask test).
__global__ void __launch_bounds__(256) dense_retry_prep_kernel(DenseState s, double delta, double rowsq_max, int rung,
                                                               int* __restrict__ n_retry) {
  __shared__ double red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const bool again = s.active[b] && s.fail[b];
  if (!again) {
    if (tid == 0) s.retry[b] = 0;
    return;
  }
  if (tid == 0 && n_retry) atomicAdd(n_retry, 1);
  double mx = 0.0;
  for (int k = tid; k < s.n; k += blockDim.x) mx = fmin(mx, -s.D[(long long)b * s.n64 + k]);
  mx = -block_min(mx, red);
  if (tid == 0) {
    s.delta_retry[b] = delta + (rung == 0 ? 1e-12 : 1e-8) * mx * rowsq_max;
    s.retry[b] = 1;
    s.fail[b] = 0;
  }
}

// Graph form of the ladder: the retry pass of a rung is the body of an IF node that only runs when some instance failed.
__global__ void dense_if_cond_kernel(cudaGraphConditionalHandle handle, int* __restrict__ n_retry) {
  if (threadIdx.x == 0) {
    cudaGraphSetConditional(handle, *n_retry > 0 ? 1u : 0u);
    *n_retry = 0;
  }
}

// End of an iteration's termination tests: the while-node of the forward graph runs its body again iff some instance
// entered this iteration (dense_resid_kernel counted them).
__global__ void dense_loop_cond_kernel(cudaGraphConditionalHandle handle, int* __restrict__ n_active) {
  if (threadIdx.x == 0) {
    cudaGraphSetConditional(handle, *n_active > 0 ? 1u : 0u);
    *n_active = 0;  // for the next iteration's count (the body graph holds kernels only)
  }
}

// coefficient arrays of the backward right-hand sides (:1529-1530): v = tau * A(X2 g), q = mu * b + A(X2 c)
__global__ void dense_fill_kernel(double* p, long long n, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// Register-resident DMMA chains: the measured FP64 tensor-core roofline denominator.
__global__ void __launch_bounds__(256) dense_dmma_peak_kernel(double* out, int iters, double seed) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = seed + i;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace dn
