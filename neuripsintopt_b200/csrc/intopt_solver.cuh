// Per-instance interior-point solver: one CTA owns one LP for the whole IPM loop.
//
// Data placement (Energy LP, m=172, n=1582): the factor of M = A.diag(d).A^T + delta.I lives in shared
// memory as an 8x8 block skyline (~60-80 KB), together with the m-vectors and one n-vector gather buffer;
// the element-wise n-vectors live in a per-CTA global workspace that stays L2 resident (CTAs are
// persistent and pull instances from an atomic counter).  A (by rows / by interleaved column groups) and
// the gather lists that assemble M are shared by every instance and stay in L2.
//
// The code is written against a tiny "CTA execution" vocabulary (CTA_FOR / WARP_ROWS / cta_sum ...)
// so the same source also compiles as plain single-threaded C++ (-DINTOPT_HOST_EMU) for the CPU
// unit tests of the arithmetic; that build is test-only and is never loaded by the package.
//
// Reference arithmetic being reproduced (JayMan91/NeurIPSIntopt):
//   forward  Interior/ip_model_whole.py:1385-1447 (_get_delta :761-859, _get_step :870-878,
//            _do_step :861-868, _indicators :880-917)
//   backward Interior/ip_model_whole.py:1502-1560, Interior/ip_model_alter.py:1516-1624
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef INTOPT_HOST_EMU
inline int blockIdx_x() { return 0; }
#define IO_DEV inline
#define IO_HD inline
#define IO_FORCEINLINE inline
struct CtaCtx {
  static constexpr int tid = 0, nthreads = 1, lane = 0, warp = 0, nwarps = 1, warpsize = 1;
  double* red;
};
#define CTA_SYNC() ((void)0)
#define WARP_SYNC() ((void)0)
IO_DEV double warp_sum(double v) { return v; }
IO_DEV double warp_min(double v) { return v; }
#else
__device__ __forceinline__ int blockIdx_x() { return (int)blockIdx.x; }
#define IO_DEV __device__ __forceinline__
#define IO_HD __host__ __device__ inline
#define IO_FORCEINLINE __device__ __forceinline__
struct CtaCtx {
  int tid, nthreads, lane, warp, nwarps;
  static constexpr int warpsize = 32;
  double* red;  // 8*32 doubles of shared scratch
  // Barrier of this execution group: id 0 = the whole CTA (__syncthreads); a warp-specialised phase runs two
  // sub-groups side by side, each with its own named barrier (bar.sync id, threads).
  int bar_id, bar_n;
};
__device__ __forceinline__ void group_barrier(const CtaCtx& cx) {
  if (cx.bar_id == 0)
    __syncthreads();
  else if (cx.bar_id == 1)  // (immediate barrier ids, so that the assembler counts the barriers the kernel uses)
    asm volatile("bar.sync 1, %0;" ::"r"(cx.bar_n) : "memory");
  else
    asm volatile("bar.sync 2, %0;" ::"r"(cx.bar_n) : "memory");
}
#define CTA_SYNC() group_barrier(cx)
#define WARP_SYNC() __syncwarp()
IO_FORCEINLINE double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
IO_FORCEINLINE double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

#ifdef INTOPT_HOST_EMU
struct alignas(16) D2 {
  double x, y;
};
#else
typedef double2 D2;
#endif
#ifdef INTOPT_HOST_EMU
struct alignas(16) U4 {
  uint32_t x, y, z, w;
};
#else
typedef uint4 U4;
#endif

// loops: every thread of the CTA strides over [0,n)
#define CTA_FOR(i, n) for (int i = cx.tid; i < (n); i += cx.nthreads)
// the same index space in batches of NU strided elements per thread: a batch first loads everything it needs
// from the (L2-resident) global n-vectors into registers, then computes, so a batch costs one L2 round trip
constexpr int NU = 1;   // plain element-wise loops
constexpr int NUC = 1;  // loops that also run a column gather (register pressure)
#define CTA_BATCH_FOR(k0, n, U) for (int k0 = cx.tid; k0 < (n); k0 += (U) * cx.nthreads)
#define BATCH_EACH(u, k, k0, n, U)     \
  _Pragma("unroll") for (int u = 0; u < (U); ++u) if (const int k = (k0) + u * cx.nthreads; k < (n))
// one warp per item, lanes stride over the item's entries
#define WARP_ROWS(i, m) for (int i = cx.warp; i < (m); i += cx.nwarps)
#define LANE_FOR(e, lo, hi) for (int e = (lo) + cx.lane; e < (hi); e += cx.warpsize)

namespace intopt {

enum { ST_OK = 0, ST_MAXITER = 1, ST_INFEAS = 2, ST_BREAKDOWN = 3, ST_NONFINITE = 4 };
constexpr int KB = 8;  // block edge, == NB of intopt_structure.hpp
constexpr int S_DSCR = KB + KB * KB;  // one diagonal-block scratch: reciprocal pivots + the factored block (row-major)

// The gather buffer is indexed through gpad(): one unused slot after every 32 elements, so that run starts a
// multiple of 16 or 32 elements apart (the Energy LP: 144 = one task's variables) no longer share a bank.
// Every gather index list built on the host (intopt_structure.hpp) is in padded coordinates.
IO_HD int gpad(int k) { return k + (k >> 5); }
// Store element k of a gather vector; the pad slot after every 32nd element is kept at zero (the run-length row
// kernel sums straight across it, and the assembly's prefix sums overwrite the whole region in between).
IO_HD void gstore(double* G, int k, double v) {
  const int p = gpad(k);
  G[p] = v;
  if ((k & 31) == 31) G[p + 1] = 0.0;
}

struct ProblemDev {
  int m, n, nnz, mb, m8, mv, n_act;  // mv = m8 + 8: length of every m-vector (slot m8 is always zero)
  const int* g_csr_ptr;
  const int* csr_col;
  const double* csr_val;  // null if a_unit
  const int* g_csc_gptr;
  const uint16_t* csc_row;
  const double* csc_val;  // null if a_unit
  int a_unit;
  const uint32_t* rs_hdr;  // run-length rows (unit A): slot-packed headers, null if unavailable
  const uint16_t* rs_row;
  int rs_slots;
  const int* perm;  // internal row -> original row
  const int* g_fb;
  const int* g_boff;
  int bsize;
  const int* dummy_pos;
  int n_dummy;
  const int* g_act_ptr;
  const int* g_act_blk;
  int meta_ints;  // total ints of {fb, boff, act_ptr, act_blk, csr_ptr, csc_gptr, asm_gptr}: staged in smem if small
  int n_asm;
  const int* g_asm_gptr;
  const int* g_asm_rptr;
  const uint32_t* asm_run;  // run-length gather lists (unit weights only), null otherwise
  const int* asm_pos;
  const unsigned char* asm_diag;
  const int* asm_k;          // 32-bit gather indices, or
  const uint16_t* asm_k16;   // 16-bit ones when n < 65535 (asm_k is then null)
  const double* asm_w;  // null if asm_unit
  // per-instance values of A (same pattern): positions in the caller's CSR value array, -1 = padding (see
  // intopt_structure.hpp); null unless the problem was created for it
  int inst_values, n_csc_pad, n_asm_pad;  // (padded lengths of csc_val / asm_w)
  const int* csr_src;
  const int* csc_src;
  const int* asm_sa;
  const int* asm_sb;
  int asm_unit;
};

// The small schedule / pointer arrays that sit on the critical path of every blocked step.  They point into
// shared memory when staged (stage_meta) and into global memory otherwise.
struct Meta {
  const int* fb;
  const int* boff;
  const int* act_ptr;
  const int* act_blk;
  const int* csr_ptr;
  const int* csc_gptr;
  const int* asm_gptr;
  const int* asm_rptr;
  // value streams of A (by rows, by column groups, assembly weights): the problem's own, or this instance's copies
  // in the CTA's workspace when the values differ per instance (materialize_instance_values)
  const double* csr_val;
  const double* csc_val;
  const double* asm_w;
};

// Sum N partial values over the CTA; every thread receives the totals (fixed order => deterministic).
template <int N>
IO_DEV void cta_sum(const CtaCtx& cx, double (&v)[N]) {
#ifdef INTOPT_HOST_EMU
  (void)cx;
  (void)v;
#else
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  if (cx.lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) cx.red[k * 32 + cx.warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double s = 0.0;
    for (int w = 0; w < cx.nwarps; ++w) s += cx.red[k * 32 + w];
    v[k] = s;
  }
  __syncthreads();
#endif
}

template <int N>
IO_DEV void cta_min(const CtaCtx& cx, double (&v)[N]) {
#ifdef INTOPT_HOST_EMU
  (void)cx;
  (void)v;
#else
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_min(v[k]);
  if (cx.lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) cx.red[k * 32 + cx.warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double s = cx.red[k * 32];
    for (int w = 1; w < cx.nwarps; ++w) s = fmin(s, cx.red[k * 32 + w]);
    v[k] = s;
  }
  __syncthreads();
#endif
}

// ---------------------------------------------------------------------------------------------
// K1: M = A diag(d) A^T + delta I, assembled entry by entry into the block skyline L.
// `g` is the gather buffer holding d (g[n] == 0 is the padding slot).  Each warp walks whole groups
// of 32 entries whose term lists are interleaved, so index loads are fully coalesced.
// ---------------------------------------------------------------------------------------------
template <class IT>
IO_FORCEINLINE double gather_sum_unit(const IT* __restrict__ kk, const double* g, int len) {
  double s0 = 0.0, s1 = 0.0;
  int t = 0;
  for (; t + 8 <= len; t += 8) {  // eight independent index loads in flight per lane
    int k[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) k[u] = kk[(t + u) * 32];
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
      s0 += g[k[u]];
      s1 += g[k[u + 1]];
    }
  }
  for (; t < len; ++t) s0 += g[kk[t * 32]];
  return s0 + s1;
}

template <class IT>
IO_FORCEINLINE double gather_sum_weighted(const IT* __restrict__ kk, const double* __restrict__ ww, const double* g,
                                          int len) {
  double s0 = 0.0, s1 = 0.0;
  int t = 0;
  for (; t + 4 <= len; t += 4) {
    int k[4];
    double w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      k[u] = kk[(t + u) * 32];
      w[u] = ww[(t + u) * 32];
    }
    s0 = fma(w[0], g[k[0]], s0);
    s1 = fma(w[1], g[k[1]], s1);
    s0 = fma(w[2], g[k[2]], s0);
    s1 = fma(w[3], g[k[3]], s1);
  }
  for (; t < len; ++t) s0 = fma(ww[t * 32], g[kk[t * 32]], s0);
  return s0 + s1;
}

// ---------------------------------------------------------------------------------------------
// Double-double helpers for the prefix-sum assembly.  A run of consecutive gather indices sums to
// P[start+len] - P[start]; keeping the prefix sums P as unevaluated pairs hi+lo (about 106 bits) makes that
// difference exact to ~1e-32 * P[n], i.e. far below one ulp of any run sum, however wide the dynamic range
// of d is (d = x/t spans 1e-12..1e12 near convergence).
// ---------------------------------------------------------------------------------------------
IO_FORCEINLINE void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
IO_FORCEINLINE D2 dd_add_d(D2 a, double b) {  // (hi,lo) + double
  double s, e;
  two_sum(a.x, b, s, e);
  e += a.y;
  D2 r;
  r.x = s + e;
  r.y = e - (r.x - s);
  return r;
}
IO_FORCEINLINE D2 dd_add(D2 a, D2 b) {  // (hi,lo) + (hi,lo)
  double s, e;
  two_sum(a.x, b.x, s, e);
  e += a.y + b.y;
  D2 r;
  r.x = s + e;
  r.y = e - (r.x - s);
  return r;
}
IO_FORCEINLINE double dd_diff(D2 hi, D2 lo) {  // hi - lo rounded to double
  double s, e;
  two_sum(hi.x, -lo.x, s, e);
  return s + (e + (hi.y - lo.y));
}

// Exclusive double-double prefix sums of d[0..n): PP[gpad(k)] = sum_{k'<k} d[k'] for k = 0..n.  Chunks of 8
// consecutive elements per thread, chunk totals scanned in two more levels.  `tot` holds ceil(n/8) + 40 D2.
IO_DEV void dd_prefix_scan(const CtaCtx& cx, const double* __restrict__ d, int n, D2* PP, D2* tot) {
  const int nch = (n + 7) >> 3, ngr = (nch + 31) >> 5;
  D2* tot2 = tot + nch;
  CTA_FOR(t, nch) {
    D2 acc;
    acc.x = 0.0;
    acc.y = 0.0;
    const int k0 = t << 3, k1 = k0 + 8 < n ? k0 + 8 : n;
    for (int k = k0; k < k1; ++k) {
      PP[gpad(k)] = acc;  // exclusive, local to the chunk
      acc = dd_add_d(acc, d[k]);
    }
    tot[t] = acc;
  }
  CTA_SYNC();
  CTA_FOR(gq, ngr) {  // exclusive scan of 32 chunk totals each
    D2 acc;
    acc.x = 0.0;
    acc.y = 0.0;
    const int t0 = gq << 5, t1 = t0 + 32 < nch ? t0 + 32 : nch;
    for (int t = t0; t < t1; ++t) {
      const D2 v = tot[t];
      tot[t] = acc;
      acc = dd_add(acc, v);
    }
    tot2[gq] = acc;
  }
  CTA_SYNC();
  if (cx.tid == 0) {  // exclusive scan of the group totals (a handful), grand total to PP[n]
    D2 acc;
    acc.x = 0.0;
    acc.y = 0.0;
    for (int gq = 0; gq < ngr; ++gq) {
      const D2 v = tot2[gq];
      tot2[gq] = acc;
      acc = dd_add(acc, v);
    }
    PP[gpad(n)] = acc;
  }
  CTA_SYNC();
  CTA_FOR(t, nch) {
    const D2 off = dd_add(tot2[t >> 5], tot[t]);
    const int k0 = t << 3, k1 = k0 + 8 < n ? k0 + 8 : n;
    for (int k = k0; k < k1; ++k) PP[gpad(k)] = dd_add(off, PP[gpad(k)]);
  }
  CTA_SYNC();
}

// 4*NQ run headers of one matrix entry (NQ 16-byte loads in flight together): every run adds the difference of two
// double-double prefix sums; padding headers (len 0) contribute exactly 0
template <int NQ>
IO_FORCEINLINE void asm_run_batch(const uint32_t* __restrict__ hp, const D2* PP, double& s0, double& s1) {
  uint32_t h[4 * NQ];
#pragma unroll
  for (int c = 0; c < NQ; ++c) {
    const U4 q = *reinterpret_cast<const U4*>(hp + 4 * c);
    h[4 * c] = q.x;
    h[4 * c + 1] = q.y;
    h[4 * c + 2] = q.z;
    h[4 * c + 3] = q.w;
  }
#pragma unroll
  for (int u = 0; u < 4 * NQ; u += 2) {
    const int a0 = (int)(h[u] & 0xffffu), a1 = (int)(h[u + 1] & 0xffffu);
    const D2 lo0 = PP[a0], hi0 = PP[a0 + (int)(h[u] >> 16)];
    const D2 lo1 = PP[a1], hi1 = PP[a1 + (int)(h[u + 1] >> 16)];
    s0 += dd_diff(hi0, lo0);
    s1 += dd_diff(hi1, lo1);
  }
}

// ---------------------------------------------------------------------------------------------
// K1: M = A diag(d) A^T + delta I, assembled entry by entry into the block skyline L.
// `d` is the (global) weight vector, `region` the shared gather region of 2*gather_len(n) doubles.
//  * unit-weight problems (0/1 matrices): the term list of an entry is a few runs of consecutive indices;
//    a run costs two 16-byte gathers from the double-double prefix sums of d instead of one gather per term;
//  * general problems: one gather per term from a copy of d in `region`, indices warp-interleaved.
// ---------------------------------------------------------------------------------------------
IO_DEV void assemble_normal_matrix(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* __restrict__ d,
                                   double* region, D2* scan_tmp, double delta, double* L, int dbg = 0) {
  const int n = P.n;
#ifdef INTOPT_PROFILE
  const int ngrp = (dbg & 1) ? 0 : (P.n_asm >> 5);
#else
  (void)dbg;
  const int ngrp = P.n_asm >> 5;
#endif
  if (P.asm_run) {
    D2* PP = reinterpret_cast<D2*>(region);
    dd_prefix_scan(cx, d, n, PP, scan_tmp);  // scan_tmp may alias L (scan_in_factor): the zero fill comes after
    CTA_FOR(e, P.bsize) L[e] = 0.0;
    CTA_SYNC();
    CTA_FOR(e, P.n_dummy) L[P.dummy_pos[e]] = 1.0;
    // Lane l of a warp owns entry 32*grp + l; its R run headers (R % 4 == 0, the same for the whole group: entries are
    // sorted by run count) are contiguous and fetched with up to three 16-byte loads that are all in flight
    // together; groups of 4 / 8 headers take the shorter batches, so padding headers cost nothing beyond the
    // rounding of R to a multiple of four.
#ifndef INTOPT_HOST_EMU
    // A group costs one L2 round trip for its headers (and one for its target position) before any arithmetic can
    // start, and a warp walks ~36 groups: the first four headers, the position and the diagonal flag of the warp's NEXT
    // group are requested before the current group is summed (same sums, same order: bit-identical).
    {
      struct Pre {
        int r0, R, pos, dg;
        U4 h;
      };
      auto fetch = [&](int g) {
        Pre f;
        f.r0 = M.asm_rptr[g];
        f.R = M.asm_rptr[g + 1] - f.r0;
        f.h.x = f.h.y = f.h.z = f.h.w = 0u;  // (len 0 runs contribute exactly 0)
        if (f.R > 0) f.h = *reinterpret_cast<const U4*>(P.asm_run + (size_t)f.r0 * 32 + (size_t)cx.lane * f.R);
        const int e = g * 32 + cx.lane;
        f.pos = P.asm_pos[e];
        f.dg = P.asm_diag[e];
        return f;
      };
      int grp = cx.warp;
      Pre cur{};
      if (grp < ngrp) cur = fetch(grp);
      while (grp < ngrp) {
        const int nxt = grp + cx.nwarps;
        Pre nx = cur;
        if (nxt < ngrp) nx = fetch(nxt);
        double s0 = 0.0, s1 = 0.0;
        {
          const uint32_t h[4] = {cur.h.x, cur.h.y, cur.h.z, cur.h.w};
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const int a0 = (int)(h[u] & 0xffffu), a1 = (int)(h[u + 1] & 0xffffu);
            const D2 lo0 = PP[a0], hi0 = PP[a0 + (int)(h[u] >> 16)];
            const D2 lo1 = PP[a1], hi1 = PP[a1 + (int)(h[u + 1] >> 16)];
            s0 += dd_diff(hi0, lo0);
            s1 += dd_diff(hi1, lo1);
          }
        }
        const uint32_t* __restrict__ hp = P.asm_run + (size_t)cur.r0 * 32 + (size_t)cx.lane * cur.R;
        for (int rb = 4; rb < cur.R;) {
          const int left = cur.R - rb;
          if (left >= 12) {
            asm_run_batch<3>(hp + rb, PP, s0, s1);
            rb += 12;
          } else if (left >= 8) {
            asm_run_batch<2>(hp + rb, PP, s0, s1);
            rb += 8;
          } else {
            asm_run_batch<1>(hp + rb, PP, s0, s1);
            rb += 4;
          }
        }
        if (cur.pos >= 0) L[cur.pos] = cur.dg ? (s0 + s1) + delta : (s0 + s1);
        cur = nx;
        grp = nxt;
      }
    }
#else
    WARP_ROWS(grp, ngrp) {
      const int r0 = M.asm_rptr[grp], R = M.asm_rptr[grp + 1] - r0;
      LANE_FOR(l, 0, 32) {
        const uint32_t* __restrict__ hp = P.asm_run + (size_t)r0 * 32 + (size_t)l * R;
        double s0 = 0.0, s1 = 0.0;
        for (int rb = 0; rb < R;) {
          const int left = R - rb;
          if (left >= 12) {
            asm_run_batch<3>(hp + rb, PP, s0, s1);
            rb += 12;
          } else if (left >= 8) {
            asm_run_batch<2>(hp + rb, PP, s0, s1);
            rb += 8;
          } else {
            asm_run_batch<1>(hp + rb, PP, s0, s1);
            rb += 4;
          }
        }
        const int e = grp * 32 + l, pos = P.asm_pos[e];
        if (pos >= 0) L[pos] = P.asm_diag[e] ? (s0 + s1) + delta : (s0 + s1);
      }
    }
#endif
    CTA_SYNC();
    return;
  }
  CTA_FOR(e, P.bsize) L[e] = 0.0;
  CTA_FOR(k, n) region[gpad(k)] = d[k];
  if (cx.tid == 0) region[gpad(n)] = 0.0;
  CTA_SYNC();
  CTA_FOR(e, P.n_dummy) L[P.dummy_pos[e]] = 1.0;
  const double* g = region;
  WARP_ROWS(grp, ngrp) {
    const int g0 = M.asm_gptr[grp], len = M.asm_gptr[grp + 1] - g0;
    LANE_FOR(l, 0, 32) {
      const int e = grp * 32 + l;
      const size_t base = (size_t)g0 * 32 + l;
      double s;
      if (P.asm_k16) {
        s = P.asm_unit ? gather_sum_unit(P.asm_k16 + base, g, len)
                       : gather_sum_weighted(P.asm_k16 + base, M.asm_w + base, g, len);
      } else {
        s = P.asm_unit ? gather_sum_unit(P.asm_k + base, g, len)
                       : gather_sum_weighted(P.asm_k + base, M.asm_w + base, g, len);
      }
      const int pos = P.asm_pos[e];
      if (pos >= 0) L[pos] = P.asm_diag[e] ? s + delta : s;
    }
  }
  CTA_SYNC();
}

// Optional per-phase cycle accounting (debug aid for profiles/): thread 0 of every CTA adds the cycles since
// the previous mark to bucket `k` of a global array.  Disabled (null pointer) in normal runs.
struct PhaseTimer {
  unsigned long long* buckets;
  long long last;
};
#if defined(INTOPT_HOST_EMU) || !defined(INTOPT_PROFILE)  // the product build carries no probes
#define PHASE_MARK(pt, k) ((void)0)
#define PHASE_START(pt) ((void)0)
#else
#define PHASE_START(pt)                                  \
  if ((pt).buckets && cx.tid == 0) (pt).last = clock64()
#define PHASE_MARK(pt, k)                                                   \
  if ((pt).buckets && cx.tid == 0) {                                        \
    const long long now_ = clock64();                                       \
    atomicAdd((pt).buckets + (k), (unsigned long long)(now_ - (pt).last));  \
    (pt).last = now_;                                                       \
  }
#endif

// ---------------------------------------------------------------------------------------------
// 8x8 block storage.  Row-major, but the four 16-byte chunks of row r are XOR-swizzled by (r>>1)&3, so a
// 128-bit access to the same logical chunk of 8 different rows touches 8 different bank groups
// (conflict-free), as does a 128-bit access to the four chunks of one row.
// ---------------------------------------------------------------------------------------------
IO_HD int bswz(int r) { return (r >> 1) & 3; }
IO_HD int bidx(int r, int k) { return r * 8 + ((((k >> 1) ^ ((r >> 1) & 3))) << 1) + (k & 1); }

IO_FORCEINLINE void load_row(const double* blk, int r, double (&v)[KB]) {
  const D2* p = reinterpret_cast<const D2*>(blk + r * KB);
  const int sw = bswz(r);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const D2 t = p[c ^ sw];
    v[2 * c] = t.x;
    v[2 * c + 1] = t.y;
  }
}
IO_FORCEINLINE void store_row(double* blk, int r, const double (&v)[KB]) {
  D2* p = reinterpret_cast<D2*>(blk + r * KB);
  const int sw = bswz(r);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    D2 t;
    t.x = v[2 * c];
    t.y = v[2 * c + 1];
    p[c ^ sw] = t;
  }
}
IO_FORCEINLINE void load_vec8(const double* v8, double (&v)[KB]) {
  const D2* p = reinterpret_cast<const D2*>(v8);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const D2 t = p[c];
    v[2 * c] = t.x;
    v[2 * c + 1] = t.y;
  }
}

// Cholesky of one 8x8 diagonal block (lower part), then its inverse written over it: on exit the block
// holds Linv = L_KK^-1 (lower).  Executed by ONE warp.  `scr` is 8 + 64 doubles of scratch.
// A non-positive / non-finite pivot sets *fail and continues with pivot 1.
#ifdef INTOPT_HOST_EMU
// single-threaded test build: plain column Cholesky + forward substitution
IO_DEV void diag_block_factor_invert(const CtaCtx& cx, double* blk, double* scr, int* fail,
                                     const double* panel = nullptr) {
  double* col = scr;
  double* inv = scr + KB;
  if (panel)  // fused look-ahead update: C -= P P^T with P = L_(K+1,K)
    for (int i = 0; i < KB; ++i)
      for (int k = 0; k <= i; ++k) {
        double s = 0.0;
        for (int j = 0; j < KB; ++j) s = fma(panel[bidx(i, j)], panel[bidx(k, j)], s);
        blk[bidx(i, k)] -= s;
      }
  for (int j = 0; j < KB; ++j) {
    LANE_FOR(i, j, KB) {
      double s = blk[bidx(i, j)];
      for (int k = 0; k < j; ++k) s = fma(-blk[bidx(i, k)], blk[bidx(j, k)], s);
      col[i] = s;
    }
    double piv = col[j];
    if (!(piv > 0.0) || !(piv < 1e300)) {
      *fail = 1;
      piv = 1.0;
    }
    const double dj = sqrt(piv);
    const double rinv = 1.0 / dj;
    LANE_FOR(i, j, KB) blk[bidx(i, j)] = (i == j) ? dj : col[i] * rinv;
  }
  LANE_FOR(c, 0, KB) {
    double x[KB];
    for (int i = 0; i < KB; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s = fma(-blk[bidx(i, k)], x[k], s);
      x[i] = (i >= c) ? s / blk[bidx(i, i)] : 0.0;
    }
    for (int i = 0; i < KB; ++i) inv[i * KB + c] = x[i];
  }
  for (int i = 0; i < KB; ++i)
    for (int k = 0; k <= i; ++k) blk[bidx(i, k)] = inv[i * KB + k];
}
#else
// CUDA build: the block lives in registers (lane i & 7 owns row i, lanes 8..31 mirror lanes 0..7), pivots
// and columns travel by warp shuffles, pivots use rsqrt so the dependent chain has no fp64 divide/sqrt.
// Factor only: L (row-major rows) and the reciprocal pivots are left in `scr` (KB + KB*KB doubles: rinv, then L) for
// whoever forms the inverse - this warp right away (diag_block_factor_invert) or another warp off the pivot chain
// (diag_block_invert_published).  `blk` itself is not written.
IO_DEV void diag_block_factor_publish(const CtaCtx& cx, const double* blk, double* scr, int* fail,
                                      const double* panel = nullptr) {
  const unsigned full = 0xffffffffu;
  const int i = cx.lane & 7;
  double a[KB], rinv[KB];
  load_row(blk, i, a);
  if (panel) {
    // fused look-ahead update in registers: row i of C -= P_i . P_k^T with P = L_(K+1,K).  The four lane octets (which
    // all hold row i) take two of the eight dot products each and exchange the results by shuffle: a quarter of the
    // shared-memory loads and FMAs on this warp's critical path, the same sums bit for bit.
    const int o = cx.lane >> 3;
    double own[KB], p0[KB], p1[KB];
    load_row(panel, i, own);
    load_row(panel, 2 * o, p0);
    load_row(panel, 2 * o + 1, p1);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      s0 = fma(own[j], p0[j], s0);
      s1 = fma(own[j], p1[j], s1);
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) a[k] -= __shfl_sync(full, (k & 1) ? s1 : s0, (k >> 1) * 8 + i);
  }
  bool bad = false;
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    double piv = __shfl_sync(full, a[j], j);
    // the entries of column j travel BEFORE they are scaled (lane k's a[j] is C_kj): the shuffles then overlap the
    // rsqrt instead of following it on the dependent chain; C_kj * r is formed locally, the same product lane k forms
    double ckj[KB];
#pragma unroll
    for (int k = j + 1; k < KB; ++k) ckj[k] = __shfl_sync(full, a[j], k);
    if (!(piv > 0.0) || !(piv < 1e300)) {
      bad = true;
      piv = 1.0;
    }
    const double r = rsqrt(piv);
    rinv[j] = r;
    a[j] = (i == j) ? piv * r : a[j] * r;  // column j of L (rows above the diagonal hold junk)
#pragma unroll
    for (int k = j + 1; k < KB; ++k) a[k] = fma(-a[j], ckj[k] * r, a[k]);
  }
  if (bad && cx.lane == 0) *fail = 1;
  // publish L (row i, plain row-major) so that every lane can read the rows below its own column
  double* Ls = scr + KB;
  if (cx.lane < KB) {
#pragma unroll
    for (int k = 0; k < KB; ++k) Ls[i * KB + k] = a[k];
  }
  if (cx.lane == KB) {  // (every lane holds the same reciprocal pivots)
#pragma unroll
    for (int k = 0; k < KB; ++k) scr[k] = rinv[k];
  }
  __syncwarp();
}

// Column i = lane & 7 of Linv from a published factor (see above), written over the diagonal block.  One warp.
IO_DEV void diag_block_invert_published(const CtaCtx& cx, double* blk, const double* scr) {
  const int i = cx.lane & 7;
  const double* Ls = scr + KB;
  double rinv[KB];
#pragma unroll
  for (int k = 0; k < KB; ++k) rinv[k] = scr[k];
  // column c = i of Linv by forward substitution with reciprocal pivots: x_c = rinv_c,
  // x_r = -rinv_r * sum_{k=c}^{r-1} L_rk x_k
  double x[KB];
#pragma unroll
  for (int r = 0; r < KB; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k < r) s = fma(Ls[r * KB + k], (k >= i) ? x[k] : 0.0, s);
    x[r] = (r == i) ? rinv[r] : ((r > i) ? -rinv[r] * s : 0.0);
  }
  __syncwarp();
  if (cx.lane < KB) {
#pragma unroll
    for (int r = 0; r < KB; ++r)
      if (r >= i) blk[bidx(r, i)] = x[r];
  }
  __syncwarp();
}

IO_DEV void diag_block_factor_invert(const CtaCtx& cx, double* blk, double* scr, int* fail,
                                     const double* panel = nullptr) {
  diag_block_factor_publish(cx, blk, scr, fail, panel);
  diag_block_invert_published(cx, blk, scr);
}

// One row of the panel block the pivot chain needs, L_(K+1,K) = C L_KK^-T, by substitution against the PUBLISHED factor of
// block K (its inverse is being formed by another warp at the same time): out_c = (c_c - sum_{k<c} out_k L_ck) / L_cc.
IO_FORCEINLINE void chol_panel_row_subst(double* blk, const double* scr, int row) {
  const double* Ls = scr + KB;
  double c[KB];
  load_row(blk, row, c);
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    const double xj = c[j] * scr[j];
    c[j] = xj;
#pragma unroll
    for (int k = j + 1; k < KB; ++k) c[k] = fma(-xj, Ls[k * KB + j], c[k]);
  }
  store_row(blk, row, c);
}
#endif

// ---------------------------------------------------------------------------------------------
// K2: blocked right-looking Cholesky on the block skyline with one-step look-ahead.
// Step K:  (a) panel: L_IK = C_IK * Linv_KK^T for the active block rows I > K (one thread per row);
//          (b) trailing update C_IJ -= L_IK L_JK^T, one 4x4 register tile per thread; warp 0 takes the
//              update of the next diagonal block first and then factors + inverts it while the other
//              warps finish the remaining block pairs.
// On exit: strictly-lower blocks hold L, diagonal blocks hold Linv_KK (what the solves need).
// Returns false (for every thread) on a bad pivot (the reference's cho_factor failure,
// ip_model_whole.py:742,755-758).
// ---------------------------------------------------------------------------------------------
// index of the row of the lower triangle that holds flat pair index p (pairs (a,b), b <= a, row-major)
IO_FORCEINLINE int tri_row(int p) {
  int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while ((a + 1) * (a + 2) / 2 <= p) ++a;
  while (a * (a + 1) / 2 > p) --a;
  return a;
}

IO_DEV void chol_update_task(const ProblemDev& P, const Meta& M, double* L, int K, int a0, int task) {
  // task = (pair of active block rows (I >= J), tile row tr, tile col tc): a 4x4 tile of block (I,J)
  (void)P;
  const int pi = task >> 2, tr = (task >> 1) & 1, tc = task & 1;
  const int ia = tri_row(pi), ib = pi - ia * (ia + 1) / 2;
  const int I = M.act_blk[a0 + ia], J = M.act_blk[a0 + ib];
  const int fI = M.fb[I], oI = M.boff[I];
  const double* A = L + oI + (K - fI) * 64 + tr * 32;               // rows 4tr..4tr+3 of L_IK
  const double* Bm = L + M.boff[J] + (K - M.fb[J]) * 64 + tc * 32;  // rows 4tc..4tc+3 of L_JK
  double* dst = L + oI + (J - fI) * 64 + tr * 32;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {  // logical 16-byte chunk of the K dimension
    D2 av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      av[i] = reinterpret_cast<const D2*>(A + i * KB)[c ^ bswz(4 * tr + i)];
      bv[i] = reinterpret_cast<const D2*>(Bm + i * KB)[c ^ bswz(4 * tc + i)];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i].y, bv[j].y, fma(av[i].x, bv[j].x, acc[i][j]));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    D2* drow = reinterpret_cast<D2*>(dst + i * KB);
    const int sw = bswz(4 * tr + i);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      D2 t = drow[(2 * tc + h) ^ sw];
      t.x -= acc[i][2 * h];
      t.y -= acc[i][2 * h + 1];
      drow[(2 * tc + h) ^ sw] = t;
    }
  }
}

#ifndef INTOPT_HOST_EMU
// The same update of one 8x8 block, C_IJ -= L_IK L_JK^T, by ONE WARP on the FP64 tensor core: two mma.sync.m8n8k4.f64
// (SASS: DMMA.8x8x4).  Lane l holds A[l/4][l%4 + 4s] and - because the second operand is the TRANSPOSE of a block stored the
// same way - B[l/4][l%4 + 4s] of the other block, and the two entries C[l/4][2(l%4) .. +1] of the result: 2 + 2 eight-byte
// and one 16-byte shared-memory access per lane for 512 multiply-adds, against 48 + 8 sixteen-byte accesses per 128 in the
// register-tiled version - the trailing updates were bound by shared-memory bandwidth (1150 wavefronts per step).
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// (UNR block pairs of one warp side by side: a pair is a chain of schedule lookups, loads, two dependent DMMAs and a
// read-modify-write, ~250 cycles of latency and almost no issue slots)
template <int UNR>
__device__ __forceinline__ void chol_update_pairs_dmma(const Meta& M, double* L, int K, int a0, int pair0, int stride, int np,
                                                       int lane) {
  const int r = lane >> 2, q = lane & 3;
  const int o_lo = bidx(r, q), o_hi = bidx(r, q + 4), o_c = r * KB + ((q ^ bswz(r)) << 1);
  // (a pair index beyond the list is clamped to the last pair: its loads and DMMAs are harmless duplicates, only the
  // store is skipped - no conditional code around the register arrays)
  double a_lo[UNR], a_hi[UNR], b_lo[UNR], b_hi[UNR];
  D2 t[UNR];
  double* cdst[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int want = pair0 + u * stride;
    const int pair = want < np ? want : np - 1;
    const int ia = tri_row(pair), ib = pair - ia * (ia + 1) / 2;
    const int I = M.act_blk[a0 + ia], J = M.act_blk[a0 + ib];
    const int fI = M.fb[I], oI = M.boff[I];
    const double* A = L + oI + (K - fI) * 64;
    const double* Bm = L + M.boff[J] + (K - M.fb[J]) * 64;
    cdst[u] = L + oI + (J - fI) * 64 + o_c;
    a_lo[u] = A[o_lo];
    a_hi[u] = A[o_hi];
    b_lo[u] = Bm[o_lo];
    b_hi[u] = Bm[o_hi];
    t[u] = *reinterpret_cast<const D2*>(cdst[u]);
  }
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    double acc0 = 0.0, acc1 = 0.0;
    dmma_884(acc0, acc1, a_lo[u], b_lo[u]);
    dmma_884(acc0, acc1, a_hi[u], b_hi[u]);
    t[u].x -= acc0;
    t[u].y -= acc1;
    if (pair0 + u * stride < np) *reinterpret_cast<D2*>(cdst[u]) = t[u];
  }
}
#endif

// one row of a panel block: row <- row * Linv^T (Linv lower triangular, the inverted diagonal block of the step)
IO_FORCEINLINE void chol_panel_row(double* blk, const double* Linv, int row) {
  double cin[KB], out[KB];
  load_row(blk, row, cin);
#pragma unroll
  for (int c = 0; c < KB; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k <= c) s = fma(cin[k], Linv[bidx(c, k)], s);
    out[c] = s;
  }
  store_row(blk, row, out);
}

IO_DEV bool block_cholesky(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double* L, double* scratch,
                           int* fail_flag, PhaseTimer* pt = nullptr) {
  const int mb = P.mb;
  if (cx.tid == 0) *fail_flag = 0;
  CTA_SYNC();
#ifndef INTOPT_HOST_EMU
  if (cx.nwarps > 2) {
    if (cx.warp == 0) diag_block_factor_publish(cx, L + M.boff[0] + (0 - M.fb[0]) * 64, scratch, fail_flag);
  } else
#endif
  if (cx.warp == 0) diag_block_factor_invert(cx, L + M.boff[0] + (0 - M.fb[0]) * 64, scratch, fail_flag);
  CTA_SYNC();
  const int o_tid = cx.nwarps > 1 ? cx.tid - cx.warpsize : 0;
  const int o_n = cx.nwarps > 1 ? cx.nthreads - cx.warpsize : 1;
#ifndef INTOPT_HOST_EMU
  if (cx.nwarps > 2) {
    // The pivot chain runs on warp 0: factor diagonal block K+1, and for that it needs ONE block of step K's panel,
    // L_(K+1,K), and the FACTOR of block K - not its inverse.  Warp 0 therefore forms that block itself, by substitution
    // against the factor it published a step earlier, and goes straight on to the next diagonal block; the inverse of
    // block K (which only the substitution sweeps act through) is formed by warp 1 in the meantime, warps 2.. form the
    // rest of the panel the same way as warp 0, all meet on a named barrier (warp 0 only ARRIVES there, after its block is
    // written: it never waits) and apply the trailing updates on the tensor core.  One full barrier per step; panel,
    // inversion and updates are all off the chain.
    for (int K = 0; K < mb; ++K) {
      double* diagK = L + M.boff[K] + (K - M.fb[K]) * 64;
      double* scrK = scratch + (K & 1) * S_DSCR;  // factor + reciprocal pivots of diagonal block K (published by warp 0)
      const int a0 = M.act_ptr[K], na = M.act_ptr[K + 1] - a0;
      const int lead = (na > 0 && M.act_blk[a0] == K + 1) ? 1 : 0;
      const int np = na * (na + 1) / 2;
      if (cx.warp == 0) {
        if (pt) { PHASE_MARK(*pt, 14); }
        const double* lead_panel = nullptr;
        if (lead) {
          double* blk = L + M.boff[K + 1] + (K - M.fb[K + 1]) * 64;
          if (cx.lane < KB) chol_panel_row_subst(blk, scrK, cx.lane);
          __syncwarp();
          lead_panel = blk;
        }
        asm volatile("bar.arrive 3, %0;" ::"r"(cx.nthreads) : "memory");
        if (pt) { PHASE_MARK(*pt, 11); }
        if (K + 1 < mb)
          diag_block_factor_publish(cx, L + M.boff[K + 1] + (K + 1 - M.fb[K + 1]) * 64, scratch + ((K + 1) & 1) * S_DSCR,
                                    fail_flag, lead_panel);
        if (pt) { PHASE_MARK(*pt, 13); }
      } else {
        // the inverse of diagonal block K (only the substitution sweeps act through it: every panel row is formed by
        // substitution against the published factor) is written here, off everybody's chain, by warp 1, while warps 2.. form
        // the panel
        if (cx.warp == 1) {
          diag_block_invert_published(cx, diagK, scrK);
        } else {
          for (int idx = cx.tid - 2 * cx.warpsize; idx < (na - lead) * KB; idx += cx.nthreads - 2 * cx.warpsize) {
            const int I = M.act_blk[a0 + lead + (idx >> 3)];
            chol_panel_row_subst(L + M.boff[I] + (K - M.fb[I]) * 64, scrK, idx & 7);
          }
        }
        asm volatile("bar.sync 3, %0;" ::"r"(cx.nthreads) : "memory");
        // one warp per block pair (two pairs of a warp in flight measured 1 % slower, three 2 %)
        if (na <= 32) {
          // Lane l holds the addressing of active block row l of this column; a pair's three block addresses then come from
          // five shuffles instead of two levels of dependent schedule lookups, and the pair's (row, column) position in the
          // lower triangle is advanced incrementally instead of being recomputed (a square root) per pair.
          int r_I = 0, r_f = 0, r_bo = 0, r_pk = 0;
          if (cx.lane < na) {
            r_I = M.act_blk[a0 + cx.lane];
            r_f = M.fb[r_I];
            r_bo = M.boff[r_I];
            r_pk = r_bo + (K - r_f) * 64;
          }
          const int stride = cx.nwarps - 1;
          int pair = lead + cx.warp - 1;
          int ia = tri_row(pair), ib = pair - ia * (ia + 1) / 2;
          const int rr = cx.lane >> 2, q = cx.lane & 3;
          const int o_lo = bidx(rr, q), o_hi = bidx(rr, q + 4), o_c = rr * KB + ((q ^ bswz(rr)) << 1);
          while (pair < np) {
            const unsigned full = 0xffffffffu;
            const int pkA = __shfl_sync(full, r_pk, ia), pkB = __shfl_sync(full, r_pk, ib);
            const int boA = __shfl_sync(full, r_bo, ia), fA = __shfl_sync(full, r_f, ia), J = __shfl_sync(full, r_I, ib);
            const double* A = L + pkA;
            const double* Bm = L + pkB;
            double* cdst = L + boA + (J - fA) * 64 + o_c;
            const double a_lo = A[o_lo], a_hi = A[o_hi], b_lo = Bm[o_lo], b_hi = Bm[o_hi];
            D2 t = *reinterpret_cast<const D2*>(cdst);
            double acc0 = 0.0, acc1 = 0.0;
            dmma_884(acc0, acc1, a_lo, b_lo);
            dmma_884(acc0, acc1, a_hi, b_hi);
            t.x -= acc0;
            t.y -= acc1;
            *reinterpret_cast<D2*>(cdst) = t;
            pair += stride;
            ib += stride;
            while (ib > ia) {
              ib -= ia + 1;
              ++ia;
            }
          }
        } else {
          for (int pair = lead + cx.warp - 1; pair < np; pair += cx.nwarps - 1)
            chol_update_pairs_dmma<1>(M, L, K, a0, pair, cx.nwarps - 1, np, cx.lane);
        }
      }
      CTA_SYNC();
    }
    return *fail_flag == 0;
  }
#endif
  for (int K = 0; K < mb; ++K) {
    const double* Linv = L + M.boff[K] + (K - M.fb[K]) * 64;
    // (a) panel rows
    const int a0 = M.act_ptr[K], na = M.act_ptr[K + 1] - a0;
    CTA_FOR(idx, na * KB) {
      const int I = M.act_blk[a0 + (idx >> 3)];
      chol_panel_row(L + M.boff[I] + (K - M.fb[I]) * 64, Linv, idx & 7);
    }
    CTA_SYNC();
    if (pt) { PHASE_MARK(*pt, 11); }
    // (b) trailing update with look-ahead
    const int np = na * (na + 1) / 2;
    const int lead = (na > 0 && M.act_blk[a0] == K + 1) ? 1 : 0;  // pair 0 updates block (K+1,K+1)
    if (cx.warp == 0) {
      if (pt) { PHASE_MARK(*pt, 12); }
      if (K + 1 < mb) {
        // the update of the next diagonal block by this step's panel block (K+1,K) is fused into its factorisation
        const double* lead_panel = lead ? L + M.boff[K + 1] + (K - M.fb[K + 1]) * 64 : nullptr;
        diag_block_factor_invert(cx, L + M.boff[K + 1] + (K + 1 - M.fb[K + 1]) * 64, scratch, fail_flag, lead_panel);
      }
      if (pt) { PHASE_MARK(*pt, 13); }
    }
    if (cx.nwarps == 1 || cx.warp > 0) {
      for (int task = lead * 4 + o_tid; task < np * 4; task += o_n) chol_update_task(P, M, L, K, a0, task);
    }
    CTA_SYNC();
    if (pt) { PHASE_MARK(*pt, 14); }
  }
  return *fail_flag == 0;
}

// ---------------------------------------------------------------------------------------------
// Solve L L^T z = r in place for NR right-hand sides at once (blocked substitution; the diagonal blocks
// hold Linv).  One barrier per block step: while the CTA applies block K's update to the rows below,
// warp 0 first updates the rows of the NEXT pivot block and immediately multiplies them by its inverted
// diagonal block, so the next step's pivot values are ready when the barrier opens.
// `ybuf` is 2*NR*8 doubles of scratch (double buffered, 16-byte aligned).
// ---------------------------------------------------------------------------------------------
template <int NR>
IO_FORCEINLINE double* rhs_sel(double* const (&r)[NR], int q) {
  return (NR > 1 && q) ? r[NR - 1] : r[0];  // NR is 1 or 2
}

// y[idx] (idx = q*8 + i) = row i of Linv (or of Linv^T) times the 8-vector v_q
template <int NR, bool TRANS>
IO_FORCEINLINE void pivot_apply(const double* Linv, double* const (&r)[NR], int blk, int idx, double* yout) {
  const int q = idx >> 3, i = idx & 7;
  double rk[KB];
  load_vec8(rhs_sel<NR>(r, q) + blk * KB, rk);
  // two independent half sums (even / odd terms): half the dependent FMA chain on warp 0's critical path
  double s = 0.0, s2 = 0.0;
  if (!TRANS) {
    double li[KB];
    load_row(Linv, i, li);
#pragma unroll
    for (int j = 0; j < KB; j += 2) {
      if (j <= i) s = fma(li[j], rk[j], s);
      if (j + 1 <= i) s2 = fma(li[j + 1], rk[j + 1], s2);
    }
  } else {
#pragma unroll
    for (int j = 0; j < KB; j += 2) {
      if (j >= i) s = fma(Linv[bidx(j, i)], rk[j], s);
      if (j + 1 >= i) s2 = fma(Linv[bidx(j + 1, i)], rk[j + 1], s2);
    }
  }
  yout[idx] = s + s2;
}

template <int NR>
IO_DEV void block_solve(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* L, double* const (&r)[NR],
                        double* ybuf) {
  const int mb = P.mb;
  const int o_tid = cx.nwarps > 1 ? cx.tid - cx.warpsize : 0;
  const int o_n = cx.nwarps > 1 ? cx.nthreads - cx.warpsize : 1;
  // ---- forward: L y = r --------------------------------------------------------------------------
  if (cx.warp == 0) {
    LANE_FOR(idx, 0, NR * KB) pivot_apply<NR, false>(L + M.boff[0] + (0 - M.fb[0]) * 64, r, 0, idx, ybuf);
  }
  CTA_SYNC();
  for (int K = 0; K < mb; ++K) {
    const double* ycur = ybuf + (K & 1) * NR * KB;
    double* ynext = ybuf + ((K + 1) & 1) * NR * KB;
    const int a0 = M.act_ptr[K], na = M.act_ptr[K + 1] - a0;
    const int lead = (na > 0 && M.act_blk[a0] == K + 1) ? 1 : 0;  // the next pivot block is updated by this step
    if (cx.warp == 0) {
      LANE_FOR(idx, 0, NR * KB) {
        rhs_sel<NR>(r, idx >> 3)[K * KB + (idx & 7)] = ycur[idx];  // y_K is final: store it
        if (lead) {
          const int q = idx >> 3, rr = idx & 7;
          double row[KB], yk[KB];
          load_row(L + M.boff[K + 1] + (K - M.fb[K + 1]) * 64, rr, row);
          load_vec8(ycur + q * KB, yk);
          double sacc = 0.0, sacc2 = 0.0;  // two half sums: half the dependent chain
#pragma unroll
          for (int k = 0; k < KB; k += 2) {
            sacc = fma(row[k], yk[k], sacc);
            sacc2 = fma(row[k + 1], yk[k + 1], sacc2);
          }
          rhs_sel<NR>(r, q)[(K + 1) * KB + rr] -= sacc + sacc2;
        }
      }
      WARP_SYNC();
      if (K + 1 < mb) {
        LANE_FOR(idx, 0, NR * KB)
        pivot_apply<NR, false>(L + M.boff[K + 1] + (K + 1 - M.fb[K + 1]) * 64, r, K + 1, idx, ynext);
      }
    }
    if (cx.nwarps == 1 || cx.warp > 0) {
      // one item = one row of a block, applied to every right-hand side (the row is read from shared memory once)
      const int nrest = na - lead;
      for (int idx = o_tid; idx < nrest * KB; idx += o_n) {
        const int I = M.act_blk[a0 + lead + (idx >> 3)], rr = idx & 7;
        double row[KB];
        load_row(L + M.boff[I] + (K - M.fb[I]) * 64, rr, row);
#pragma unroll
        for (int q = 0; q < NR; ++q) {
          double yk[KB];
          load_vec8(ycur + q * KB, yk);
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < KB; ++k) sacc = fma(row[k], yk[k], sacc);
          rhs_sel<NR>(r, q)[I * KB + rr] -= sacc;
        }
      }
    }
    CTA_SYNC();
  }
  // ---- backward: L^T z = y -----------------------------------------------------------------------
  if (cx.warp == 0) {
    LANE_FOR(idx, 0, NR * KB)
    pivot_apply<NR, true>(L + M.boff[mb - 1] + (mb - 1 - M.fb[mb - 1]) * 64, r, mb - 1, idx, ybuf + ((mb - 1) & 1) * NR * KB);
  }
  CTA_SYNC();
  for (int I = mb - 1; I >= 0; --I) {
    const int f = M.fb[I];
    const double* zcur = ybuf + (I & 1) * NR * KB;
    double* znext = ybuf + ((I + 1) & 1) * NR * KB;  // buffer of parity (I-1)&1
    const int lead = (I > 0 && f <= I - 1) ? 1 : 0;  // block (I, I-1) exists: row block I-1 is updated by this step
    if (cx.warp == 0) {
      LANE_FOR(idx, 0, NR * KB) rhs_sel<NR>(r, idx >> 3)[I * KB + (idx & 7)] = zcur[idx];  // z_I is final
      if (lead) {
        LANE_FOR(idx, 0, NR * 4) {  // the four 16-byte chunks (two columns each) of block (I, I-1)
          const int q = idx >> 2, c = idx & 3;
          const double* blk = L + M.boff[I] + (I - 1 - f) * 64;
          double zi[KB];
          load_vec8(zcur + q * KB, zi);
          double sx = 0.0, sy = 0.0, sx2 = 0.0, sy2 = 0.0;  // half sums: half the dependent chain
#pragma unroll
          for (int i = 0; i < KB; i += 2) {
            const D2 t2 = reinterpret_cast<const D2*>(blk + i * KB)[c ^ bswz(i)];
            const D2 t3 = reinterpret_cast<const D2*>(blk + (i + 1) * KB)[c ^ bswz(i + 1)];
            sx = fma(t2.x, zi[i], sx);
            sy = fma(t2.y, zi[i], sy);
            sx2 = fma(t3.x, zi[i + 1], sx2);
            sy2 = fma(t3.y, zi[i + 1], sy2);
          }
          D2* dst = reinterpret_cast<D2*>(rhs_sel<NR>(r, q) + (I - 1) * KB) + c;
          D2 t2 = *dst;
          t2.x -= sx + sx2;
          t2.y -= sy + sy2;
          *dst = t2;
        }
      }
      WARP_SYNC();
      if (I > 0) {
        LANE_FOR(idx, 0, NR * KB)
        pivot_apply<NR, true>(L + M.boff[I - 1] + (I - 1 - M.fb[I - 1]) * 64, r, I - 1, idx, znext);
      }
    }
    if (cx.nwarps == 1 || cx.warp > 0) {
      const int nk = I - f - lead;  // block columns f .. I-1-lead
      for (int idx = o_tid; idx < nk * 4; idx += o_n) {  // (two columns of a block, for every right-hand side)
        const int Kc = f + (idx >> 2), c = idx & 3;
        const double* blk = L + M.boff[I] + (Kc - f) * 64;
        D2 col[KB];
#pragma unroll
        for (int i = 0; i < KB; ++i) col[i] = reinterpret_cast<const D2*>(blk + i * KB)[c ^ bswz(i)];
#pragma unroll
        for (int q = 0; q < NR; ++q) {
          double zi[KB];
          load_vec8(zcur + q * KB, zi);
          double sx = 0.0, sy = 0.0;
#pragma unroll
          for (int i = 0; i < KB; ++i) {
            sx = fma(col[i].x, zi[i], sx);
            sy = fma(col[i].y, zi[i], sy);
          }
          D2* dst = reinterpret_cast<D2*>(rhs_sel<NR>(r, q) + Kc * KB) + c;
          D2 t2 = *dst;
          t2.x -= sx;
          t2.y -= sy;
          *dst = t2;
        }
      }
    }
    CTA_SYNC();
  }
}

IO_FORCEINLINE double run_total(const double* gp, int len) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = 0;
  for (; j + 4 <= len; j += 4) {
    const double a0 = gp[j], a1 = gp[j + 1], a2 = gp[j + 2], a3 = gp[j + 3];
    s0 += a0;
    s1 += a1;
    s2 += a2;
    s3 += a3;
  }
  for (; j < len; ++j) s0 += gp[j];
  return (s0 + s1) + (s2 + s3);
}

// Sum of a run of at most 16 consecutive entries (the row kernel's pieces, RS_MAXLEN): fully predicated, no loop.
// The 16 offsets are visited in an order rotated by `rot` (chosen on the host per lane so that start + rot == lane
// mod 16): at every step the 32 lanes of a warp then touch 16 different 8-byte bank pairs, two lanes each - the
// conflict-free pattern - wherever the runs start.
IO_FORCEINLINE double run_total16(const double* gp, int len, int rot) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const int o0 = (j + rot) & 15, o1 = (j + 1 + rot) & 15, o2 = (j + 2 + rot) & 15, o3 = (j + 3 + rot) & 15;
    if (o0 < len) s0 += gp[o0];
    if (o1 < len) s1 += gp[o1];
    if (o2 < len) s2 += gp[o2];
    if (o3 < len) s3 += gp[o3];
  }
  return (s0 + s1) + (s2 + s3);
}
// header of a run piece: start | len << 16 | rot << 24
IO_FORCEINLINE int rs_start(uint32_t h) { return (int)(h & 0xffffu); }
IO_FORCEINLINE int rs_len(uint32_t h) { return (int)((h >> 16) & 0xffu); }
IO_FORCEINLINE int rs_rot(uint32_t h) { return (int)(h >> 24); }

// out = A src (NV = 1), or out = A src and out2 = A src2 in one pass over A (NV = 2: the index / header stream,
// the address arithmetic and the scan bookkeeping are shared by the two products).
template <int NV>
IO_DEV void spmv_rows_n(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* src, const double* src2,
                        double* out, double* out2) {
  if (P.rs_hdr) {
    // unit-weight A in run-length form: every lane sums one run of consecutive src entries, rows are whole
    // lane ranges of a slot and are combined by a segmented shuffle scan (fixed order => deterministic)
#ifdef INTOPT_HOST_EMU
    for (int slot = 0; slot < P.rs_slots; ++slot)
      for (int l = 0; l < 32; ++l) {
        const uint32_t h = P.rs_hdr[slot * 32 + l];
        const int row = P.rs_row[slot * 32 + l];
        if (row == 0xffff) continue;
        const bool first = l == 0 || P.rs_row[slot * 32 + l - 1] != row;
        const double s = run_total16(src + rs_start(h), rs_len(h), rs_rot(h));
        out[row] = first ? s : out[row] + s;
        if (NV == 2) {
          const double s2 = run_total16(src2 + rs_start(h), rs_len(h), rs_rot(h));
          out2[row] = first ? s2 : out2[row] + s2;
        }
      }
#else
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
      if (nxt < P.rs_slots) {  // next slot's headers are in flight while this one is summed
        hn = P.rs_hdr[nxt * 32 + cx.lane];
        rown = P.rs_row[nxt * 32 + cx.lane];
      }
      double s = run_total16(src + rs_start(h), rs_len(h), rs_rot(h));
      double s2 = 0.0;
      if (NV == 2) s2 = run_total16(src2 + rs_start(h), rs_len(h), rs_rot(h));
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, s, off);
        const int r = __shfl_up_sync(0xffffffffu, row, off);
        const bool take = cx.lane >= off && r == row;
        if (take) s += v;
        if (NV == 2) {
          const double v2 = __shfl_up_sync(0xffffffffu, s2, off);
          if (take) s2 += v2;
        }
      }
      const int rnext = __shfl_down_sync(0xffffffffu, row, 1);
      if (row != 0xffff && (cx.lane == 31 || rnext != row)) {
        out[row] = s;
        if (NV == 2) out2[row] = s2;
      }
    }
#endif
    return;
  }
  const int W = cx.warpsize;
  WARP_ROWS(i, P.m) {
    const int lo = M.csr_ptr[i], hi = M.csr_ptr[i + 1];
    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
    for (int e = lo + cx.lane; e < hi; e += 4 * W) {
      const int e1 = e + W, e2 = e + 2 * W, e3 = e + 3 * W;
      const int c0 = P.csr_col[e];
      const int c1 = e1 < hi ? P.csr_col[e1] : c0;
      const int c2 = e2 < hi ? P.csr_col[e2] : c0;
      const int c3 = e3 < hi ? P.csr_col[e3] : c0;
      if (P.a_unit) {
        s0 += src[c0];
        if (e1 < hi) s1 += src[c1];
        if (e2 < hi) s0 += src[c2];
        if (e3 < hi) s1 += src[c3];
        if (NV == 2) {
          t0 += src2[c0];
          if (e1 < hi) t1 += src2[c1];
          if (e2 < hi) t0 += src2[c2];
          if (e3 < hi) t1 += src2[c3];
        }
      } else {
        const double a0 = M.csr_val[e], a1 = e1 < hi ? M.csr_val[e1] : 0.0, a2 = e2 < hi ? M.csr_val[e2] : 0.0,
                     a3 = e3 < hi ? M.csr_val[e3] : 0.0;
        s0 = fma(a0, src[c0], s0);
        if (e1 < hi) s1 = fma(a1, src[c1], s1);
        if (e2 < hi) s0 = fma(a2, src[c2], s0);
        if (e3 < hi) s1 = fma(a3, src[c3], s1);
        if (NV == 2) {
          t0 = fma(a0, src2[c0], t0);
          if (e1 < hi) t1 = fma(a1, src2[c1], t1);
          if (e2 < hi) t0 = fma(a2, src2[c2], t0);
          if (e3 < hi) t1 = fma(a3, src2[c3], t1);
        }
      }
    }
    const double s = warp_sum(s0 + s1);
    if (cx.lane == 0) out[i] = s;
    if (NV == 2) {
      const double t = warp_sum(t0 + t1);
      if (cx.lane == 0) out2[i] = t;
    }
  }
}
IO_DEV void spmv_rows(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* src, double* out) {
  spmv_rows_n<1>(cx, P, M, src, nullptr, out, nullptr);
}

// (A^T y)_k through the interleaved column groups (lanes of a warp own 32 consecutive columns).  Row indices
// come as 16-byte chunks of eight; the chunks of a column (at most four for up to 32 terms) are all requested
// requested two at a time.
struct IdxChunk {
  uint32_t w[4];
};
IO_FORCEINLINE IdxChunk load_idx_chunk(const uint16_t* p) {
  const U4 q = *reinterpret_cast<const U4*>(p);
  IdxChunk c;
  c.w[0] = q.x;
  c.w[1] = q.y;
  c.w[2] = q.z;
  c.w[3] = q.w;
  return c;
}
#define IDX_LO(c, j) ((int)((c).w[j] & 0xffffu))
#define IDX_HI(c, j) ((int)((c).w[j] >> 16))

template <int NY>
IO_FORCEINLINE void col_dot_n(const ProblemDev& P, const Meta& M, int k, const double* y1, const double* y2, double& r1,
                              double& r2) {
  const int g = k >> 5, l = k & 31, g0 = M.csc_gptr[g], len = M.csc_gptr[g + 1] - g0;  // len % 8 == 0
  const uint16_t* __restrict__ rr = P.csc_row + (((size_t)(g0 >> 3)) * 32 + l) * 8;
  double a1 = 0.0, b1 = 0.0, a2 = 0.0, b2 = 0.0;
  constexpr int NC = 2;  // index chunks in flight (3 and 4 measured slower: registers)
  for (int t0 = 0; t0 < len; t0 += 8 * NC) {
    IdxChunk q[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (t0 + 8 * c < len) q[c] = load_idx_chunk(rr + (size_t)((t0 >> 3) + c) * 256);
    if (P.a_unit) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (t0 + 8 * c < len) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i0 = IDX_LO(q[c], j), i1 = IDX_HI(q[c], j);
            a1 += y1[i0];
            b1 += y1[i1];
            if (NY == 2) {
              a2 += y2[i0];
              b2 += y2[i1];
            }
          }
        }
    } else {
      const double* __restrict__ vv = M.csc_val + (size_t)(g0 + t0) * 32 + l;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (t0 + 8 * c < len) {
          double wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = vv[(8 * c + j) * 32];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i0 = IDX_LO(q[c], j), i1 = IDX_HI(q[c], j);
            a1 = fma(wv[2 * j], y1[i0], a1);
            b1 = fma(wv[2 * j + 1], y1[i1], b1);
            if (NY == 2) {
              a2 = fma(wv[2 * j], y2[i0], a2);
              b2 = fma(wv[2 * j + 1], y2[i1], b2);
            }
          }
        }
    }
  }
  r1 = a1 + b1;
  if (NY == 2) r2 = a2 + b2;
}

IO_FORCEINLINE double col_dot(const ProblemDev& P, const Meta& M, int k, const double* yv) {
  double s, unused_;
  col_dot_n<1>(P, M, k, yv, nullptr, s, unused_);
  return s;
}
IO_FORCEINLINE void col_dot2(const ProblemDev& P, const Meta& M, int k, const double* y1, const double* y2, double& s1,
                             double& s2) {
  col_dot_n<2>(P, M, k, y1, y2, s1, s2);
}

// out[k] = (A^T y)_k for every column k (NY = 1 or 2 vectors at once).
template <int NY>
IO_DEV void at_mult(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* y1, const double* y2,
                    double* out1, double* out2) {
  const int n = P.n;
  CTA_FOR(k, n) {
    double a, b2;
    col_dot_n<NY>(P, M, k, y1, y2, a, b2);
    out1[k] = a;
    if (NY == 2) out2[k] = b2;
  }
}

IO_DEV Meta global_meta(const ProblemDev& P) {
  Meta M;
  M.fb = P.g_fb;
  M.boff = P.g_boff;
  M.act_ptr = P.g_act_ptr;
  M.act_blk = P.g_act_blk;
  M.csr_ptr = P.g_csr_ptr;
  M.csc_gptr = P.g_csc_gptr;
  M.asm_gptr = P.g_asm_gptr;
  M.asm_rptr = P.g_asm_rptr;
  M.csr_val = P.csr_val;
  M.csc_val = P.csc_val;
  M.asm_w = P.asm_w;
  return M;
}

// Copies the schedule arrays into shared memory (`dst` holds P.meta_ints ints) and returns pointers to the copies.
IO_DEV Meta stage_meta(const CtaCtx& cx, const ProblemDev& P, int* dst) {
  const int cnt[8] = {P.mb, P.mb + 1, P.mb + 1, P.n_act, P.m + 1, ((P.n + 31) >> 5) + 1, (P.n_asm >> 5) + 1,
                      (P.n_asm >> 5) + 1};
  const int* src[8] = {P.g_fb, P.g_boff, P.g_act_ptr, P.g_act_blk, P.g_csr_ptr, P.g_csc_gptr, P.g_asm_gptr,
                       P.g_asm_rptr};
  int* d = dst;
  for (int a = 0; a < 8; ++a) {
    CTA_FOR(i, cnt[a]) d[i] = src[a][i];
    d += cnt[a];
  }
  Meta M;
  d = dst;
  M.fb = d;
  d += cnt[0];
  M.boff = d;
  d += cnt[1];
  M.act_ptr = d;
  d += cnt[2];
  M.act_blk = d;
  d += cnt[3];
  M.csr_ptr = d;
  d += cnt[4];
  M.csc_gptr = d;
  d += cnt[5];
  M.asm_gptr = d;
  d += cnt[6];
  M.asm_rptr = d;
  M.csr_val = P.csr_val;
  M.csc_val = P.csc_val;
  M.asm_w = P.asm_w;
  CTA_SYNC();
  return M;
}

// Per-instance values of A: the three value streams the kernels read (by rows, by column groups, assembly weights)
// then live in the CTA's workspace (`buf` = nnz + n_csc_pad + n_asm_pad doubles, laid out like the problem's own);
// instance_value_streams points the schedule at them once per CTA, fill_instance_values copies one instance's
// numbers (av, in the caller's CSR order) in.
IO_DEV void instance_value_streams(const ProblemDev& P, Meta& M, double* buf) {
  M.csr_val = buf;
  M.csc_val = buf + P.nnz;
  M.asm_w = buf + P.nnz + P.n_csc_pad;
}
IO_DEV void fill_instance_values(const CtaCtx& cx, const ProblemDev& P, const double* av, double* buf) {
  double* cv = buf;
  double* sv = cv + P.nnz;
  double* wv = sv + P.n_csc_pad;
  CTA_FOR(e, P.nnz) cv[e] = av[P.csr_src[e]];
  CTA_FOR(i, P.n_csc_pad) {
    const int src = P.csc_src[i];
    sv[i] = src >= 0 ? av[src] : 0.0;
  }
  CTA_FOR(i, P.n_asm_pad) {
    const int sa = P.asm_sa[i];
    wv[i] = sa >= 0 ? av[sa] * av[P.asm_sb[i]] : 0.0;
  }
  CTA_SYNC();
}

// workspace n-vector slots (global, per CTA)
enum { W_X = 0, W_T, W_D, W_R2, W_P, W_W, W_DX, W_DXDT, W_ATY, W_NVEC };
// doubles of the n-vector block of a workspace slot; even, so that whatever follows (a spilled factor / gather
// region, read with 16-byte accesses) stays 16-byte aligned
IO_HD long long ws_nvec_doubles(int n) { return ((long long)W_NVEC * n + 1) & ~1LL; }
// shared-memory m-vector slots
enum { S_B = 0, S_Y, S_R1, S_Q, S_V, S_NVEC };
constexpr int S_SCRATCH = 4 * KB + 2 * S_DSCR + 8 * 32 + 4;  // ybuf (double buffered, 2 rhs) + diag scratch (double buffered) + red + flag + mbarrier

struct SmemLayout {
  double *L, *G, *bs, *y, *r1, *q, *v, *ybuf, *dscr, *red;  // G: gather region of 2*gather_len(n) doubles
  D2* scan;  // scratch of the double-double prefix scan: scan_len(n) D2
  int* flag;
  unsigned long long* mbar;  // mbarrier of the cold-start factor prefetch (bulk async copy)
  int* meta;  // P.meta_ints ints when the schedule is staged in shared memory
};

IO_HD int gather_len(int n) { return (gpad(n) + 3) & ~1; }
IO_HD int scan_len(int n) { return ((n + 7) >> 3) + 40; }  // D2 entries of prefix-scan scratch
IO_HD bool scan_in_factor(int bsize, int n) { return bsize >= 2 * scan_len(n); }

// Carve shared memory (and, for spilled pieces, the tail of the per-CTA workspace).  ALLSMEM = factor and
// gather buffers both live in shared memory: every pointer is then derived unconditionally from the shared
// array, which lets the compiler emit LDS/STS instead of generic loads (the common, fast instantiation).
template <bool ALLSMEM>
IO_DEV SmemLayout carve(const ProblemDev& P, double* smem, double* ws_tail, int factor_in_smem, int gather_in_smem) {
  SmemLayout s;
  double* p = smem;
  if (ALLSMEM) {
    s.L = p;
    p += P.bsize;
    s.G = p;
    p += 2 * gather_len(P.n);
  } else {
    if (factor_in_smem) {
      s.L = p;
      p += P.bsize;
    } else {
      s.L = ws_tail;
      ws_tail += P.bsize;
    }
    if (gather_in_smem) {  // gather buffer + the primal iterate x
      s.G = p;
      p += 2 * gather_len(P.n);
    } else {
      s.G = ws_tail;
      ws_tail += 2 * gather_len(P.n);
    }
  }
  s.bs = p + S_B * P.mv;
  s.y = p + S_Y * P.mv;
  s.r1 = p + S_R1 * P.mv;
  s.q = p + S_Q * P.mv;
  s.v = p + S_V * P.mv;
  p += S_NVEC * P.mv;
  s.ybuf = p;
  s.dscr = p + 4 * KB;
  s.red = s.dscr + 2 * S_DSCR;
  s.flag = (int*)(s.red + 8 * 32);
  s.mbar = reinterpret_cast<unsigned long long*>(s.red + 8 * 32 + 2);
  // the prefix-scan scratch aliases the factor storage when it fits (the scan runs before the factor is zero-filled)
  const int own_scan = scan_in_factor(P.bsize, P.n) ? 0 : 2 * scan_len(P.n);
  s.scan = own_scan ? reinterpret_cast<D2*>(s.red + 8 * 32 + 4) : reinterpret_cast<D2*>(s.L);
  s.meta = (int*)(s.red + 8 * 32 + 4 + own_scan);
#ifndef INTOPT_HOST_EMU
  if (ALLSMEM) {
    __builtin_assume(__isShared(s.L));
    __builtin_assume(__isShared(s.G));
  }
  __builtin_assume(__isShared(s.bs));
  __builtin_assume(__isShared(s.y));
  __builtin_assume(__isShared(s.r1));
  __builtin_assume(__isShared(s.q));
  __builtin_assume(__isShared(s.v));
  __builtin_assume(__isShared(s.ybuf));
  __builtin_assume(__isShared(s.dscr));
  __builtin_assume(__isShared(s.red));
#endif
  return s;
}

// ---------------------------------------------------------------------------------------------
// Cold-start factor prefetch.  CUDA build, factor in shared memory: one elected thread issues bulk async
// copies (TMA, cp.async.bulk global -> shared) that complete on an mbarrier, so the 60-80 KB factor streams in
// from L2 underneath the instance's set-up phase; otherwise a plain vectorised copy at the wait point.
// ---------------------------------------------------------------------------------------------
#ifndef INTOPT_HOST_EMU
IO_FORCEINLINE uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
IO_FORCEINLINE void mbar_init(unsigned long long* mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
IO_FORCEINLINE void mbar_wait(unsigned long long* mbar, unsigned parity) {
  const uint32_t addr = smem_addr_u32(mbar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
#endif

// Begin copying the precomputed factor `src` (P.bsize doubles, 16-byte aligned) into L.  Call after a CTA
// barrier that ends every earlier use of L.
IO_DEV void cold_copy_begin(const CtaCtx& cx, const ProblemDev& P, const double* src, double* L, unsigned long long* mbar,
                            bool l_in_smem) {
#ifdef INTOPT_HOST_EMU
  (void)cx; (void)P; (void)src; (void)L; (void)mbar; (void)l_in_smem;
#else
  if (l_in_smem && cx.tid == 0) {
    const uint32_t bytes = (uint32_t)P.bsize * 8u, bar = smem_addr_u32(mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to L vs the async writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t len = bytes - off < 32768u ? bytes - off : 32768u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr_u32(L) + off),
                   "l"((const char*)src + off), "r"(len), "r"(bar)
                   : "memory");
    }
  }
#endif
}
// Complete the copy: on return every thread may read L.
IO_DEV void cold_copy_finish(const CtaCtx& cx, const ProblemDev& P, const double* src, double* L, unsigned long long* mbar,
                             bool l_in_smem, unsigned parity) {
#ifndef INTOPT_HOST_EMU
  if (l_in_smem) {
    mbar_wait(mbar, parity);
    return;
  }
#else
  (void)mbar; (void)l_in_smem; (void)parity;
#endif
  const D2* __restrict__ s2 = reinterpret_cast<const D2*>(src);
  D2* d2 = reinterpret_cast<D2*>(L);
  CTA_FOR(e, P.bsize >> 1) d2[e] = s2[e];  // bsize is a multiple of 64
  CTA_SYNC();
}

// Optional mapping between the caller's own layout / precision and the standard form, applied inside the solver
// kernels (saves the separate pack / unpack launches of the host glue).  `in == nullptr` / `out == nullptr`: off.
//   input  : standard-form entry j (< n_in) = in_scale[j] * in[inst][in_idx[j]]   (in_idx[j] < 0: 0; in_scale null: 1)
//            forward: the cost vector c, backward: the upstream gradient; the packed row is written to the kernel's
//            regular input buffer (c / grad_x), which must then be writable
//   output : user entry i (< n_out) = std[out_idx[i]], or out_const[i] where out_idx[i] < 0 (out_const null: 0)
//            forward: x = x_hat with the presolve-fixed variables re-inserted, backward: dL/dc
struct UserMap {
  const void* in;
  int in_f32;
  long long in_stride;
  const int* in_idx;
  const double* in_scale;
  void* out;
  int out_f32;
  long long out_stride;
  int n_out;
  const int* out_idx;
  const double* out_const;
};
IO_FORCEINLINE double um_load(const UserMap& u, long long inst, int k) {
  const long long o = inst * u.in_stride + k;
  return u.in_f32 ? (double)reinterpret_cast<const float*>(u.in)[o] : reinterpret_cast<const double*>(u.in)[o];
}
IO_FORCEINLINE void um_store(const UserMap& u, long long inst, int i, double v) {
  const long long o = inst * u.out_stride + i;
  if (u.out_f32)
    reinterpret_cast<float*>(u.out)[o] = (float)v;
  else
    reinterpret_cast<double*>(u.out)[o] = v;
}

struct FwdParams {
  int B;
  const double* c;
  long long c_stride;
  const double* b_batch;  // original row order, B x m, or null
  long long b_stride;
  const double* b_problem;  // internal row order or null
  const double* init;       // B x (2n+m+2) or null
  UserMap um;               // optional user-layout input / output (see UserMap)
  const double* a_vals;     // per-instance values of A, B x nnz (caller's CSR order), or null
  long long a_stride;
  double* vals_ws;          // per CTA slot: nnz + n_csc_pad + n_asm_pad doubles (only with a_vals)
  long long vals_ws_stride;
  double* xhat;
  double* state;
  int* status;
  int* iters;
  double* ws;  // per CTA slot: W_NVEC n-vectors (+ spilled factor / gather buffer)
  long long ws_stride;
  int* counter;  // instance dispenser (zeroed before launch)
  int factor_in_smem, gather_in_smem;
  // Cold-start factor: without a warm start every instance begins at x = t = 1, so the first normal matrix
  // A.A^T + delta.I and its factor are constants of the problem; they are computed once (cold_factor_build,
  // same code => same bits) and only copied in at iteration 1.  Null / *cold_ok == 0: assemble + factor as usual.
  const double* cold_factor;
  const int* cold_ok;
  double alpha0, tol, thr, damping;
  int max_iter;
  int plain;    // 1: pc=False upstream — one pass per iteration with centring gamma = beta * mean(x t) (:1392)
  double beta;
#ifdef INTOPT_PROFILE
  unsigned long long* phase_cycles;  // null unless switched on (profiling build only)
  int debug;                         // switches for bottleneck isolation (env INTOPT_DEBUG)
#endif
};

// Factor L (block_cholesky) and form the two products out1 = A src1, out2 = A src2 (out2 may be null: one product),
// side by side when the CTA has at least 4 warps: the products need nothing from the factorisation.  On return
// (after a full CTA barrier) *S.flag tells whether the factorisation succeeded.
constexpr int RHS_WARPS = 2;
IO_DEV void factor_and_rhs(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double* L, const SmemLayout& S,
                           const double* src1, const double* src2, double* out1, double* out2, PhaseTimer* pt) {
#ifndef INTOPT_HOST_EMU
  if (cx.nwarps >= 4 + RHS_WARPS - 1) {
    const int ga = cx.nwarps - RHS_WARPS;
    if (cx.warp < ga) {
      CtaCtx ca = cx;
      ca.nthreads = ga * 32;
      ca.nwarps = ga;
      ca.bar_id = 1;
      ca.bar_n = ga * 32;
      block_cholesky(ca, P, M, L, S.dscr, S.flag, pt);
    } else {
      CtaCtx cb = cx;
      cb.tid = cx.tid - ga * 32;
      cb.warp = cx.warp - ga;
      cb.nthreads = RHS_WARPS * 32;
      cb.nwarps = RHS_WARPS;
      cb.bar_id = 2;
      cb.bar_n = RHS_WARPS * 32;
      if (out2)
        spmv_rows_n<2>(cb, P, M, src1, src2, out1, out2);
      else
        spmv_rows_n<1>(cb, P, M, src1, nullptr, out1, nullptr);
    }
    __syncthreads();
    return;
  }
#endif
  block_cholesky(cx, P, M, L, S.dscr, S.flag, pt);
  if (out2)
    spmv_rows_n<2>(cx, P, M, src1, src2, out1, out2);
  else
    spmv_rows_n<1>(cx, P, M, src1, nullptr, out1, nullptr);
  CTA_SYNC();
}

// ---- where the forward's element-wise n-vectors live ---------------------------------------------------------------
// Nine n-vectors (x, t, c, D, r_d, w, p, d_x, d_x*d_t) are only ever touched element-wise, and always by the same thread:
// element k belongs to thread k % nthreads.  In the workspace in global memory every pass over them is a chain of L2 round
// trips (the loops are not batched: registers).  Tensor memory is the other place that can hold them: 256 KB per SM
// that this kernel does not otherwise use, addressed per lane (a thread reads and writes only its own 32-bit cells, which
// is exactly the access pattern here), with a load-to-use latency of some 40 cycles instead of an L2 hit's ~700.  The loops
// below are written once against an accessor; NVGlobal is the workspace version, NVTmem the tensor-memory version
// (tcgen05.ld / tcgen05.st, shape 32x32b.x2 = one double per thread and access).
enum { V_X = 0, V_T, V_C, V_D, V_R2, V_W, V_P, V_DX, V_DXDT, V_SLOTS };
constexpr int NV_EMAX = 7;     // elements per thread and vector: 9 vectors * 7 elements * 2 columns = 126 <= 128
constexpr int NV_COLS = 128;   // tensor-memory columns of one group of four warps (warps w and w + 4 share lanes)

struct NVGlobal {
  static constexpr bool kTmem = false;
  typedef double Val;
  double* v[V_SLOTS];
  int n;
  IO_DEV Val ld(int s, int, int k) const { return k < n ? v[s][k] : 1.0; }
  IO_DEV void ld_wait() const {}
  IO_DEV double get(Val& r) const { return r; }
  IO_DEV void st(int s, int, int k, double val) const {
    if (k < n) v[s][k] = val;
  }
  IO_DEV void st_wait() const {}
};

#ifndef INTOPT_HOST_EMU
struct NVTmem {
  static constexpr bool kTmem = true;
  struct Val {
    uint32_t lo, hi;
  };
  uint32_t base;  // this thread's lane (bits 31:16) and the first column of its warp group
  __device__ __forceinline__ Val ld(int s, int j, int) const {
    Val r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                 : "=r"(r.lo), "=r"(r.hi)
                 : "r"(base + 2u * (unsigned)(s * NV_EMAX + j))
                 : "memory");
    return r;
  }
  __device__ __forceinline__ void ld_wait() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
  __device__ __forceinline__ double get(Val& r) const {
    asm volatile("" : "+r"(r.lo), "+r"(r.hi));  // (pins the first use behind the wait)
    return __hiloint2double((int)r.hi, (int)r.lo);
  }
  __device__ __forceinline__ void st(int s, int j, int, double val) const {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(base + 2u * (unsigned)(s * NV_EMAX + j)),
                 "r"((uint32_t)__double2loint(val)), "r"((uint32_t)__double2hiint(val))
                 : "memory");
  }
  __device__ __forceinline__ void st_wait() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
};
#endif

// With tensor memory every thread runs the same number of trips (the accesses are warp-collective) and the element index
// is checked where it matters; with the workspace the loop is CTA_FOR.  Per thread the elements come in ascending order.
#define NV_FOR(j, k) for (int j = 0, k = cx.tid; TM ? j < nv_trips : k < n; ++j, k += cx.nthreads)

// PLAIN = false: Mehrotra predictor-corrector (pc=True upstream, every reference driver).  PLAIN = true: pc=False.  A
// template parameter, and one kernel per value, rather than a run-time switch: the run-time version cost the
// predictor-corrector path 0.45 %, both instantiations inside one kernel 3.8 % (register allocation).
template <bool PLAIN, class NV>
IO_DEV void forward_instance_nv(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const FwdParams& a, int inst, double* wsi,
                                const SmemLayout& S, const NV& nv, bool cold = false, unsigned cold_parity = 0) {
  constexpr bool TM = NV::kTmem;
  if (a.a_vals)  // (M already points at the CTA's value streams, see instance_value_streams)
    fill_instance_values(cx, P, a.a_vals + (long long)inst * a.a_stride, a.vals_ws + (long long)blockIdx_x() * a.vals_ws_stride);
  const int m = P.m, n = P.n, mv = P.mv;
  const int nv_trips = (n + cx.nthreads - 1) / cx.nthreads;
  double *L = S.L, *G = S.G, *bs = S.bs, *y = S.y, *r1 = S.r1, *q = S.q, *v = S.v;
  // D is also kept in the workspace (the assembly's prefix sums read it in their own order)
  double* __restrict__ Dg = wsi + (long long)W_D * n;
  // A^T y, carried along the iteration: y only ever changes by alpha * (v + q dtau), and A^T(v + q dtau) is what the
  // direction was built from, so the product never has to be formed again after the starting point
  double* __restrict__ aty = wsi + (long long)W_ATY * n;
  const double* __restrict__ c = a.c + (long long)inst * a.c_stride;

#ifdef INTOPT_PROFILE
  PhaseTimer pt{a.phase_cycles, 0};
#else
  PhaseTimer pt{nullptr, 0};
#endif
  PHASE_START(pt);
  if (a.um.in) {  // the caller's cost vector -> this instance's standard-form row (kept for the backward)
    double* cw = const_cast<double*>(a.c) + (long long)inst * a.c_stride;
    NV_FOR(j, k) {
      double ck = 0.0;
      if (k < n) {
        const int src = a.um.in_idx[k];
        ck = src >= 0 ? (a.um.in_scale ? a.um.in_scale[k] : 1.0) * um_load(a.um, inst, src) : 0.0;
        cw[k] = ck;
      }
      if (TM) nv.st(V_C, j, k, ck);
    }
    CTA_SYNC();
  } else if (TM) {
    NV_FOR(j, k) nv.st(V_C, j, k, k < n ? c[k] : 0.0);
  }
  const bool l_in_smem = a.factor_in_smem != 0;
  if (cold) cold_copy_begin(cx, P, a.cold_factor, L, S.mbar, l_in_smem);
  // ---- load b, initial point (_initialization :647-661); padded slots of m-vectors are zero ---------
  CTA_FOR(i, mv) {
    double bv = 0.0;
    if (i < m) bv = a.b_batch ? a.b_batch[(long long)inst * a.b_stride + P.perm[i]] : a.b_problem[i];
    bs[i] = bv;
    y[i] = 0.0;
    r1[i] = 0.0;
    q[i] = 0.0;
    v[i] = 0.0;
  }
  if (cx.tid == 0) G[gpad(n)] = 0.0;
  double tau = 1.0, kappa = 1.0;
  const long long slen = 2LL * n + m + 2;
  CTA_SYNC();
  if (a.init) {
    const double* s0 = a.init + (long long)inst * slen;
    NV_FOR(j, k) {
      double xk = 1.0, tk = 1.0;
      if (k < n) {
        xk = s0[k];
        tk = s0[n + m + k];
        gstore(G, k, xk);
      }
      nv.st(V_X, j, k, xk);
      nv.st(V_T, j, k, tk);
    }
    CTA_FOR(i, m) y[i] = s0[n + P.perm[i]];
    tau = s0[2 * n + m];
    kappa = s0[2 * n + m + 1];
    CTA_SYNC();
    CTA_FOR(k, n) aty[k] = col_dot(P, M, k, y);
  } else {
    NV_FOR(j, k) {
      if (k < n) {
        aty[k] = 0.0;       // y = 0
        gstore(G, k, 1.0);  // G mirrors x at every loop top: the residual A x gathers from shared memory
      }
      nv.st(V_X, j, k, 1.0);
      nv.st(V_T, j, k, 1.0);
    }
  }
  nv.st_wait();
  CTA_SYNC();

  // ---- cold-start residual norms for the indicators (:884,907-910) -----------------------------
  double rp0, rd0, rg0;
  int nonfinite = 0;
  {
    double acc[3] = {0.0, 0.0, 0.0};
    NV_FOR(j, k) {
      typename NV::Val cr = nv.ld(V_C, j, k);
      nv.ld_wait();
      const double ck = nv.get(cr);
      if (k < n) {
        const double e = ck - 1.0;
        acc[0] = fma(e, e, acc[0]);
        acc[1] += ck;
        if (!(fabs(ck) < 1e300)) acc[2] += 1.0;
      }
    }
    double accp = 0.0;  // ||b - A 1||
    if (P.a_unit) {  // row sum of a 0/1 matrix = its entry count (exactly what summing the ones gives)
      CTA_FOR(i, m) {
        const double e = bs[i] - (double)(M.csr_ptr[i + 1] - M.csr_ptr[i]);
        accp = fma(e, e, accp);
      }
    } else {
      WARP_ROWS(i, m) {
        double s = 0.0;
        LANE_FOR(e, M.csr_ptr[i], M.csr_ptr[i + 1]) s += M.csr_val[e];
        s = warp_sum(s);
        if (cx.lane == 0) {
          const double e = bs[i] - s;
          accp = fma(e, e, accp);
        }
      }
    }
    double red[4] = {acc[0], acc[1], acc[2], accp};
    cta_sum<4>(cx, red);
    rd0 = fmax(1.0, sqrt(red[0]));
    rg0 = fmax(1.0, fabs(1.0 + red[1]));
    nonfinite = red[2] > 0.0;
    rp0 = fmax(1.0, sqrt(red[3]));
  }

  int iters = 0, status = ST_OK;
  bool solved = true;
  if (nonfinite) status = ST_NONFINITE;  // the reference raises LinAlgError here (:1332-1341)

  if (cold) cold_copy_finish(cx, P, a.cold_factor, L, S.mbar, l_in_smem, cold_parity);
  PHASE_MARK(pt, 0);
  while (status != ST_NONFINITE) {
    // ---- residuals at the current iterate (:765-770; the same r_p, r_d as _indicators :887-894) ----
    double acc[3] = {0.0, 0.0, 0.0};
    NV_FOR(j, k) {
      typename NV::Val xr = nv.ld(V_X, j, k), tr = nv.ld(V_T, j, k), cr = nv.ld(V_C, j, k);
      const bool ok = k < n;
      const double av = ok ? aty[k] : 0.0;
      nv.ld_wait();
      const double xk = nv.get(xr), tk = nv.get(tr), ck = nv.get(cr);
      const double r = -(av + tk - ck * tau);  // :766
      const double dk = xk / tk;               // :772
      nv.st(V_R2, j, k, r);
      nv.st(V_D, j, k, dk);
      if (ok) {
        if (TM) Dg[k] = dk;
        acc[0] = fma(xk, tk, acc[0]);
        acc[1] = fma(ck, xk, acc[1]);
        acc[2] = fma(r, r, acc[2]);
      }
    }
    nv.st_wait();
    if (iters > 0 && a.thr > 0.0) {
      // Early stop (:1408): with thr > 0 the loop usually ends because mu <= thr, which makes `go` false whatever the
      // residuals are.  If moreover tau is not tiny, both infeasibility tests (:1409-1413) are false without looking
      // at the residuals either, so the primal residual (a product with A) need not be formed on the way out.  mu comes
      // from the same fixed-order reduction as below: the decision is bit-identical to the full test sequence.
      double xt1[1] = {acc[0]};
      cta_sum<1>(cx, xt1);
      const double mu_e = (xt1[0] + tau * kappa) / (double)(n + 1);
      if (!(tau < a.tol * fmax(1.0, kappa)) && !(mu_e > a.thr)) {
        if (a.max_iter < iters)
          status = ST_MAXITER;
        else if (!solved)
          status = ST_BREAKDOWN;
        break;
      }
    }
    double accm[2] = {0.0, 0.0};
    PHASE_MARK(pt, 16);
    if (iters == 0 && !a.init && P.a_unit) {
      // cold start of a 0/1 matrix: x = 1, so A x is the entry count of every row (exactly what the product returns)
      CTA_FOR(i, m) r1[i] = (double)(M.csr_ptr[i + 1] - M.csr_ptr[i]);
    } else {
      spmv_rows(cx, P, M, G, r1);  // A x (G mirrors x)
    }
    CTA_SYNC();
    PHASE_MARK(pt, 17);
    CTA_FOR(i, m) {
      const double r = -(r1[i] - bs[i] * tau);  // :765
      r1[i] = r;
      accm[0] = fma(r, r, accm[0]);
      accm[1] = fma(bs[i], y[i], accm[1]);
    }
    double red[5] = {acc[0], acc[1], acc[2], accm[0], accm[1]};
    cta_sum<5>(cx, red);  // (its barriers also order the G writes before the assembly reads)
    const double xt_ = red[0], cx_ = red[1], nd2 = red[2], np2 = red[3], by_ = red[4];
    const double mu = (xt_ + tau * kappa) / (double)(n + 1);  // :770
    if (iters > 0) {
      // _indicators (:880-917) and the loop tests (:1408-1423), in the reference's order
      const double rho_p = sqrt(np2) / rp0;
      const double rho_d = sqrt(nd2) / rd0;
      const double rho_A = fabs(cx_ - by_) / (tau + fabs(by_));
      const double rho_g = fabs(kappa + cx_ - by_) / rg0;
      const double rho_mu = mu;  // mu_0 = 1
      const bool go = (rho_p > a.tol || rho_d > a.tol || rho_A > a.tol) && (mu > a.thr);
      const bool inf1 = rho_p < a.tol && rho_d < a.tol && rho_g < a.tol && tau < a.tol * fmax(1.0, kappa);
      const bool inf2 = rho_mu < a.tol && tau < a.tol * fmin(1.0, kappa);
      if (inf1 || inf2) {
        status = ST_INFEAS;
        break;
      }
      if (a.max_iter < iters) {
        status = ST_MAXITER;
        break;
      }
      if (!solved) {
        status = ST_BREAKDOWN;
        break;
      }
      if (!go) break;
    }
    ++iters;
    const double r3 = -(-cx_ + by_ - kappa);  // :767
    // pass-0 centring (:1392-1394): 0 with the predictor-corrector, beta * mean(x t) without; x*t - 0 and 1*r are exact,
    // so the predictor-corrector path computes the same bits as before
    const double gamma0 = PLAIN ? a.beta * (xt_ / (double)n) : 0.0;
    const double eta0 = 1.0 - gamma0;
    const double gm0 = gamma0 * mu;
    PHASE_MARK(pt, 1);

    // ---- M = A D A^T + delta I, factor (:772-777) ------------------------------------------------
    const bool reuse_cold = cold && iters == 1;  // x = t = 1: L already holds chol(A A^T + delta I)
    // The failure ladder of _get_delta (:798-822).  When cho_factor fails the reference retries the SAME system with
    // scipy.linalg.solve(assume_a='sym'), then 'pos', and only then reports "not solved" (:855-857).  A non-positive pivot
    // here means M = A D A^T + delta I is numerically singular (rank-deficient A with tiny damping, or D spanning > 1e16
    // near convergence); the device analogue of the ladder is two more attempts at this point with the diagonal raised by
    // eps * dmax, eps = 1e-12 then 1e-8, dmax the largest diagonal entry of M: the smallest perturbations that make the
    // pivots positive again, far below the accuracy a step needs.  The whole block (assembly, right-hand sides,
    // factorisation) simply runs again: one copy of the code, nothing added to the first pass.
    int attempt = 0;
  factor_again:
    if (!reuse_cold) {
#ifdef INTOPT_PROFILE
      assemble_normal_matrix(cx, P, M, Dg, G, S.scan, a.damping, L, a.debug);
#else
      assemble_normal_matrix(cx, P, M, Dg, G, S.scan, a.damping, L);
#endif
      if (attempt > 0) {
        double mx[1] = {0.0};
        CTA_FOR(e, P.n_asm) {
          const int pos = P.asm_pos[e];
          if (pos >= 0 && P.asm_diag[e]) mx[0] = fmin(mx[0], -L[pos]);
        }
        cta_min<1>(cx, mx);
        const double boost = (attempt == 1 ? 1e-12 : 1e-8) * -mx[0];
        CTA_FOR(e, P.n_asm) {
          const int pos = P.asm_pos[e];
          if (pos >= 0 && P.asm_diag[e]) L[pos] += boost;
        }
        CTA_SYNC();
      }
    }
    PHASE_MARK(pt, 2);
    // right-hand sides of the (p,q) solve (:800) and of the predictor (gamma = 0, eta = 1): G = D*c and, in the
    // upper half of the gather region (only the assembly's prefix sums use it), G2 = D*w
    double* G2 = G + gather_len(n);
    NV_FOR(j, k) {
      typename NV::Val dr = nv.ld(V_D, j, k), cr = nv.ld(V_C, j, k), xr = nv.ld(V_X, j, k), tr = nv.ld(V_T, j, k),
                       rr = nv.ld(V_R2, j, k);
      nv.ld_wait();
      const double dk = nv.get(dr), ck = nv.get(cr), xk = nv.get(xr), tk = nv.get(tr), rk = nv.get(rr);
      double wk;
      if (!PLAIN) {
        const double rxt = -(xk * tk);    // :788 with gamma = 0
        wk = rk - (1.0 / xk) * rxt;       // :801 with eta(0) = 1
      } else {
        const double rxt = -(xk * tk - gm0);
        wk = eta0 * rk - (1.0 / xk) * rxt;
      }
      nv.st(V_W, j, k, wk);
      if (k < n) {
        gstore(G, k, dk * ck);
        gstore(G2, k, dk * wk);
      }
    }
    nv.st_wait();
    CTA_SYNC();  // G, G2 complete
    // The factorisation is a dependent chain that keeps few warps busy, and the two products A(D c), A(D w) do not
    // depend on it: with enough warps the CTA splits, the last RHS_WARPS warps form the right-hand sides while the
    // others factor (each group on its own named barrier).
    if (reuse_cold) {
      solved = true;
      spmv_rows_n<2>(cx, P, M, G, G2, q, v);
      CTA_SYNC();
      CTA_FOR(i, m) {
        q[i] = bs[i] + q[i];  // r = r2 + A(Dinv r1), :666 with (r1,r2) = (c,b)
        v[i] = (PLAIN ? eta0 * r1[i] : r1[i]) + v[i];  // :785,666
      }
    } else {
      factor_and_rhs(cx, P, M, L, S, G, G2, q, v, &pt);
      solved = *S.flag == 0;
      if (!solved && attempt < 2) {
        ++attempt;
        goto factor_again;
      }
      CTA_FOR(i, m) {
        q[i] = bs[i] + q[i];
        v[i] = (PLAIN ? eta0 * r1[i] : r1[i]) + v[i];  // :785,666
      }
    }
    CTA_SYNC();
    PHASE_MARK(pt, 3);
    if (!solved) {  // deltas = 0 (:855-857): the iterate is unchanged, the tests run at loop top
      NV_FOR(j, k) {  // G must mirror x again
        typename NV::Val xr = nv.ld(V_X, j, k);
        nv.ld_wait();
        const double xk = nv.get(xr);
        if (k < n) gstore(G, k, xk);
      }
      CTA_SYNC();
      continue;
    }
    PHASE_MARK(pt, 4);
    {
      double* const rhs2[2] = {q, v};
      block_solve<2>(cx, P, M, L, rhs2, S.ybuf);
    }
    PHASE_MARK(pt, 5);
    double cp_, bq_, cu_, bv_;
    {
      double acc2[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      NV_FOR(j, k) {
        typename NV::Val dr = nv.ld(V_D, j, k), cr = nv.ld(V_C, j, k), wr = nv.ld(V_W, j, k);
        const bool ok = k < n;
        double atq = 0.0, atv = 0.0;
        if (ok) col_dot2(P, M, k, q, v, atq, atv);
        nv.ld_wait();
        const double dk = nv.get(dr), ck = nv.get(cr), wk = nv.get(wr);
        const double pk = dk * (atq - ck);  // :678 for (c,b)
        const double uk = dk * (atv - wk);  // :678 for the predictor
        nv.st(V_P, j, k, pk);
        nv.st(V_DX, j, k, uk);
        if (ok) {
          acc2[0] = fma(ck, pk, acc2[0]);
          acc2[2] = fma(ck, uk, acc2[2]);
          if (!(fabs(pk) < 1e300)) acc2[4] += 1.0;
        }
      }
      nv.st_wait();
      CTA_FOR(i, m) {
        acc2[1] = fma(bs[i], q[i], acc2[1]);
        acc2[3] = fma(bs[i], v[i], acc2[3]);
        if (!(fabs(q[i]) < 1e300)) acc2[4] += 1.0;
      }
      cta_sum<5>(cx, acc2);
      cp_ = acc2[0];
      bq_ = acc2[1];
      cu_ = acc2[2];
      bv_ = acc2[3];
      if (acc2[4] > 0.0) {  // NaN in p or q (:805-806) -> unsolved
        solved = false;
        NV_FOR(j, k) {  // G must mirror x again
          typename NV::Val xr = nv.ld(V_X, j, k);
          nv.ld_wait();
          const double xk = nv.get(xr);
          if (k < n) gstore(G, k, xk);
        }
        CTA_SYNC();
        continue;
      }
    }

    PHASE_MARK(pt, 6);
    // ---- predictor (pass 0) and corrector (pass 1) (:783-847) ---------------------------------------
    double gamma = gamma0, dtau = 0.0, dkappa = 0.0, minr = 0.0;
    for (int pass = 0; pass < 2; ++pass) {  // (:846-847: without the corrector the first direction is the step)
      const double eta = 1.0 - gamma;
      const double gm = gamma * mu;
      double rtk = -(tau * kappa - gm);     // :789
      if (pass == 1) rtk -= dtau * dkappa;  // :795
      if (pass == 1) {
        NV_FOR(j, k) {
          typename NV::Val xr = nv.ld(V_X, j, k), tr = nv.ld(V_T, j, k), rr = nv.ld(V_R2, j, k), dr = nv.ld(V_D, j, k),
                           er = nv.ld(V_DXDT, j, k);
          nv.ld_wait();
          const double xk = nv.get(xr), tk = nv.get(tr), rk = nv.get(rr), dk = nv.get(dr), ek = nv.get(er);
          const double rxt = -(xk * tk - gm) - ek;         // :788,793
          const double wk = eta * rk - (1.0 / xk) * rxt;  // :801
          nv.st(V_W, j, k, wk);
          if (k < n) gstore(G, k, dk * wk);
        }
        nv.st_wait();
        CTA_SYNC();
        PHASE_MARK(pt, 21);
        spmv_rows(cx, P, M, G, v);
        CTA_SYNC();
        PHASE_MARK(pt, 22);
        CTA_FOR(i, m) v[i] = eta * r1[i] + v[i];
        CTA_SYNC();
        PHASE_MARK(pt, 7);
        {
          double* const rhs1[1] = {v};
          block_solve<1>(cx, P, M, L, rhs1, S.ybuf);
        }
        PHASE_MARK(pt, 8);
        double acc3[2] = {0.0, 0.0};
        NV_FOR(j, k) {
          typename NV::Val dr = nv.ld(V_D, j, k), cr = nv.ld(V_C, j, k), wr = nv.ld(V_W, j, k);
          const bool ok = k < n;
          const double atv = ok ? col_dot(P, M, k, v) : 0.0;
          nv.ld_wait();
          const double dk = nv.get(dr), ck = nv.get(cr), wk = nv.get(wr);
          const double uk = dk * (atv - wk);  // :678
          nv.st(V_DX, j, k, uk);
          if (ok) acc3[0] = fma(ck, uk, acc3[0]);
        }
        nv.st_wait();
        CTA_FOR(i, m) acc3[1] = fma(bs[i], v[i], acc3[1]);
        cta_sum<2>(cx, acc3);
        cu_ = acc3[0];
        bv_ = acc3[1];
      }
      dtau = ((eta * r3 + 1.0 / tau * rtk - (-cu_ + bv_)) / (1.0 / tau * kappa + (-cp_ + bq_)));  // :837-838
      dkappa = (rtk - kappa * dtau) / tau;                                                         // :842
      // ratio test (:871-875): min over x/-dx and t/-dt of the decreasing components = 1 / max(-dx/x, -dt/t); the
      // reciprocals come from 1/x (needed for d_t anyway) and 1/t = D/x, so the loop has one division instead of three
      double mn[1] = {INFINITY};
      NV_FOR(j, k) {
        typename NV::Val xr = nv.ld(V_X, j, k), tr = nv.ld(V_T, j, k), ur = nv.ld(V_DX, j, k), pr = nv.ld(V_P, j, k),
                         dr = nv.ld(V_D, j, k), er = xr;
        if (pass == 1) er = nv.ld(V_DXDT, j, k);
        nv.ld_wait();
        const double xk = nv.get(xr), tk = nv.get(tr), uk = nv.get(ur), pk = nv.get(pr), dk = nv.get(dr);
        double rxt = -(xk * tk - gm);
        if (pass == 1) rxt -= nv.get(er);
        const double dxk = uk + pk * dtau;  // :839
        const double rx = 1.0 / xk;
        const double dtk = rx * (rxt - tk * dxk);  // :841
        nv.st(V_DX, j, k, dxk);
        if (pass == 0) nv.st(V_DXDT, j, k, dxk * dtk);
        if (k < n) {
          // -mn[0] tracks max(-dx/x, -dt/t) over the decreasing components (as a min of the negated values)
          if (dxk < 0.0) mn[0] = fmin(mn[0], dxk * rx);
          if (dtk < 0.0) mn[0] = fmin(mn[0], dtk * (dk * rx));
        }
      }
      nv.st_wait();
      cta_min<1>(cx, mn);
      minr = mn[0] < 0.0 ? -1.0 / mn[0] : INFINITY;
      PHASE_MARK(pt, 23);
      // _get_step (:870-878) with alpha0 = 1 (:843); an empty candidate set contributes 1
      double alpha = fmin(1.0, minr);
      if (dtau < 0.0) alpha = fmin(alpha, tau / -dtau);
      if (dkappa < 0.0) alpha = fmin(alpha, kappa / -dkappa);
      const double oma = 1.0 - alpha;
      gamma = oma * oma * fmin(0.1, oma);  // :844
      if (pass == 1 || PLAIN) {
        // ---- step (:1402-1403).  alpha0*min(x/-dx) == min(alpha0*x/-dx) bit for bit
        double al = fmin(1.0, a.alpha0 * minr);
        if (dtau < 0.0) al = fmin(al, a.alpha0 * tau / -dtau);  // evaluation order of :874,876
        if (dkappa < 0.0) al = fmin(al, a.alpha0 * kappa / -dkappa);
        NV_FOR(j, k) {
          typename NV::Val xr = nv.ld(V_X, j, k), tr = nv.ld(V_T, j, k), ur = nv.ld(V_DX, j, k), wr = nv.ld(V_W, j, k),
                           cr = nv.ld(V_C, j, k), er = xr;
          if (pass == 1) er = nv.ld(V_DXDT, j, k);
          const bool ok = k < n;
          const double av = ok ? aty[k] : 0.0;
          nv.ld_wait();
          const double xk = nv.get(xr), tk = nv.get(tr), dxk = nv.get(ur), wk = nv.get(wr), ck = nv.get(cr);
          const double ek = pass == 1 ? nv.get(er) : 0.0;  // Mehrotra term of the corrector only (:793)
          const double rxt = -(xk * tk - gm) - ek;
          const double rx = 1.0 / xk;
          const double dtk = rx * (rxt - tk * dxk);
          const double xn = xk + al * dxk;
          nv.st(V_X, j, k, xn);
          nv.st(V_T, j, k, tk + al * dtk);
          if (ok) {
            gstore(G, k, xn);
            // A^T d_y = A^T v + dtau A^T q, recovered from d_x = D (A^T v - w) + D (A^T q - c) dtau with 1/D = t/x
            aty[k] = av + al * (dxk * tk * rx + wk + dtau * ck);
          }
        }
        nv.st_wait();
        CTA_FOR(i, m) y[i] = y[i] + al * (v[i] + q[i] * dtau);  // d_y = v + q d_tau (:840)
        tau = tau + al * dtau;
        kappa = kappa + al * dkappa;
        CTA_SYNC();
        break;
      }
    }
    PHASE_MARK(pt, 9);
  }

  // ---- epilogue (:1441-1447): x_hat = x / tau, save the iterate ------------------------------------
  double* st = a.state + (long long)inst * slen;
  double* xh = a.xhat ? a.xhat + (long long)inst * n : nullptr;
  NV_FOR(j, k) {
    typename NV::Val xr = nv.ld(V_X, j, k), tr = nv.ld(V_T, j, k);
    nv.ld_wait();
    const double xk = nv.get(xr), tk = nv.get(tr);
    if (k < n) {
      if (xh) xh[k] = xk / tau;
      st[k] = xk;
      st[n + m + k] = tk;
    }
  }
  if (a.um.out) {  // x in the caller's layout and precision, presolve-fixed variables re-inserted (_postprocess :1241-1285)
    // (any element, not only this thread's: G mirrors x whenever the loop is left, and a barrier has passed since)
    CTA_FOR(i, a.um.n_out) {
      const int k = a.um.out_idx[i];
      um_store(a.um, inst, i, k >= 0 ? G[gpad(k)] / tau : (a.um.out_const ? a.um.out_const[i] : 0.0));
    }
  }
  CTA_FOR(i, m) st[n + P.perm[i]] = y[i];
  if (cx.tid == 0) {
    st[2 * n + m] = tau;
    st[2 * n + m + 1] = kappa;
    a.status[inst] = status;
    a.iters[inst] = iters;
  }
  CTA_SYNC();
  PHASE_MARK(pt, 10);
}

template <bool PLAIN>
IO_DEV void forward_instance_t(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const FwdParams& a, int inst, double* wsi,
                               const SmemLayout& S, bool cold = false, unsigned cold_parity = 0) {
  const int n = P.n;
  NVGlobal nv;
  nv.n = n;
  nv.v[V_X] = wsi + (long long)W_X * n;
  nv.v[V_T] = wsi + (long long)W_T * n;
  nv.v[V_C] = const_cast<double*>(a.c) + (long long)inst * a.c_stride;  // (never stored to)
  nv.v[V_D] = wsi + (long long)W_D * n;
  nv.v[V_R2] = wsi + (long long)W_R2 * n;
  nv.v[V_W] = wsi + (long long)W_W * n;
  nv.v[V_P] = wsi + (long long)W_P * n;
  nv.v[V_DX] = wsi + (long long)W_DX * n;
  nv.v[V_DXDT] = wsi + (long long)W_DXDT * n;
  forward_instance_nv<PLAIN>(cx, P, M, a, inst, wsi, S, nv, cold, cold_parity);
}

#ifndef INTOPT_HOST_EMU
// The predictor-corrector forward with the nine element-wise n-vectors in tensor memory (tmem_base: the CTA's allocation of
// NV_COLS columns per group of four warps).
IO_DEV void forward_instance_tmem(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const FwdParams& a, int inst,
                                  double* wsi, const SmemLayout& S, uint32_t tmem_base, bool cold, unsigned cold_parity) {
  NVTmem nv;
  nv.base = tmem_base + ((uint32_t)(32 * (cx.warp & 3)) << 16) + (uint32_t)(NV_COLS * (cx.warp >> 2));
  forward_instance_nv<false>(cx, P, M, a, inst, wsi, S, nv, cold, cold_parity);
}
#endif

#ifdef INTOPT_HOST_EMU  // (the CUDA kernels pick the instantiation at launch: one kernel per value of PLAIN)
IO_DEV void forward_instance(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const FwdParams& a, int inst, double* wsi,
                             const SmemLayout& S, bool cold = false, unsigned cold_parity = 0) {
  if (a.plain)
    forward_instance_t<true>(cx, P, M, a, inst, wsi, S, cold, cold_parity);
  else
    forward_instance_t<false>(cx, P, M, a, inst, wsi, S, cold, cold_parity);
}
#endif

// The factor every cold-started instance uses at its first iteration: chol(A A^T + delta I) in the block-skyline
// layout, computed by exactly the code path an instance would run (D = x/t = 1), hence bit-identical to it.
IO_DEV void cold_factor_build(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double delta, double* wsi,
                              const SmemLayout& S, double* out, int* ok_out) {
  const int n = P.n;
  double* D = wsi + (long long)W_D * n;
  CTA_FOR(k, n) D[k] = 1.0;
  if (cx.tid == 0) S.G[gpad(n)] = 0.0;
  CTA_SYNC();
  assemble_normal_matrix(cx, P, M, D, S.G, S.scan, delta, S.L);
  const bool ok = block_cholesky(cx, P, M, S.L, S.dscr, S.flag);
  CTA_FOR(e, P.bsize) out[e] = S.L[e];
  if (cx.tid == 0) *ok_out = ok ? 1 : 0;
  CTA_SYNC();
}

struct BwdParams {
  int B, mode, nvar;
  double damping;
  const double* state;
  const double* c;
  long long c_stride;
  const double* b_batch;
  long long b_stride;
  const double* b_problem;
  const double* gx;  // B x nvar
  UserMap um;        // optional user-layout input / output (see UserMap)
  const double* a_vals;  // per-instance values of A or null
  long long a_stride;
  double* vals_ws;
  long long vals_ws_stride;
  double* gc;        // B x nvar
  int* status;
  double* ws;
  long long ws_stride;
  int* counter;
  int factor_in_smem, gather_in_smem;
};

// c_grad[var_index] = dc, zeros elsewhere (:1559-1560).  With a mapped output the result goes straight to the caller's
// layout and precision and the fp64 staging row grad_c_dev is NOT written (27 MB per Energy batch saved); without one it
// goes to grad_c_dev.  `val(k)` evaluates entry k of dc.
template <class F>
IO_DEV void backward_emit(const CtaCtx& cx, const BwdParams& a, int inst, double* gc, F val) {
  if (a.um.out) {
    CTA_FOR(i, a.um.n_out) {
      const int k = a.um.out_idx[i];
      um_store(a.um, inst, i, k >= 0 ? val(k) : 0.0);
    }
  } else {
    CTA_FOR(k, a.nvar) gc[k] = val(k);
  }
}

// K5: dL/dc at the saved iterate, VJP form:
//   whole: w = M^-1 (tau A(X2 g)), ug = X2 (A^T w - tau g)/mu, q = M^-1 (mu b + A(X2 c)), p = X2 (A^T q - c)/mu,
//          s = (x.g + c.ug - b.w)/(b.q - c.p + mu/tau^2), dc = ug + p s        (ip_model_whole.py:1515-1556)
//   alter: dc = -X2 (g - A^T M^-1 A (X2 g)) / mu  with x,t normalised by tau   (ip_model_alter.py:1524-1620)
// with M = A diag(x^2) A^T + damping I in both.
// (its five element-wise n-vectors - x^2, g, c, u_g, p - through the same accessor as the forward's: workspace or tensor
// memory)
enum { B_X2 = 0, B_G, B_C, B_UG, B_PP };
template <class NV>
IO_DEV void backward_instance_nv(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const BwdParams& a, int inst, double* wsi,
                                 const SmemLayout& S, const NV& nv) {
  constexpr bool TM = NV::kTmem;
  if (a.a_vals)
    fill_instance_values(cx, P, a.a_vals + (long long)inst * a.a_stride, a.vals_ws + (long long)blockIdx_x() * a.vals_ws_stride);
  const int m = P.m, n = P.n, nvar = a.nvar, mv = P.mv;
  const int nv_trips = (n + cx.nthreads - 1) / cx.nthreads;
  double *L = S.L, *G = S.G, *bs = S.bs, *q = S.q, *v = S.v;
  double* x2 = wsi + (long long)W_D * n;  // (also with tensor memory: the assembly's prefix sums read it in their own order)
  const long long slen = 2LL * n + m + 2;
  const double* st = a.state + (long long)inst * slen;
  const double* __restrict__ c = a.c + (long long)inst * a.c_stride;
  const double* __restrict__ g = a.gx + (long long)inst * nvar;
  double* gc = a.gc + (long long)inst * nvar;
  if (a.um.in) {  // the caller's upstream gradient -> del_x[var_index] (:1518)
    double* gw = const_cast<double*>(a.gx) + (long long)inst * nvar;
    CTA_FOR(k, nvar) gw[k] = um_load(a.um, inst, a.um.in_idx[k]);
    CTA_SYNC();
  }
  const double* xs = st;
  const double* ts = st + n + m;
  const double tau = st[2 * n + m], kappa = st[2 * n + m + 1];
  const bool whole = a.mode == 0;

  CTA_FOR(i, mv) {
    double bv = 0.0;
    if (i < m)
      bv = a.b_batch ? a.b_batch[(long long)inst * a.b_stride + P.perm[i]] : (a.b_problem ? a.b_problem[i] : 0.0);
    bs[i] = bv;
    q[i] = 0.0;
    v[i] = 0.0;
  }
  if (cx.tid == 0) G[gpad(n)] = 0.0;
  double xt[2] = {0.0, 0.0};
  NV_FOR(j, k) {
    double xx = 1.0, gk = 0.0, ck = 0.0;
    if (k < n) {
      const double xk = whole ? xs[k] : xs[k] / tau;  // alter saves x/tau, t/tau (ip_model_alter.py:1459-1461)
      const double tk = whole ? ts[k] : ts[k] / tau;
      xx = xk * xk;
      xt[0] = fma(xk, tk, xt[0]);
      if (k < nvar) {
        gk = g[k];
        xt[1] = fma(xk, gk, xt[1]);
      }
      if (TM) {
        x2[k] = xx;
        ck = c[k];
      }
    }
    nv.st(B_X2, j, k, xx);
    if (TM) {
      nv.st(B_G, j, k, gk);
      nv.st(B_C, j, k, ck);
    }
  }
  nv.st_wait();
  cta_sum<2>(cx, xt);
  const double mu = whole ? (xt[0] + tau * kappa) / (double)(n + 1) : xt[0] / (double)n;  // :1517 | alter :1525
  const double xg = xt[1];

  CTA_SYNC();
  assemble_normal_matrix(cx, P, M, x2, G, S.scan, a.damping, L);  // :1525-1527,1535
  double* G2 = G + gather_len(n);  // upper half of the gather region: X2 c beside X2 g, one pass over A for both
  NV_FOR(j, k) {
    typename NV::Val xr = nv.ld(B_X2, j, k), cr = nv.ld(B_C, j, k), gr = xr;
    if (TM || k < nvar) gr = nv.ld(B_G, j, k);  // (g has nvar entries; its tensor-memory copy is zero beyond them)
    nv.ld_wait();
    const double xx = nv.get(xr), gk = nv.get(gr), ck = nv.get(cr);
    if (k < n) {
      gstore(G, k, (k < nvar) ? xx * gk : 0.0);
      if (whole) gstore(G2, k, xx * ck);
    }
  }
  CTA_SYNC();  // G, G2 complete
  factor_and_rhs(cx, P, M, L, S, G, whole ? G2 : nullptr, v, whole ? q : nullptr, nullptr);  // A (X2 g_hat) [, A (X2 c)]
  const bool ok = *S.flag == 0;
  if (!ok) {
    // the reference raises LinAlgError (:1546-1547); here: zero gradient + status
    backward_emit(cx, a, inst, gc, [](int) { return 0.0; });
    if (cx.tid == 0) a.status[inst] = ST_BREAKDOWN;
    CTA_SYNC();
    return;
  }
  if (!whole) {
    double* const rhs1[1] = {v};
    block_solve<1>(cx, P, M, L, rhs1, S.ybuf);
    at_mult<1>(cx, P, M, v, nullptr, G, nullptr);
    CTA_SYNC();
    backward_emit(cx, a, inst, gc, [&](int k) { return -x2[k] * (g[k] - G[k]) / mu; });
    if (cx.tid == 0) a.status[inst] = ST_OK;
    CTA_SYNC();
    return;
  }
  CTA_FOR(i, m) {
    v[i] = tau * v[i];            // r1 g = tau A X2 g (:1529)
    q[i] = mu * bs[i] + q[i];     // r2 = mu b + A(X2 c) (:1530)
  }
  CTA_SYNC();
  {
    double* const rhs2[2] = {q, v};
    block_solve<2>(cx, P, M, L, rhs2, S.ybuf);
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  at_mult<2>(cx, P, M, q, v, G, G + gather_len(n));
  CTA_SYNC();
  NV_FOR(j, k) {
    typename NV::Val xr = nv.ld(B_X2, j, k), cr = nv.ld(B_C, j, k), gr = xr;
    if (TM || k < nvar) gr = nv.ld(B_G, j, k);
    const bool in = k < n;
    const double atq = in ? G[k] : 0.0, atv = in ? G[gather_len(n) + k] : 0.0;
    nv.ld_wait();
    const double xx = nv.get(xr), ck = nv.get(cr);
    const double gk = (TM || k < nvar) ? nv.get(gr) : 0.0;
    const double ugk = xx * (atv - tau * gk) / mu;  // :1549 applied to g
    const double pk = xx * (atq - ck) / mu;         // :1548
    nv.st(B_UG, j, k, ugk);
    nv.st(B_PP, j, k, pk);
    if (in) {
      acc[0] = fma(ck, ugk, acc[0]);
      acc[1] = fma(ck, pk, acc[1]);
    }
  }
  nv.st_wait();
  CTA_FOR(i, m) {
    acc[2] = fma(bs[i], v[i], acc[2]);
    acc[3] = fma(bs[i], q[i], acc[3]);
  }
  cta_sum<4>(cx, acc);  // (its barriers: every thread is past its reads of G)
  const double s = (xg + acc[0] - acc[2]) / (acc[3] - acc[1] + mu / (tau * tau));  // :1552-1553
  if (TM) {
    // dc = u_g + p s (:1554-1556) is formed by the owner of each element and passed through the gather region: the
    // emission may ask for any element
    NV_FOR(j, k) {
      typename NV::Val ur = nv.ld(B_UG, j, k), pr = nv.ld(B_PP, j, k);
      nv.ld_wait();
      const double dc = nv.get(ur) + nv.get(pr) * s;
      if (k < n) G[k] = dc;
    }
    CTA_SYNC();
    backward_emit(cx, a, inst, gc, [&](int k) { return G[k]; });
  } else {
    const double* ug = wsi + (long long)W_DX * n;
    const double* pp = wsi + (long long)W_P * n;
    backward_emit(cx, a, inst, gc, [&](int k) { return ug[k] + pp[k] * s; });  // :1554-1556
  }
  if (cx.tid == 0) a.status[inst] = ST_OK;
  CTA_SYNC();
}

IO_DEV void backward_instance(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const BwdParams& a, int inst, double* wsi,
                              const SmemLayout& S) {
  const int n = P.n;
  NVGlobal nv;
  nv.n = n;
  for (int sidx = 0; sidx < V_SLOTS; ++sidx) nv.v[sidx] = nullptr;
  nv.v[B_X2] = wsi + (long long)W_D * n;
  nv.v[B_G] = const_cast<double*>(a.gx) + (long long)inst * a.nvar;  // (never stored to)
  nv.v[B_C] = const_cast<double*>(a.c) + (long long)inst * a.c_stride;  // (never stored to)
  nv.v[B_UG] = wsi + (long long)W_DX * n;
  nv.v[B_PP] = wsi + (long long)W_P * n;
  backward_instance_nv(cx, P, M, a, inst, wsi, S, nv);
}

#ifndef INTOPT_HOST_EMU
IO_DEV void backward_instance_tmem(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const BwdParams& a, int inst,
                                   double* wsi, const SmemLayout& S, uint32_t tmem_base) {
  NVTmem nv;
  nv.base = tmem_base + ((uint32_t)(32 * (cx.warp & 3)) << 16) + (uint32_t)(NV_COLS * (cx.warp >> 2));
  backward_instance_nv(cx, P, M, a, inst, wsi, S, nv);
}
#endif

}  // namespace intopt
