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
#define IO_DEV __device__ __forceinline__
#define IO_HD __host__ __device__ inline
#define IO_FORCEINLINE __device__ __forceinline__
struct CtaCtx {
  int tid, nthreads, lane, warp, nwarps;
  static constexpr int warpsize = 32;
  double* red;  // 8*32 doubles of shared scratch
};
#define CTA_SYNC() __syncthreads()
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
constexpr int KSEG = 4;  // == MAXSEG of intopt_structure.hpp (checked in intopt_bind.hpp)
constexpr int KREC = 12 + 4 * KSEG;  // == ST_REC: ints per step record of the substitution schedule

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
  const int* g_act_off;
  // concurrent leading segments (intopt_structure.hpp): nseg == 0 disables
  int nseg, seg_end, seg_tmax, nbord;
  int seg_start[KSEG], seg_len[KSEG];
  const int* g_act_ns;
  const int* g_seg_bord;
  // substitution schedule (intopt_structure.hpp: st_fwd / st_bwd, one KREC-int record per step)
  int nsteps, n_wblk;
  const int* g_st_fwd;
  const int* g_st_bwd;
  const int* g_wblk_col;  // column K of W slot w (n_wblk entries)
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
  int asm_unit;
};

// The small schedule / pointer arrays that sit on the critical path of every blocked step.  They point into
// shared memory when staged (stage_meta) and into global memory otherwise.
struct Meta {
  const int* fb;
  const int* boff;
  const int* act_ptr;
  const int* act_blk;
  const int* act_off;
  const int* csr_ptr;
  const int* csc_gptr;
  const int* asm_gptr;
  const int* asm_rptr;
  const int* act_ns;
  const int* seg_bord;
  const int* st_fwd;
  const int* st_bwd;
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
  const int ngrp = (dbg & 1) ? 0 : (P.n_asm >> 5);
  if (P.asm_run) {
    D2* PP = reinterpret_cast<D2*>(region);
    dd_prefix_scan(cx, d, n, PP, scan_tmp);  // scan_tmp may alias L (scan_in_factor): the zero fill comes after
    CTA_FOR(e, P.bsize) L[e] = 0.0;
    CTA_SYNC();
    CTA_FOR(e, P.n_dummy) L[P.dummy_pos[e]] = 1.0;
    // Lane l of a warp owns entry 32*grp + l; its R run headers are contiguous and fetched 12 at a time with
    // three 16-byte loads that are all in flight together (one L2 round trip per group).
    WARP_ROWS(grp, ngrp) {
      const int r0 = M.asm_rptr[grp], R = M.asm_rptr[grp + 1] - r0;
      LANE_FOR(l, 0, 32) {
        const uint32_t* __restrict__ hp = P.asm_run + (size_t)r0 * 32 + (size_t)l * R;
        double s0 = 0.0, s1 = 0.0;
        for (int rb = 0; rb < R; rb += 12) {
          uint32_t h[12];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (rb + 4 * c < R) {
              const U4 q = *reinterpret_cast<const U4*>(hp + rb + 4 * c);
              h[4 * c] = q.x;
              h[4 * c + 1] = q.y;
              h[4 * c + 2] = q.z;
              h[4 * c + 3] = q.w;
            } else {
              h[4 * c] = h[4 * c + 1] = h[4 * c + 2] = h[4 * c + 3] = 0u;
            }
          }
#pragma unroll
          for (int u = 0; u < 12; u += 2) {  // padding headers (len 0) contribute exactly 0
            const int a0 = (int)(h[u] & 0xffffu), a1 = (int)(h[u + 1] & 0xffffu);
            const D2 lo0 = PP[a0], hi0 = PP[a0 + (int)(h[u] >> 16)];
            const D2 lo1 = PP[a1], hi1 = PP[a1 + (int)(h[u + 1] >> 16)];
            s0 += dd_diff(hi0, lo0);
            s1 += dd_diff(hi1, lo1);
          }
        }
        const int e = grp * 32 + l, pos = P.asm_pos[e];
        if (pos >= 0) L[pos] = P.asm_diag[e] ? (s0 + s1) + delta : (s0 + s1);
      }
    }
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
                       : gather_sum_weighted(P.asm_k16 + base, P.asm_w + base, g, len);
      } else {
        s = P.asm_unit ? gather_sum_unit(P.asm_k + base, g, len)
                       : gather_sum_weighted(P.asm_k + base, P.asm_w + base, g, len);
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
#ifdef INTOPT_HOST_EMU
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
IO_DEV void diag_block_factor_invert(const CtaCtx& cx, double* blk, double* scr, int* fail,
                                     const double* panel = nullptr) {
  const unsigned full = 0xffffffffu;
  const int i = cx.lane & 7;
  double a[KB], rinv[KB];
  load_row(blk, i, a);
  if (panel) {  // fused look-ahead update in registers: row i of C -= P_i . P_k^T with P = L_(K+1,K)
    double own[KB];
    load_row(panel, i, own);
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      double pk[KB];
      load_row(panel, k, pk);  // same address on every lane: broadcast
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < KB; ++j) s = fma(own[j], pk[j], s);
      a[k] -= s;
    }
  }
  bool bad = false;
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    double piv = __shfl_sync(full, a[j], j);
    if (!(piv > 0.0) || !(piv < 1e300)) {
      bad = true;
      piv = 1.0;
    }
    const double r = rsqrt(piv);
    rinv[j] = r;
    a[j] = (i == j) ? piv * r : a[j] * r;  // column j of L (rows above the diagonal hold junk)
#pragma unroll
    for (int k = j + 1; k < KB; ++k) {
      const double lkj = __shfl_sync(full, a[j], k);
      a[k] = fma(-a[j], lkj, a[k]);
    }
  }
  if (bad && cx.lane == 0) *fail = 1;
  // publish L (row i, plain row-major) so that every lane can read the rows below its own column
  double* Ls = scr + KB;
  if (cx.lane < KB) {
#pragma unroll
    for (int k = 0; k < KB; ++k) Ls[i * KB + k] = a[k];
  }
  __syncwarp();
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

// ---- concurrent leading segments ------------------------------------------------------------------------
// Per-thread view of step t of the segment phase: column K[s] = seg_start[s] + t of every segment that is still
// running (na[s] == 0 and K[s] == -1 otherwise).  Everything is indexed with compile-time s (registers).
struct SegStep {
  int K[KSEG], a0[KSEG], na[KSEG], ns[KSEG], ld[KSEG];
};
IO_FORCEINLINE SegStep seg_step(const ProblemDev& P, const Meta& M, int t) {
  SegStep g;
#pragma unroll
  for (int s = 0; s < KSEG; ++s) {
    const bool on = s < P.nseg && t < P.seg_len[s];
    const int K = on ? P.seg_start[s] + t : 0;
    g.K[s] = on ? K : -1;
    g.a0[s] = M.act_ptr[K];
    g.na[s] = on ? M.act_ptr[K + 1] - g.a0[s] : 0;
    g.ns[s] = on ? M.act_ns[K] : 0;  // active rows inside the segment (they come first in the list)
    g.ld[s] = (g.ns[s] > 0 && M.act_blk[g.a0[s]] == K + 1) ? 1 : 0;  // block (K+1,K) exists: look-ahead pivot
  }
  return g;
}
#define SEG_PICK(arr, s) ((s) == 0 ? (arr)[0] : (s) == 1 ? (arr)[1] : (s) == 2 ? (arr)[2] : (arr)[3])
#define SEG_PICK5(arr, s) ((s) == 0 ? (arr)[0] : (s) == 1 ? (arr)[1] : (s) == 2 ? (arr)[2] : (s) == 3 ? (arr)[3] : (arr)[4])
#define SEG_PICK4(arr, s, f) ((s) == 0 ? (arr)[0].f : (s) == 1 ? (arr)[1].f : (s) == 2 ? (arr)[2].f : (arr)[3].f)

// acc += rows 4tr..4tr+3 of L_IK times rows 4tc..4tc+3 of L_JK, transposed (one 4x4 tile of L_IK L_JK^T)
IO_FORCEINLINE void chol_tile_acc(const Meta& M, const double* L, int I, int J, int K, int tr, int tc, double (&acc)[4][4]) {
  const double* A = L + M.boff[I] + (K - M.fb[I]) * 64 + tr * 32;
  const double* Bm = L + M.boff[J] + (K - M.fb[J]) * 64 + tc * 32;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
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
}
IO_FORCEINLINE void chol_tile_sub(const Meta& M, double* L, int I, int J, int tr, int tc, const double (&acc)[4][4]) {
  double* dst = L + M.boff[I] + (J - M.fb[I]) * 64 + tr * 32;
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

// one panel row: row `rr` of block (I,K) times Linv_KK^T, in place
IO_FORCEINLINE void chol_panel_row(const Meta& M, double* L, int I, int K, int rr) {
  const double* Linv = L + M.boff[K] + (K - M.fb[K]) * 64;
  double* blk = L + M.boff[I] + (K - M.fb[I]) * 64;
  double cin[KB], out[KB];
  load_row(blk, rr, cin);
#pragma unroll
  for (int c = 0; c < KB; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k <= c) s = fma(cin[k], Linv[bidx(c, k)], s);
    out[c] = s;
  }
  store_row(blk, rr, out);
}

// Segment phase of the factorisation: step t handles column seg_start[s] + t of every segment at once.  Warp s
// factors the segment's next diagonal block (look-ahead, its update by the fresh panel block fused in) while the
// remaining warps apply the trailing updates: targets inside a segment or in (tail row x segment column) belong to
// one column; (tail x tail) targets sum the contributions of all running segments in fixed order (deterministic).
IO_DEV void chol_segment_phase(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double* L, double* scratch,
                               int* fail_flag, int dbg) {
  double* myscr = scratch + (cx.warp < KSEG ? cx.warp : 0) * (KB + KB * KB);
  for (int s = cx.warp; s < P.nseg; s += cx.nwarps) {
    const int K = P.seg_start[s];
    diag_block_factor_invert(cx, L + M.boff[K] + (K - M.fb[K]) * 64, myscr, fail_flag);
  }
  CTA_SYNC();
  const int nbp = P.nbord * (P.nbord + 1) / 2;
  for (int t = 0; t < P.seg_tmax; ++t) {
    const SegStep g = seg_step(P, M, t);
    // (a) panels of every running column
    int poff[KSEG + 1];
    poff[0] = 0;
#pragma unroll
    for (int s = 0; s < KSEG; ++s) poff[s + 1] = poff[s] + g.na[s] * KB;
    if (!(dbg & 16)) CTA_FOR(idx, poff[KSEG]) {
      const int s = (idx >= poff[1]) + (idx >= poff[2]) + (idx >= poff[3]);
      const int loc = idx - SEG_PICK(poff, s);
      const int K = SEG_PICK(g.K, s);
      chol_panel_row(M, L, M.act_blk[SEG_PICK(g.a0, s) + (loc >> 3)], K, loc & 7);
    }
    CTA_SYNC();
    // (b) look-ahead pivots on the first warps, trailing updates on the others
    int nd = 0;  // running segments come first?  not necessarily: warp s serves segment s
#pragma unroll
    for (int s = 0; s < KSEG; ++s)
      if (g.ld[s]) nd = s + 1;
    const int nded = cx.nwarps > nd ? nd : 0;  // warps dedicated to the pivots (none if the CTA is that small)
    for (int s = cx.warp; s < nd; s += cx.nwarps) {
      if (SEG_PICK(g.ld, s) && !(dbg & 8)) {
        const int K = SEG_PICK(g.K, s);
        const double* lead_panel = L + M.boff[K + 1] + (K - M.fb[K + 1]) * 64;
        diag_block_factor_invert(cx, L + M.boff[K + 1] + (K + 1 - M.fb[K + 1]) * 64, myscr, fail_flag, lead_panel);
      }
    }
    if (cx.warp >= nded && !(dbg & 4)) {
      const int w_tid = cx.tid - nded * cx.warpsize, w_n = cx.nthreads - nded * cx.warpsize;
      // private targets of column K[s]: pairs (ia >= ib) of its active rows with ib inside the segment
      int toff[KSEG + 1];
      toff[0] = 0;
#pragma unroll
      for (int s = 0; s < KSEG; ++s) {
        const int npair = g.ns[s] * (g.ns[s] + 1) / 2 + (g.na[s] - g.ns[s]) * g.ns[s];
        toff[s + 1] = toff[s] + (npair - g.ld[s]) * 4;
      }
      const int ntask = toff[KSEG] + nbp * 4;
      for (int task = w_tid; task < ntask; task += w_n) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        if (task < toff[KSEG]) {
          const int s = (task >= toff[1]) + (task >= toff[2]) + (task >= toff[3]);
          const int loc = task - SEG_PICK(toff, s) + SEG_PICK(g.ld, s) * 4;
          const int pi = loc >> 2, tr = (loc >> 1) & 1, tc = loc & 1;
          const int ns = SEG_PICK(g.ns, s), a0 = SEG_PICK(g.a0, s), K = SEG_PICK(g.K, s);
          const int ntri = ns * (ns + 1) / 2;
          int ia, ib;
          if (pi < ntri) {
            ia = tri_row(pi);
            ib = pi - ia * (ia + 1) / 2;
          } else {
            const int r = pi - ntri;
            ia = ns + r / ns;
            ib = r - (ia - ns) * ns;
          }
          const int I = M.act_blk[a0 + ia], J = M.act_blk[a0 + ib];
          chol_tile_acc(M, L, I, J, K, tr, tc, acc);
          chol_tile_sub(M, L, I, J, tr, tc, acc);
        } else {
          const int loc = task - toff[KSEG];
          const int pi = loc >> 2, tr = (loc >> 1) & 1, tc = loc & 1;
          const int ia = tri_row(pi), ib = pi - ia * (ia + 1) / 2;
          const int I = M.seg_bord[ia], J = M.seg_bord[ib];
          const int fI = M.fb[I], fJ = M.fb[J];
          bool any = false;
#pragma unroll
          for (int s = 0; s < KSEG; ++s) {
            const int K = g.K[s];
            if (K >= 0 && fI <= K && fJ <= K) {
              chol_tile_acc(M, L, I, J, K, tr, tc, acc);
              any = true;
            }
          }
          if (any) chol_tile_sub(M, L, I, J, tr, tc, acc);
        }
      }
    }
    CTA_SYNC();
  }
}

IO_DEV void fuse_coupling_blocks(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double* L, double* W);

IO_DEV bool block_cholesky(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double* L, double* W, double* scratch,
                           int* fail_flag, PhaseTimer* pt = nullptr, int dbg = 0) {
  const int mb = P.mb;
  if (cx.tid == 0) *fail_flag = 0;
  CTA_SYNC();
  const int K0 = P.nseg > 1 ? P.seg_end : 0;  // first column of the sequential part
  if (P.nseg > 1) chol_segment_phase(cx, P, M, L, scratch, fail_flag, dbg);
  if (K0 < mb) {
    if (cx.warp == 0) diag_block_factor_invert(cx, L + M.boff[K0] + (K0 - M.fb[K0]) * 64, scratch, fail_flag);
    CTA_SYNC();
  }
  if (pt) { PHASE_MARK(*pt, 11); }
  const int o_tid = cx.nwarps > 1 ? cx.tid - cx.warpsize : 0;
  const int o_n = cx.nwarps > 1 ? cx.nthreads - cx.warpsize : 1;
  for (int K = K0; K < mb; ++K) {
    // (a) panel rows
    const int a0 = M.act_ptr[K], na = M.act_ptr[K + 1] - a0;
    if (!(dbg & 16)) CTA_FOR(idx, na * KB) chol_panel_row(M, L, M.act_blk[a0 + (idx >> 3)], K, idx & 7);
    CTA_SYNC();
    if (pt) { PHASE_MARK(*pt, 12); }
    // (b) trailing update with look-ahead
    const int np = na * (na + 1) / 2;
    const int lead = (na > 0 && M.act_blk[a0] == K + 1) ? 1 : 0;  // pair 0 updates block (K+1,K+1)
    if (cx.warp == 0) {
      if (K + 1 < mb && !(dbg & 8)) {
        // the update of the next diagonal block by this step's panel block (K+1,K) is fused into its factorisation
        const double* lead_panel = lead ? L + M.boff[K + 1] + (K - M.fb[K + 1]) * 64 : nullptr;
        diag_block_factor_invert(cx, L + M.boff[K + 1] + (K + 1 - M.fb[K + 1]) * 64, scratch, fail_flag, lead_panel);
      }
      if (pt) { PHASE_MARK(*pt, 13); }
    }
    if ((cx.nwarps == 1 || cx.warp > 0) && !(dbg & 4)) {
      for (int task = lead * 4 + o_tid; task < np * 4; task += o_n) chol_update_task(P, M, L, K, a0, task);
    }
    CTA_SYNC();
    if (pt) { PHASE_MARK(*pt, 14); }
  }
  const bool ok = *fail_flag == 0;
  if (ok && !(dbg & 32)) fuse_coupling_blocks(cx, P, M, L, W);
  return ok;
}

// ---------------------------------------------------------------------------------------------
// Solve L L^T z = r in place for NR right-hand sides at once (blocked substitution; the diagonal blocks hold
// Linv).  One barrier per step: while the CTA applies the step's update to the other rows, the pivot warp
// first updates the rows of the NEXT pivot block and immediately multiplies them by its inverted diagonal
// block (in place), so the next step's pivot values are ready when the barrier opens.  With concurrent
// leading segments a step handles one column of every segment (pivot warp s serves segment s).
// ---------------------------------------------------------------------------------------------
template <int NR>
IO_FORCEINLINE double* rhs_sel(double* const (&r)[NR], int q) {
  return (NR > 1 && q) ? r[NR - 1] : r[0];  // NR is 1 or 2
}

// r_blk <- Linv r_blk (or Linv^T r_blk), all NR right-hand sides, in place; executed by one full warp
template <int NR, bool TRANS>
IO_FORCEINLINE void pivot_inplace(const CtaCtx& cx, const double* Linv, double* const (&r)[NR], int blk) {
#ifdef INTOPT_HOST_EMU
  (void)cx;
  double tmp[NR * KB];
  for (int idx = 0; idx < NR * KB; ++idx) {
    const int q = idx >> 3, i = idx & 7;
    const double* rk = rhs_sel<NR>(r, q) + blk * KB;
    double s = 0.0;
    for (int j = 0; j < KB; ++j) {
      if (!TRANS && j <= i) s = fma(Linv[bidx(i, j)], rk[j], s);
      if (TRANS && j >= i) s = fma(Linv[bidx(j, i)], rk[j], s);
    }
    tmp[idx] = s;
  }
  for (int idx = 0; idx < NR * KB; ++idx) rhs_sel<NR>(r, idx >> 3)[blk * KB + (idx & 7)] = tmp[idx];
#else
  const int idx = cx.lane, q = idx >> 3, i = idx & 7;
  double s = 0.0;
  if (idx < NR * KB) {
    double rk[KB];
    load_vec8(rhs_sel<NR>(r, q) + blk * KB, rk);
    if (!TRANS) {
      double li[KB];
      load_row(Linv, i, li);
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j <= i) s = fma(li[j], rk[j], s);
    } else {
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j >= i) s = fma(Linv[bidx(j, i)], rk[j], s);
    }
  }
  __syncwarp();
  if (idx < NR * KB) rhs_sel<NR>(r, q)[blk * KB + i] = s;
  __syncwarp();
#endif
}

// row rr of the block at `blk` times y_K   (forward substitution update of one row, right-hand side q)
template <int NR>
IO_FORCEINLINE double fwd_row_dot(const double* blk, double* const (&r)[NR], int q, int K, int rr) {
  double row[KB], yk[KB];
  load_row(blk, rr, row);
  load_vec8(rhs_sel<NR>(r, q) + K * KB, yk);
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int k = 0; k < KB; k += 2) {
    s0 = fma(row[k], yk[k], s0);
    s1 = fma(row[k + 1], yk[k + 1], s1);
  }
  return s0 + s1;
}
// r_Kc[2c..2c+1] -= (B^T z_I)[2c..2c+1] with B = block (I,Kc) at `blk`   (backward substitution update)
template <int NR>
IO_FORCEINLINE void bwd_chunk_update(const double* blk, double* const (&r)[NR], int q, int I, int Kc, int c) {
  double zi[KB];
  load_vec8(rhs_sel<NR>(r, q) + I * KB, zi);
  double sx = 0.0, sy = 0.0, tx = 0.0, ty = 0.0;
#pragma unroll
  for (int i = 0; i < KB; i += 2) {
    const D2 t2 = reinterpret_cast<const D2*>(blk + i * KB)[c ^ bswz(i)];
    const D2 t3 = reinterpret_cast<const D2*>(blk + (i + 1) * KB)[c ^ bswz(i + 1)];
    sx = fma(t2.x, zi[i], sx);
    sy = fma(t2.y, zi[i], sy);
    tx = fma(t3.x, zi[i + 1], tx);
    ty = fma(t3.y, zi[i + 1], ty);
  }
  D2* dst = reinterpret_cast<D2*>(rhs_sel<NR>(r, q) + Kc * KB) + c;
  D2 t2 = *dst;
  t2.x -= sx + tx;
  t2.y -= sy + ty;
  *dst = t2;
}

// After the factorisation: for every coupling block inside a chain, Lb = L_(K+1,K), form
//   W = Linv_(K+1) Lb  (kept in the W buffer, slot look_fwd)   and   V = Lb Linv_K  (stored in place of Lb),
// the blocks of the fused look-ahead pivots:  y_(K+1) = Linv_(K+1) r_(K+1) - W y_K   (forward)  and
// z_K = Linv_K^T r_K - V^T z_(K+1)  (backward).  Lb itself is not needed by the substitutions any more.
IO_DEV void fuse_coupling_blocks(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double* L, double* W) {
  // one thread per (coupling block, row i): row i of W and row i of V; V is written back after a barrier because
  // it overwrites the block every row of the same task group reads
  const int ntask = P.n_wblk * KB;
#ifdef INTOPT_HOST_EMU
  (void)cx;
  for (int w = 0; w < P.n_wblk; ++w) {  // single-threaded build: whole block at a time, same arithmetic
    const int K = P.g_wblk_col[w];
    double* Lb = L + M.boff[K + 1] + (K - M.fb[K + 1]) * 64;
    const double* Linv1 = L + M.boff[K + 1] + (K + 1 - M.fb[K + 1]) * 64;
    const double* Linv0 = L + M.boff[K] + (K - M.fb[K]) * 64;
    double vall[KB][KB];
    for (int i = 0; i < KB; ++i)
      for (int j = 0; j < KB; ++j) {
        double wv = 0.0, vv = 0.0;
        for (int k = 0; k < KB; ++k) {
          if (k <= i) wv = fma(Linv1[bidx(i, k)], Lb[bidx(k, j)], wv);
          if (j <= k) vv = fma(Lb[bidx(i, k)], Linv0[bidx(k, j)], vv);
        }
        W[w * 64 + bidx(i, j)] = wv;
        vall[i][j] = vv;
      }
    for (int i = 0; i < KB; ++i)
      for (int j = 0; j < KB; ++j) Lb[bidx(i, j)] = vall[i][j];
  }
  if (ntask >= 0) return;
#endif
  for (int base = 0; base < ntask; base += cx.nthreads) {
    const int task = base + cx.tid;
    double vrow[KB];
    double* Lb = nullptr;
    int i = 0;
    if (task < ntask) {
      const int w = task >> 3;
      i = task & 7;
      const int K = P.g_wblk_col[w];
      Lb = L + M.boff[K + 1] + (K - M.fb[K + 1]) * 64;
      const double* Linv1 = L + M.boff[K + 1] + (K + 1 - M.fb[K + 1]) * 64;
      const double* Linv0 = L + M.boff[K] + (K - M.fb[K]) * 64;
      double li[KB], bi[KB], wrow[KB];
      load_row(Linv1, i, li);
      load_row(Lb, i, bi);
#pragma unroll
      for (int j = 0; j < KB; ++j) {
        wrow[j] = 0.0;
        vrow[j] = 0.0;
      }
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        double bk[KB], ik[KB];
        load_row(Lb, k, bk);
        load_row(Linv0, k, ik);
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          if (k <= i) wrow[j] = fma(li[k], bk[j], wrow[j]);    // W = Linv_(K+1) Lb   (Linv lower triangular)
          if (j <= k) vrow[j] = fma(bi[k], ik[j], vrow[j]);    // V = Lb Linv_K
        }
      }
      store_row(W + w * 64, i, wrow);
    }
    CTA_SYNC();
    if (task < ntask) store_row(Lb, i, vrow);
    CTA_SYNC();
  }
}

// fused forward look-ahead: r_(K+1) <- Linv_(K+1) r_(K+1) - W y_K, all NR right-hand sides, one full warp
template <int NR>
IO_FORCEINLINE void pivot_fused_fwd(const CtaCtx& cx, const double* Linv, const double* Wb, double* const (&r)[NR], int K) {
#ifdef INTOPT_HOST_EMU
  (void)cx;
  double tmp[NR * KB];
  for (int idx = 0; idx < NR * KB; ++idx) {
    const int q = idx >> 3, i = idx & 7;
    const double* rn = rhs_sel<NR>(r, q) + (K + 1) * KB;
    const double* yk = rhs_sel<NR>(r, q) + K * KB;
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < KB; ++j) {
      if (j <= i) s1 = fma(Linv[bidx(i, j)], rn[j], s1);
      s2 = fma(Wb[bidx(i, j)], yk[j], s2);
    }
    tmp[idx] = s1 - s2;
  }
  for (int idx = 0; idx < NR * KB; ++idx) rhs_sel<NR>(r, idx >> 3)[(K + 1) * KB + (idx & 7)] = tmp[idx];
#else
  const int idx = cx.lane, q = idx >> 3, i = idx & 7;
  double s1 = 0.0, s2 = 0.0;
  if (idx < NR * KB) {
    double rn[KB], yk[KB], li[KB], wi[KB];
    load_vec8(rhs_sel<NR>(r, q) + (K + 1) * KB, rn);
    load_vec8(rhs_sel<NR>(r, q) + K * KB, yk);
    load_row(Linv, i, li);
    load_row(Wb, i, wi);
    double s1b = 0.0, s2b = 0.0;
#pragma unroll
    for (int j = 0; j < KB; j += 2) {
      if (j <= i) s1 = fma(li[j], rn[j], s1);
      if (j + 1 <= i) s1b = fma(li[j + 1], rn[j + 1], s1b);
      s2 = fma(wi[j], yk[j], s2);
      s2b = fma(wi[j + 1], yk[j + 1], s2b);
    }
    s1 += s1b;
    s2 += s2b;
  }
  __syncwarp();
  if (idx < NR * KB) rhs_sel<NR>(r, q)[(K + 1) * KB + i] = s1 - s2;
  __syncwarp();
#endif
}
// fused backward look-ahead: r_(I-1) <- Linv_(I-1)^T r_(I-1) - V^T z_I
template <int NR>
IO_FORCEINLINE void pivot_fused_bwd(const CtaCtx& cx, const double* Linv, const double* Vb, double* const (&r)[NR], int I) {
#ifdef INTOPT_HOST_EMU
  (void)cx;
  double tmp[NR * KB];
  for (int idx = 0; idx < NR * KB; ++idx) {
    const int q = idx >> 3, i = idx & 7;
    const double* rp = rhs_sel<NR>(r, q) + (I - 1) * KB;
    const double* zi = rhs_sel<NR>(r, q) + I * KB;
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < KB; ++j) {
      if (j >= i) s1 = fma(Linv[bidx(j, i)], rp[j], s1);
      s2 = fma(Vb[bidx(j, i)], zi[j], s2);
    }
    tmp[idx] = s1 - s2;
  }
  for (int idx = 0; idx < NR * KB; ++idx) rhs_sel<NR>(r, idx >> 3)[(I - 1) * KB + (idx & 7)] = tmp[idx];
#else
  const int idx = cx.lane, q = idx >> 3, i = idx & 7;
  double s1 = 0.0, s2 = 0.0;
  if (idx < NR * KB) {
    double rp[KB], zi[KB];
    load_vec8(rhs_sel<NR>(r, q) + (I - 1) * KB, rp);
    load_vec8(rhs_sel<NR>(r, q) + I * KB, zi);
    double s1b = 0.0, s2b = 0.0;
#pragma unroll
    for (int j = 0; j < KB; j += 2) {
      if (j >= i) s1 = fma(Linv[bidx(j, i)], rp[j], s1);
      if (j + 1 >= i) s1b = fma(Linv[bidx(j + 1, i)], rp[j + 1], s1b);
      s2 = fma(Vb[bidx(j, i)], zi[j], s2);
      s2b = fma(Vb[bidx(j + 1, i)], zi[j + 1], s2b);
    }
    s1 += s1b;
    s2 += s2b;
  }
  __syncwarp();
  if (idx < NR * KB) rhs_sel<NR>(r, q)[(I - 1) * KB + i] = s1 - s2;
  __syncwarp();
#endif
}

// One step record of the substitution schedule in registers (see intopt_structure.hpp: ST_REC).
struct StepRec {
  int ne, flags, npriv, nsh;
  int toff[KSEG];
  int k[KSEG], b[KSEG], c[KSEG], look[KSEG], poff[KSEG];  // entry fields 0..4
};
IO_FORCEINLINE StepRec load_step(const int* tab, int st) {
  const U4* rec = reinterpret_cast<const U4*>(tab + st * KREC);
  StepRec R;
  const U4 h0 = rec[0], h1 = rec[1], h2 = rec[2];
  R.ne = (int)h0.x;
  R.flags = (int)h0.y;
  R.npriv = (int)h0.z;
  R.nsh = (int)h0.w;
  R.toff[0] = (int)h1.x;
  R.toff[1] = (int)h1.y;
  R.toff[2] = (int)h1.z;
  R.toff[3] = (int)h1.w;
  R.poff[0] = (int)h2.x;
  R.poff[1] = (int)h2.y;
  R.poff[2] = (int)h2.z;
  R.poff[3] = (int)h2.w;
#pragma unroll
  for (int s = 0; s < KSEG; ++s) {
    const U4 e0 = rec[3 + s];
    R.k[s] = (int)e0.x;
    R.b[s] = (int)e0.y;
    R.c[s] = (int)e0.z;
    R.look[s] = (int)e0.w;
  }
  return R;
}

template <int NR>
IO_DEV void block_solve(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* L, const double* W,
                        double* const (&r)[NR], int dbg = 0) {
  const int nsteps = P.nsteps;
  // ================= forward: L y = r =================================================================
  for (int st = 0; st < nsteps; ++st) {
    const StepRec R = load_step(M.st_fwd, st);
    if (R.flags & 1) {  // chain heads: their pivots have no look-ahead producer
      for (int s = cx.warp; s < R.ne; s += cx.nwarps) {
        const int K = SEG_PICK(R.k, s);
        pivot_inplace<NR, false>(cx, L + M.boff[K] + (K - M.fb[K]) * 64, r, K);
      }
      CTA_SYNC();
    }
    if (!(dbg & 1))
      for (int s = cx.warp; s < R.ne; s += cx.nwarps) {
        const int K = SEG_PICK(R.k, s), look = SEG_PICK(R.look, s);
        const double* Linv = L + SEG_PICK(R.poff, s);
        if (look >= 0)
          pivot_fused_fwd<NR>(cx, Linv, W + look * 64, r, K);
        else if (look == -1)
          pivot_inplace<NR, false>(cx, Linv, r, K + 1);
      }
    {
      // every thread takes row updates; the warps that just produced a pivot come last in the task order
      const int npw = (R.ne < cx.nwarps ? R.ne : cx.nwarps) * cx.warpsize;
      const int wid = cx.tid >= npw ? cx.tid - npw : cx.tid + cx.nthreads - npw;
      const int nsh = R.nsh * KB * NR;
      const int ntask = (dbg & 2) ? 0 : R.npriv * NR + nsh;
      for (int task = wid; task < ntask; task += cx.nthreads) {
        if (task >= nsh) {  // a private row of one column
          const int pt = task - nsh;
          const int s = (pt >= R.toff[1] * NR) + (pt >= R.toff[2] * NR) + (pt >= R.toff[3] * NR);
          const int loc = pt - SEG_PICK(R.toff, s) * NR;
          const int nrow = SEG_PICK(R.c, s) * KB;
          const int q = (NR > 1 && loc >= nrow) ? 1 : 0, rem = loc - q * nrow;
          const int ai = SEG_PICK(R.b, s) + (rem >> 3), rr = rem & 7;
          const int I = M.act_blk[ai];
          rhs_sel<NR>(r, q)[I * KB + rr] -= fwd_row_dot<NR>(L + M.act_off[ai], r, q, SEG_PICK(R.k, s), rr);
        } else {  // a tail row reaching into the segments: every running column, in fixed order (longest tasks first)
          const int nrow = R.nsh * KB;
          const int q = (NR > 1 && task >= nrow) ? 1 : 0, rem = task - q * nrow;
          const int I = M.seg_bord[rem >> 3], rr = rem & 7, fI = M.fb[I];
          const double* rowblk = L + M.boff[I] - fI * 64;
          double sacc = 0.0;
#pragma unroll
          for (int s = 0; s < KSEG; ++s)
            if (R.k[s] >= 0 && fI <= R.k[s]) sacc += fwd_row_dot<NR>(rowblk + R.k[s] * 64, r, q, R.k[s], rr);
          rhs_sel<NR>(r, q)[I * KB + rr] -= sacc;
        }
      }
    }
    CTA_SYNC();
  }
  // ================= backward: L^T z = y ==============================================================
  for (int st = nsteps - 1; st >= 0; --st) {
    const StepRec R = load_step(M.st_bwd, st);
    // pivots nobody produced by look-ahead: the very last block, and every segment's last block when the tail is done
    if (R.flags & 1) {
      if (cx.warp == 0) {
        const int I = P.mb - 1;
        pivot_inplace<NR, true>(cx, L + M.boff[I] + (I - M.fb[I]) * 64, r, I);
      }
      CTA_SYNC();
    }
    if (R.flags & 2) {
      for (int s = cx.warp; s < P.nseg; s += cx.nwarps) {
        const int E = P.seg_start[s] + P.seg_len[s] - 1;
        pivot_inplace<NR, true>(cx, L + M.boff[E] + (E - M.fb[E]) * 64, r, E);
      }
      CTA_SYNC();
    }
    if (!(dbg & 1))
      for (int s = cx.warp; s < R.ne; s += cx.nwarps) {
        const int I = SEG_PICK(R.k, s), look = SEG_PICK(R.look, s);
        const double* Linv = L + SEG_PICK(R.poff, s);
        if (look >= 0)
          pivot_fused_bwd<NR>(cx, Linv, L + look, r, I);
        else if (look == -1)
          pivot_inplace<NR, true>(cx, Linv, r, I - 1);
      }
    {
      const int npw = (R.ne < cx.nwarps ? R.ne : cx.nwarps) * cx.warpsize;
      const int wid = cx.tid >= npw ? cx.tid - npw : cx.tid + cx.nthreads - npw;
      const int ntask = (dbg & 2) ? 0 : R.npriv * NR;
      for (int task = wid; task < ntask; task += cx.nthreads) {
        const int s = (task >= R.toff[1] * NR) + (task >= R.toff[2] * NR) + (task >= R.toff[3] * NR);
        const int loc = task - SEG_PICK(R.toff, s) * NR;
        const int I = SEG_PICK(R.k, s), f = SEG_PICK(R.b, s);
        const int nkk = (I - f - (SEG_PICK(R.look, s) >= 0 ? 1 : 0)) * 4;
        const int q = (NR > 1 && loc >= nkk) ? 1 : 0, rem = loc - q * nkk;
        const int Kc = f + (rem >> 2);
        bwd_chunk_update<NR>(L + SEG_PICK(R.c, s) + (Kc - f) * 64, r, q, I, Kc, rem & 3);
      }
    }
    CTA_SYNC();
  }
}

// sum of a run of at most 16 consecutive entries (the row kernel's pieces, RS_MAXLEN): fully predicated, no loop
IO_FORCEINLINE double run_total16(const double* gp, int len) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    if (j < len) s0 += gp[j];
    if (j + 1 < len) s1 += gp[j + 1];
    if (j + 2 < len) s2 += gp[j + 2];
    if (j + 3 < len) s3 += gp[j + 3];
  }
  return (s0 + s1) + (s2 + s3);
}

IO_DEV void spmv_rows(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* src, double* out) {
  if (P.rs_hdr) {
    // unit-weight A in run-length form: every lane sums one run of consecutive src entries, rows are whole
    // lane ranges of a slot and are combined by a segmented shuffle scan (fixed order => deterministic)
#ifdef INTOPT_HOST_EMU
    for (int slot = 0; slot < P.rs_slots; ++slot)
      for (int l = 0; l < 32; ++l) {
        const uint32_t h = P.rs_hdr[slot * 32 + l];
        const int row = P.rs_row[slot * 32 + l];
        if (row == 0xffff) continue;
        const double s = run_total16(src + (h & 0xffffu), (int)(h >> 16));
        const bool first = l == 0 || P.rs_row[slot * 32 + l - 1] != row;
        out[row] = first ? s : out[row] + s;
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
      double s = run_total16(src + (h & 0xffffu), (int)(h >> 16));
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, s, off);
        const int r = __shfl_up_sync(0xffffffffu, row, off);
        if (cx.lane >= off && r == row) s += v;
      }
      const int rnext = __shfl_down_sync(0xffffffffu, row, 1);
      if (row != 0xffff && (cx.lane == 31 || rnext != row)) out[row] = s;
    }
#endif
    return;
  }
  const int W = cx.warpsize;
  WARP_ROWS(i, P.m) {
    const int lo = M.csr_ptr[i], hi = M.csr_ptr[i + 1];
    double s0 = 0.0, s1 = 0.0;
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
      } else {
        s0 = fma(P.csr_val[e], src[c0], s0);
        if (e1 < hi) s1 = fma(P.csr_val[e1], src[c1], s1);
        if (e2 < hi) s0 = fma(P.csr_val[e2], src[c2], s0);
        if (e3 < hi) s1 = fma(P.csr_val[e3], src[c3], s1);
      }
    }
    const double s = warp_sum(s0 + s1);
    if (cx.lane == 0) out[i] = s;
  }
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
  constexpr int NC = 2;  // index chunks in flight
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
      const double* __restrict__ vv = P.csc_val + (size_t)(g0 + t0) * 32 + l;
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
  M.act_off = P.g_act_off;
  M.csr_ptr = P.g_csr_ptr;
  M.csc_gptr = P.g_csc_gptr;
  M.asm_gptr = P.g_asm_gptr;
  M.asm_rptr = P.g_asm_rptr;
  M.act_ns = P.g_act_ns;
  M.seg_bord = P.g_seg_bord;
  M.st_fwd = P.g_st_fwd;
  M.st_bwd = P.g_st_bwd;
  return M;
}

// Copies the schedule arrays into shared memory (`dst` holds P.meta_ints ints) and returns pointers to the copies.
IO_DEV Meta stage_meta(const CtaCtx& cx, const ProblemDev& P, int* dst) {
  constexpr int NA = 13;
  // (the 16-byte aligned tables come first: st_fwd / st_bwd are read with 128-bit loads)
  // arrays of code paths this problem never takes stay in global memory (count 0): the row pointers when the
  // run-length row kernel applies, the per-term / run-length group pointers of the assembly form not in use
  const int cnt[NA] = {KREC * P.nsteps, KREC * P.nsteps, P.mb, P.mb + 1, P.mb + 1, P.n_act, P.rs_hdr ? 0 : P.m + 1,
                       ((P.n + 31) >> 5) + 1, P.asm_run ? 0 : (P.n_asm >> 5) + 1, P.asm_run ? (P.n_asm >> 5) + 1 : 0,
                       P.mb, P.nbord, P.n_act};
  const int* src[NA] = {P.g_st_fwd, P.g_st_bwd, P.g_fb, P.g_boff, P.g_act_ptr, P.g_act_blk, P.g_csr_ptr, P.g_csc_gptr,
                        P.g_asm_gptr, P.g_asm_rptr, P.g_act_ns, P.g_seg_bord, P.g_act_off};
  const int* ptr[NA];
  int* d = dst;
  for (int a = 0; a < NA; ++a) {
    CTA_FOR(i, cnt[a]) d[i] = src[a][i];
    ptr[a] = cnt[a] ? d : src[a];
    d += cnt[a];
  }
  Meta M;
  M.st_fwd = ptr[0];
  M.st_bwd = ptr[1];
  M.fb = ptr[2];
  M.boff = ptr[3];
  M.act_ptr = ptr[4];
  M.act_blk = ptr[5];
  M.csr_ptr = ptr[6];
  M.csc_gptr = ptr[7];
  M.asm_gptr = ptr[8];
  M.asm_rptr = ptr[9];
  M.act_ns = ptr[10];
  M.seg_bord = ptr[11];
  M.act_off = ptr[12];
  CTA_SYNC();
  return M;
}

// workspace n-vector slots (global, per CTA)
enum { W_X = 0, W_T, W_D, W_R2, W_P, W_W, W_DX, W_DXDT, W_NVEC };
// shared-memory m-vector slots
enum { S_B = 0, S_Y, S_R1, S_Q, S_V, S_NVEC };
constexpr int S_SCRATCH = KSEG * (KB + KB * KB) + 8 * 32 + 4;  // diag scratch (one per concurrent segment) + red + flag + mbarrier

struct SmemLayout {
  double *L, *G, *bs, *y, *r1, *q, *v, *ybuf, *dscr, *red;  // G: gather region of 2*gather_len(n) doubles
  double* W;  // fused look-ahead blocks of the forward substitution: n_wblk x 64 doubles (upper half of G if they fit)
  D2* scan;  // scratch of the double-double prefix scan: scan_len(n) D2
  int* flag;
  unsigned long long* mbar;  // mbarrier of the cold-start factor prefetch (bulk async copy)
  int* meta;  // P.meta_ints ints when the schedule is staged in shared memory
};

IO_HD int gather_len(int n) { return (gpad(n) + 3) & ~1; }
IO_HD int scan_len(int n) { return ((n + 7) >> 3) + 40; }  // D2 entries of prefix-scan scratch
IO_HD bool scan_in_factor(int bsize, int n) { return bsize >= 2 * scan_len(n); }
// the W blocks live in the upper half of the gather region (only the assembly's prefix sums use it) when they fit
IO_HD bool w_in_gather(int n_wblk, int n) { return n_wblk * 64 <= gather_len(n); }

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
  s.W = w_in_gather(P.n_wblk, P.n) ? s.G + gather_len(P.n) : ws_tail;
  s.bs = p + S_B * P.mv;
  s.y = p + S_Y * P.mv;
  s.r1 = p + S_R1 * P.mv;
  s.q = p + S_Q * P.mv;
  s.v = p + S_V * P.mv;
  p += S_NVEC * P.mv;
  s.dscr = p;
  s.ybuf = p;  // (unused: the substitutions work in place)
  s.red = s.dscr + KSEG * (KB + KB * KB);
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

// Begin copying the precomputed factor `src` (P.bsize doubles followed by the P.n_wblk*64 doubles of the W blocks,
// 16-byte aligned) into L and W.  Call after a CTA barrier that ends every earlier use of L and W.
IO_FORCEINLINE void bulk_copy_g2s(double* dst, const double* src, uint32_t bytes, unsigned long long* mbar) {
#ifndef INTOPT_HOST_EMU
  const uint32_t bar = smem_addr_u32(mbar);
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t len = bytes - off < 32768u ? bytes - off : 32768u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr_u32(dst) + off),
                 "l"((const char*)src + off), "r"(len), "r"(bar)
                 : "memory");
  }
#else
  (void)dst; (void)src; (void)bytes; (void)mbar;
#endif
}
IO_DEV void cold_copy_begin(const CtaCtx& cx, const ProblemDev& P, const double* src, double* L, double* W,
                            unsigned long long* mbar, bool l_in_smem, bool w_in_smem) {
#ifdef INTOPT_HOST_EMU
  (void)cx; (void)P; (void)src; (void)L; (void)W; (void)mbar; (void)l_in_smem; (void)w_in_smem;
#else
  if ((l_in_smem || w_in_smem) && cx.tid == 0) {
    const uint32_t lb = l_in_smem ? (uint32_t)P.bsize * 8u : 0u, wb = w_in_smem ? (uint32_t)P.n_wblk * 512u : 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses vs the async writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(mbar)), "r"(lb + wb) : "memory");
    if (lb) bulk_copy_g2s(L, src, lb, mbar);
    if (wb) bulk_copy_g2s(W, src + P.bsize, wb, mbar);
  }
#endif
}
// Complete the copy: on return every thread may read L and W.
IO_DEV void cold_copy_finish(const CtaCtx& cx, const ProblemDev& P, const double* src, double* L, double* W,
                             unsigned long long* mbar, bool l_in_smem, bool w_in_smem, unsigned parity) {
#ifdef INTOPT_HOST_EMU
  (void)mbar; (void)parity;
  l_in_smem = w_in_smem = false;
#endif
  const D2* __restrict__ s2 = reinterpret_cast<const D2*>(src);
  if (!l_in_smem) {
    D2* d2 = reinterpret_cast<D2*>(L);
    CTA_FOR(e, P.bsize >> 1) d2[e] = s2[e];  // bsize is a multiple of 64
  }
  if (!w_in_smem) {
    D2* d2 = reinterpret_cast<D2*>(W);
    CTA_FOR(e, P.n_wblk * 32) d2[e] = s2[(P.bsize >> 1) + e];
  }
#ifndef INTOPT_HOST_EMU
  if (l_in_smem || w_in_smem) mbar_wait(mbar, parity);
#endif
  CTA_SYNC();
}

struct FwdParams {
  int B;
  const double* c;
  long long c_stride;
  const double* b_batch;  // original row order, B x m, or null
  long long b_stride;
  const double* b_problem;  // internal row order or null
  const double* init;       // B x (2n+m+2) or null
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
  unsigned long long* phase_cycles;  // null unless profiling
  int debug;                         // debug switches for bottleneck isolation (env INTOPT_DEBUG), 0 in production
};

IO_DEV void forward_instance(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const FwdParams& a, int inst, double* wsi,
                             const SmemLayout& S, bool cold = false, unsigned cold_parity = 0) {
  const int m = P.m, n = P.n, mv = P.mv;
  double *L = S.L, *G = S.G, *bs = S.bs, *y = S.y, *r1 = S.r1, *q = S.q, *v = S.v;
  double* __restrict__ x = wsi + (long long)W_X * n;
  double* __restrict__ t = wsi + (long long)W_T * n;
  double* __restrict__ D = wsi + (long long)W_D * n;
  double* __restrict__ r2 = wsi + (long long)W_R2 * n;
  double* __restrict__ p = wsi + (long long)W_P * n;
  double* __restrict__ w = wsi + (long long)W_W * n;
  double* __restrict__ dx = wsi + (long long)W_DX * n;
  double* __restrict__ dxdt = wsi + (long long)W_DXDT * n;
  const double* __restrict__ c = a.c + (long long)inst * a.c_stride;

  PhaseTimer pt{a.phase_cycles, 0};
  PHASE_START(pt);
  const bool l_in_smem = a.factor_in_smem != 0, w_in_smem = a.gather_in_smem != 0 && w_in_gather(P.n_wblk, P.n);
  if (cold) cold_copy_begin(cx, P, a.cold_factor, L, S.W, S.mbar, l_in_smem, w_in_smem);
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
    CTA_FOR(k, n) {
      x[k] = s0[k];
      gstore(G, k, s0[k]);
      t[k] = s0[n + m + k];
    }
    CTA_FOR(i, m) y[i] = s0[n + P.perm[i]];
    tau = s0[2 * n + m];
    kappa = s0[2 * n + m + 1];
  } else {
    CTA_FOR(k, n) {
      x[k] = 1.0;
      gstore(G, k, 1.0);  // G mirrors x at every loop top: the residual A x gathers from shared memory
      t[k] = 1.0;
    }
  }
  CTA_SYNC();

  // ---- cold-start residual norms for the indicators (:884,907-910) -----------------------------
  double rp0, rd0, rg0;
  int nonfinite = 0;
  {
    double acc[3] = {0.0, 0.0, 0.0};
    CTA_FOR(k, n) {
      const double ck = c[k];
      const double e = ck - 1.0;
      acc[0] = fma(e, e, acc[0]);
      acc[1] += ck;
      if (!(fabs(ck) < 1e300)) acc[2] += 1.0;
    }
    double accp = 0.0;  // ||b - A 1||
    WARP_ROWS(i, m) {
      double s = 0.0;
      if (P.a_unit) {
        LANE_FOR(e, M.csr_ptr[i], M.csr_ptr[i + 1]) s += 1.0;
      } else {
        LANE_FOR(e, M.csr_ptr[i], M.csr_ptr[i + 1]) s += P.csr_val[e];
      }
      s = warp_sum(s);
      if (cx.lane == 0) {
        const double e = bs[i] - s;
        accp = fma(e, e, accp);
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

  if (cold) cold_copy_finish(cx, P, a.cold_factor, L, S.W, S.mbar, l_in_smem, w_in_smem, cold_parity);
  PHASE_MARK(pt, 0);
  while (status != ST_NONFINITE) {
    // ---- residuals at the current iterate (:765-770; the same r_p, r_d as _indicators :887-894) ----
    double acc[3] = {0.0, 0.0, 0.0};
    CTA_BATCH_FOR(k0, n, NUC) {
      double xv[NUC], tv[NUC], cv[NUC];
      BATCH_EACH(u, k, k0, n, NUC) {
        xv[u] = x[k];
        tv[u] = t[k];
        cv[u] = c[k];
      }
      BATCH_EACH(u, k, k0, n, NUC) {
        const double xk = xv[u], tk = tv[u], ck = cv[u];
        const double aty = col_dot(P, M, k, y);
        const double r = -(aty + tk - ck * tau);  // :766
        const double dk = xk / tk;                // :772
        r2[k] = r;
        D[k] = dk;
        acc[0] = fma(xk, tk, acc[0]);
        acc[1] = fma(ck, xk, acc[1]);
        acc[2] = fma(r, r, acc[2]);
      }
    }
    double accm[2] = {0.0, 0.0};
    PHASE_MARK(pt, 16);
    spmv_rows(cx, P, M, G, r1);  // A x (G mirrors x)
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
    PHASE_MARK(pt, 1);

    // ---- M = A D A^T + delta I, factor (:772-777) ------------------------------------------------
    const bool reuse_cold = cold && iters == 1;  // x = t = 1: L already holds chol(A A^T + delta I)
    if (!reuse_cold) assemble_normal_matrix(cx, P, M, D, G, S.scan, a.damping, L, a.debug);
    PHASE_MARK(pt, 2);
    // right-hand side of the (p,q) solve (:800): G = D*c
    CTA_BATCH_FOR(k0, n, NU) {
      double dv[NU], cv[NU];
      BATCH_EACH(u, k, k0, n, NU) {
        dv[u] = D[k];
        cv[u] = c[k];
      }
      BATCH_EACH(u, k, k0, n, NU) gstore(G, k, dv[u] * cv[u]);
    }
    if (reuse_cold) {
      solved = true;
      CTA_SYNC();
    } else {
      solved = block_cholesky(cx, P, M, L, S.W, S.dscr, S.flag, &pt);  // (its barriers order the G writes)
    }
    PHASE_MARK(pt, 3);
    if (!solved) {  // deltas = 0 (:855-857): the iterate is unchanged, the tests run at loop top
      CTA_FOR(k, n) gstore(G, k, x[k]);  // G must mirror x again
      CTA_SYNC();
      continue;
    }
    spmv_rows(cx, P, M, G, q);
    CTA_SYNC();
    PHASE_MARK(pt, 18);
    // predictor: gamma = 0, eta = 1
    CTA_BATCH_FOR(k0, n, NU) {
      double xv[NU], tv[NU], rv[NU], dv[NU];
      BATCH_EACH(u, k, k0, n, NU) {
        xv[u] = x[k];
        tv[u] = t[k];
        rv[u] = r2[k];
        dv[u] = D[k];
      }
      BATCH_EACH(u, k, k0, n, NU) {
        const double rxt = -(xv[u] * tv[u]);             // :788 with gamma = 0
        const double wk = rv[u] - (1.0 / xv[u]) * rxt;   // :801 with eta(0) = 1
        w[k] = wk;
        gstore(G, k, dv[u] * wk);
      }
    }
    CTA_FOR(i, m) q[i] = bs[i] + q[i];  // r = r2 + A(Dinv r1), :666 with (r1,r2) = (c,b)
    CTA_SYNC();
    PHASE_MARK(pt, 19);
    spmv_rows(cx, P, M, G, v);
    CTA_SYNC();
    PHASE_MARK(pt, 20);
    CTA_FOR(i, m) v[i] = r1[i] + v[i];
    CTA_SYNC();
    PHASE_MARK(pt, 4);
    {
      double* const rhs2[2] = {q, v};
      block_solve<2>(cx, P, M, L, S.W, rhs2);
    }
    PHASE_MARK(pt, 5);
    double cp_, bq_, cu_, bv_;
    {
      double acc2[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      CTA_BATCH_FOR(k0, n, NUC) {
        double dv[NUC], cv[NUC], wv[NUC];
        BATCH_EACH(u, k, k0, n, NUC) {
          dv[u] = D[k];
          cv[u] = c[k];
          wv[u] = w[k];
        }
        BATCH_EACH(u, k, k0, n, NUC) {
          double atq, atv;
          col_dot2(P, M, k, q, v, atq, atv);
          const double dk = dv[u], ck = cv[u];
          const double pk = dk * (atq - ck);     // :678 for (c,b)
          const double uk = dk * (atv - wv[u]);  // :678 for the predictor
          p[k] = pk;
          dx[k] = uk;
          acc2[0] = fma(ck, pk, acc2[0]);
          acc2[2] = fma(ck, uk, acc2[2]);
          if (!(fabs(pk) < 1e300)) acc2[4] += 1.0;
        }
      }
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
        CTA_FOR(k, n) gstore(G, k, x[k]);
        CTA_SYNC();
        continue;
      }
    }

    PHASE_MARK(pt, 6);
    // ---- predictor (pass 0) and corrector (pass 1) (:783-847) ---------------------------------------
    double gamma = 0.0, dtau = 0.0, dkappa = 0.0, minr = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      const double eta = 1.0 - gamma;
      const double gm = gamma * mu;
      double rtk = -(tau * kappa - gm);     // :789
      if (pass == 1) rtk -= dtau * dkappa;  // :795
      if (pass == 1) {
        CTA_BATCH_FOR(k0, n, NU) {
          double xv[NU], tv[NU], rv[NU], dv[NU], ev[NU];
          BATCH_EACH(u, k, k0, n, NU) {
            xv[u] = x[k];
            tv[u] = t[k];
            rv[u] = r2[k];
            dv[u] = D[k];
            ev[u] = dxdt[k];
          }
          BATCH_EACH(u, k, k0, n, NU) {
            const double rxt = -(xv[u] * tv[u] - gm) - ev[u];      // :788,793
            const double wk = eta * rv[u] - (1.0 / xv[u]) * rxt;  // :801
            w[k] = wk;
            gstore(G, k, dv[u] * wk);
          }
        }
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
          block_solve<1>(cx, P, M, L, S.W, rhs1);
        }
        PHASE_MARK(pt, 8);
        double acc3[2] = {0.0, 0.0};
        CTA_BATCH_FOR(k0, n, NUC) {
          double dv[NUC], cv[NUC], wv[NUC];
          BATCH_EACH(u, k, k0, n, NUC) {
            dv[u] = D[k];
            cv[u] = c[k];
            wv[u] = w[k];
          }
          BATCH_EACH(u, k, k0, n, NUC) {
            const double atv = col_dot(P, M, k, v);
            const double uk = dv[u] * (atv - wv[u]);  // :678
            dx[k] = uk;
            acc3[0] = fma(cv[u], uk, acc3[0]);
          }
        }
        CTA_FOR(i, m) acc3[1] = fma(bs[i], v[i], acc3[1]);
        cta_sum<2>(cx, acc3);
        cu_ = acc3[0];
        bv_ = acc3[1];
      }
      dtau = ((eta * r3 + 1.0 / tau * rtk - (-cu_ + bv_)) / (1.0 / tau * kappa + (-cp_ + bq_)));  // :837-838
      dkappa = (rtk - kappa * dtau) / tau;                                                         // :842
      double mn[1] = {INFINITY};
      CTA_BATCH_FOR(k0, n, NU) {
        double xv[NU], tv[NU], uv[NU], pv[NU], ev[NU];
        BATCH_EACH(u, k, k0, n, NU) {
          xv[u] = x[k];
          tv[u] = t[k];
          uv[u] = dx[k];
          pv[u] = p[k];
          ev[u] = pass == 1 ? dxdt[k] : 0.0;
        }
        BATCH_EACH(u, k, k0, n, NU) {
          const double xk = xv[u], tk = tv[u];
          double rxt = -(xk * tk - gm);
          if (pass == 1) rxt -= ev[u];
          const double dxk = uv[u] + pv[u] * dtau;           // :839
          const double dtk = (1.0 / xk) * (rxt - tk * dxk);  // :841
          dx[k] = dxk;
          if (pass == 0) dxdt[k] = dxk * dtk;
          if (dxk < 0.0) mn[0] = fmin(mn[0], xk / -dxk);  // :871-875
          if (dtk < 0.0) mn[0] = fmin(mn[0], tk / -dtk);
        }
      }
      cta_min<1>(cx, mn);
      minr = mn[0];
      PHASE_MARK(pt, 23);
      // _get_step (:870-878) with alpha0 = 1 (:843); an empty candidate set contributes 1
      double alpha = fmin(1.0, minr);
      if (dtau < 0.0) alpha = fmin(alpha, tau / -dtau);
      if (dkappa < 0.0) alpha = fmin(alpha, kappa / -dkappa);
      const double oma = 1.0 - alpha;
      gamma = oma * oma * fmin(0.1, oma);  // :844
      if (pass == 1) {
        // ---- step (:1402-1403).  alpha0*min(x/-dx) == min(alpha0*x/-dx) bit for bit
        double al = fmin(1.0, a.alpha0 * minr);
        if (dtau < 0.0) al = fmin(al, a.alpha0 * tau / -dtau);  // evaluation order of :874,876
        if (dkappa < 0.0) al = fmin(al, a.alpha0 * kappa / -dkappa);
        CTA_BATCH_FOR(k0, n, NU) {
          double xv[NU], tv[NU], uv[NU], ev[NU];
          BATCH_EACH(u, k, k0, n, NU) {
            xv[u] = x[k];
            tv[u] = t[k];
            uv[u] = dx[k];
            ev[u] = dxdt[k];
          }
          BATCH_EACH(u, k, k0, n, NU) {
            const double xk = xv[u], tk = tv[u];
            const double rxt = -(xk * tk - gm) - ev[u];
            const double dxk = uv[u];
            const double dtk = (1.0 / xk) * (rxt - tk * dxk);
            const double xn = xk + al * dxk;
            x[k] = xn;
            gstore(G, k, xn);
            t[k] = tk + al * dtk;
          }
        }
        CTA_FOR(i, m) y[i] = y[i] + al * (v[i] + q[i] * dtau);  // d_y = v + q d_tau (:840)
        tau = tau + al * dtau;
        kappa = kappa + al * dkappa;
        CTA_SYNC();
      }
    }
    PHASE_MARK(pt, 9);
  }

  // ---- epilogue (:1441-1447): x_hat = x / tau, save the iterate ------------------------------------
  double* st = a.state + (long long)inst * slen;
  double* xh = a.xhat + (long long)inst * n;
  CTA_BATCH_FOR(k0, n, NU) {
    double xv[NU], tv[NU];
    BATCH_EACH(u, k, k0, n, NU) {
      xv[u] = x[k];
      tv[u] = t[k];
    }
    BATCH_EACH(u, k, k0, n, NU) {
      xh[k] = xv[u] / tau;
      st[k] = xv[u];
      st[n + m + k] = tv[u];
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
  const bool ok = block_cholesky(cx, P, M, S.L, S.W, S.dscr, S.flag);
  CTA_FOR(e, P.bsize) out[e] = S.L[e];
  CTA_FOR(e, P.n_wblk * 64) out[P.bsize + e] = S.W[e];
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
  double* gc;        // B x nvar
  int* status;
  double* ws;
  long long ws_stride;
  int* counter;
  int factor_in_smem, gather_in_smem;
};

// K5: dL/dc at the saved iterate, VJP form:
//   whole: w = M^-1 (tau A(X2 g)), ug = X2 (A^T w - tau g)/mu, q = M^-1 (mu b + A(X2 c)), p = X2 (A^T q - c)/mu,
//          s = (x.g + c.ug - b.w)/(b.q - c.p + mu/tau^2), dc = ug + p s        (ip_model_whole.py:1515-1556)
//   alter: dc = -X2 (g - A^T M^-1 A (X2 g)) / mu  with x,t normalised by tau   (ip_model_alter.py:1524-1620)
// with M = A diag(x^2) A^T + damping I in both.
IO_DEV void backward_instance(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const BwdParams& a, int inst, double* wsi,
                              const SmemLayout& S) {
  const int m = P.m, n = P.n, nvar = a.nvar, mv = P.mv;
  double *L = S.L, *G = S.G, *bs = S.bs, *q = S.q, *v = S.v;
  double* x2 = wsi + (long long)W_D * n;
  const long long slen = 2LL * n + m + 2;
  const double* st = a.state + (long long)inst * slen;
  const double* __restrict__ c = a.c + (long long)inst * a.c_stride;
  const double* __restrict__ g = a.gx + (long long)inst * nvar;
  double* gc = a.gc + (long long)inst * nvar;
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
  CTA_FOR(k, n) {
    const double xk = whole ? xs[k] : xs[k] / tau;  // alter saves x/tau, t/tau (ip_model_alter.py:1459-1461)
    const double tk = whole ? ts[k] : ts[k] / tau;
    const double xx = xk * xk;
    x2[k] = xx;
    xt[0] = fma(xk, tk, xt[0]);
    if (k < nvar) xt[1] = fma(xk, g[k], xt[1]);
  }
  cta_sum<2>(cx, xt);
  const double mu = whole ? (xt[0] + tau * kappa) / (double)(n + 1) : xt[0] / (double)n;  // :1517 | alter :1525
  const double xg = xt[1];

  CTA_SYNC();
  assemble_normal_matrix(cx, P, M, x2, G, S.scan, a.damping, L);  // :1525-1527,1535
  CTA_FOR(k, n) gstore(G, k, (k < nvar) ? x2[k] * g[k] : 0.0);
  const bool ok = block_cholesky(cx, P, M, L, S.W, S.dscr, S.flag);
  if (!ok) {
    // the reference raises LinAlgError (:1546-1547); here: zero gradient + status
    CTA_FOR(k, nvar) gc[k] = 0.0;
    if (cx.tid == 0) a.status[inst] = ST_BREAKDOWN;
    CTA_SYNC();
    return;
  }
  spmv_rows(cx, P, M, G, v);  // A (X2 g_hat)
  CTA_SYNC();
  if (!whole) {
    double* const rhs1[1] = {v};
    block_solve<1>(cx, P, M, L, S.W, rhs1);
    at_mult<1>(cx, P, M, v, nullptr, G, nullptr);
    CTA_SYNC();
    CTA_FOR(k, nvar) gc[k] = -x2[k] * (g[k] - G[k]) / mu;
    if (cx.tid == 0) a.status[inst] = ST_OK;
    CTA_SYNC();
    return;
  }
  CTA_FOR(k, n) gstore(G, k, x2[k] * c[k]);
  CTA_FOR(i, m) v[i] = tau * v[i];  // r1 g = tau A X2 g (:1529)
  CTA_SYNC();
  spmv_rows(cx, P, M, G, q);
  CTA_SYNC();
  CTA_FOR(i, m) q[i] = mu * bs[i] + q[i];  // r2 = mu b + A(X2 c) (:1530)
  CTA_SYNC();
  {
    double* const rhs2[2] = {q, v};
    block_solve<2>(cx, P, M, L, S.W, rhs2);
  }
  double* ug = wsi + (long long)W_DX * n;
  double* pp = wsi + (long long)W_P * n;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  at_mult<2>(cx, P, M, q, v, G, G + gather_len(n));
  CTA_SYNC();
  CTA_FOR(k, n) {
    const double atq = G[k], atv = G[gather_len(n) + k];
    const double gk = (k < nvar) ? g[k] : 0.0;
    const double ugk = x2[k] * (atv - tau * gk) / mu;  // :1549 applied to g
    const double pk = x2[k] * (atq - c[k]) / mu;       // :1548
    ug[k] = ugk;
    pp[k] = pk;
    acc[0] = fma(c[k], ugk, acc[0]);
    acc[1] = fma(c[k], pk, acc[1]);
  }
  CTA_FOR(i, m) {
    acc[2] = fma(bs[i], v[i], acc[2]);
    acc[3] = fma(bs[i], q[i], acc[3]);
  }
  cta_sum<4>(cx, acc);
  const double s = (xg + acc[0] - acc[2]) / (acc[3] - acc[1] + mu / (tau * tau));  // :1552-1553
  CTA_FOR(k, nvar) gc[k] = ug[k] + pp[k] * s;                                      // :1554-1556
  if (cx.tid == 0) a.status[inst] = ST_OK;
  CTA_SYNC();
}

}  // namespace intopt
