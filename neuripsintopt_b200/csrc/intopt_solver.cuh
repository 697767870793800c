// Per-instance interior-point solver: one CTA owns one LP for the whole IPM loop.
//
// The factor of M = A.diag(d).A^T + delta.I lives in shared memory in envelope (skyline) form, the
// m-vectors live in shared memory, the n-vectors in a per-instance global workspace that stays in
// L1/L2.  A (CSR + CSC) and the assembly gather lists are shared by every instance and stay in L2.
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

#ifdef INTOPT_HOST_EMU
#define IO_DEV inline
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
#define IO_DEV __device__ inline
#define IO_FORCEINLINE __device__ __forceinline__
struct CtaCtx {
  int tid, nthreads, lane, warp, nwarps;
  static constexpr int warpsize = 32;
  double* red;  // >= 4*32 doubles of shared scratch
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

// loops: every thread of the CTA strides over [0,n)
#define CTA_FOR(i, n) for (int i = cx.tid; i < (n); i += cx.nthreads)
// one warp per row, lanes stride over the row's entries
#define WARP_ROWS(i, m) for (int i = cx.warp; i < (m); i += cx.nwarps)
#define LANE_FOR(e, lo, hi) for (int e = (lo) + cx.lane; e < (hi); e += cx.warpsize)

namespace intopt {

enum { ST_OK = 0, ST_MAXITER = 1, ST_INFEAS = 2, ST_BREAKDOWN = 3, ST_NONFINITE = 4 };

struct ProblemDev {
  int m, n, nnz;
  const int* csr_ptr;
  const int* csr_col;
  const double* csr_val;
  const int* csc_ptr;
  const int* csc_row;
  const double* csc_val;
  const double* b;  // internal row order, may be null
  const int* perm;  // internal row -> original row
  const int* sky_first;
  const int* sky_off;
  int sky_size;
  int n_asm;
  const int* asm_ptr;
  const int* asm_pos;
  const unsigned char* asm_diag;
  const int* asm_k;
  const double* asm_w;
  int asm_unit;
};

// Sum N partial values over the CTA; every thread receives the totals (fixed order => deterministic).
template <int N>
IO_DEV void cta_sum(const CtaCtx& cx, double (&v)[N]) {
#ifdef INTOPT_HOST_EMU
  (void)cx;
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
// M = A diag(d) A^T + delta I, assembled entry by entry into the envelope storage L (K1).
// ---------------------------------------------------------------------------------------------
IO_DEV void assemble_normal_matrix(const CtaCtx& cx, const ProblemDev& P, const double* __restrict__ d,
                                   double delta, double* L) {
  CTA_FOR(e, P.sky_size) L[e] = 0.0;
  CTA_SYNC();
  CTA_FOR(e, P.n_asm) {
    const int lo = P.asm_ptr[e], hi = P.asm_ptr[e + 1];
    double s = 0.0;
    if (P.asm_unit) {
      for (int t = lo; t < hi; ++t) s += d[P.asm_k[t]];
    } else {
      for (int t = lo; t < hi; ++t) s = fma(P.asm_w[t], d[P.asm_k[t]], s);
    }
    if (P.asm_diag[e]) s += delta;
    L[P.asm_pos[e]] = s;
  }
  CTA_SYNC();
}

// ---------------------------------------------------------------------------------------------
// Envelope Cholesky, left-looking by columns (K2).  On exit L holds the strict lower factor in
// place, dg[j] = L_jj and dgi[j] = 1/L_jj.  Returns false (for every thread) on a non-positive or
// non-finite pivot (the reference's cho_factor failure, ip_model_whole.py:742,755-758).
// ---------------------------------------------------------------------------------------------
IO_DEV bool envelope_cholesky(const CtaCtx& cx, const ProblemDev& P, double* L, double* dg, double* dgi,
                              int* fail_flag) {
  const int m = P.m;
  const int* __restrict__ first = P.sky_first;
  const int* __restrict__ off = P.sky_off;
  if (cx.tid == 0) *fail_flag = 0;
  CTA_SYNC();
  for (int j = 0; j < m; ++j) {
    const int fj = first[j];
    const double* Lj = L + off[j] - fj;  // Lj[k] = L(j,k)
    // phase 1: s_ij = M_ij - sum_k L_ik L_jk for every row i >= j that reaches column j
    CTA_FOR(r, m - j) {
      const int i = j + r;
      const int fi = first[i];
      if (fi <= j) {
        double* Li = L + off[i] - fi;
        const int k0 = fi > fj ? fi : fj;
        double s = Li[j];
        for (int k = k0; k < j; ++k) s = fma(-Li[k], Lj[k], s);
        Li[j] = s;
      }
    }
    CTA_SYNC();
    // phase 2: pivot and scale the column
    double piv = Lj[j];
    if (!(piv > 0.0) || !(piv < 1e300)) {
      if (cx.tid == 0) *fail_flag = 1;
      piv = 1.0;
    }
    const double dj = sqrt(piv);
    const double rinv = 1.0 / dj;
    CTA_FOR(r, m - j) {
      const int i = j + r;
      if (r == 0) {
        dg[j] = dj;
        dgi[j] = rinv;
      } else if (first[i] <= j) {
        L[off[i] - first[i] + j] *= rinv;
      }
    }
    CTA_SYNC();
  }
  return *fail_flag == 0;
}

// Solve L L^T z = r in place (r -> z), column-oriented substitution over the envelope.
// r[j] is kept unscaled while a sweep runs (consumers multiply by 1/L_jj themselves), so one
// barrier per column suffices; a parallel scaling pass finishes each sweep.
IO_DEV void envelope_solve(const CtaCtx& cx, const ProblemDev& P, const double* L, const double* dg,
                           const double* dgi, double* r) {
  const int m = P.m;
  const int* __restrict__ first = P.sky_first;
  const int* __restrict__ off = P.sky_off;
  (void)dg;
  // forward: L y = r
  for (int j = 0; j < m - 1; ++j) {
    const double yj = r[j] * dgi[j];
    CTA_FOR(q, m - j - 1) {
      const int i = j + 1 + q;
      if (first[i] <= j) r[i] = fma(-L[off[i] - first[i] + j], yj, r[i]);
    }
    CTA_SYNC();
  }
  CTA_FOR(j, m) r[j] *= dgi[j];
  CTA_SYNC();
  // backward: L^T z = y ; row i of L is column i of L^T
  for (int i = m - 1; i > 0; --i) {
    const double zi = r[i] * dgi[i];
    const int fi = first[i];
    const double* Li = L + off[i] - fi;
    CTA_FOR(q, i - fi) {
      const int k = fi + q;
      r[k] = fma(-Li[k], zi, r[k]);
    }
    CTA_SYNC();
  }
  CTA_FOR(j, m) r[j] *= dgi[j];
  CTA_SYNC();
}

// out[i] = base[i]*scale + sum_e A_ie * src[col_e]   (row-wise, one warp per row)
IO_DEV void spmv_rows(const CtaCtx& cx, const ProblemDev& P, const double* __restrict__ src, double* out) {
  WARP_ROWS(i, P.m) {
    double s = 0.0;
    LANE_FOR(e, P.csr_ptr[i], P.csr_ptr[i + 1]) s = fma(P.csr_val[e], src[P.csr_col[e]], s);
    s = warp_sum(s);
    if (cx.lane == 0) out[i] = s;
  }
}

IO_FORCEINLINE double col_dot(const ProblemDev& P, int k, const double* yv) {
  double s = 0.0;
  for (int e = P.csc_ptr[k]; e < P.csc_ptr[k + 1]; ++e) s = fma(P.csc_val[e], yv[P.csc_row[e]], s);
  return s;
}

struct FwdParams {
  int B;
  const double* c;
  long long c_stride;
  const double* b_batch;  // original row order, B x m, or null
  long long b_stride;
  const double* init;     // B x (2n+m+2) or null
  double* xhat;
  double* state;
  int* status;
  int* iters;
  double* ws;  // per instance: 9 n-vectors (+ sky_size if the factor spills)
  long long ws_stride;
  int factor_in_smem;
  double alpha0, tol, thr, damping;
  int max_iter;
};

// workspace n-vector slots
enum { W_X = 0, W_T, W_D, W_R2, W_P, W_W, W_DW, W_DX, W_DXDT, W_NVEC };

// shared-memory m-vector slots (after the factor)
enum { S_B = 0, S_Y, S_R1, S_Q, S_V, S_DG, S_DGI, S_NVEC };

IO_DEV void forward_instance(const CtaCtx& cx, const ProblemDev& P, const FwdParams& a, int inst, double* smem) {
  const int m = P.m, n = P.n;
  double* wsi = a.ws + (long long)inst * a.ws_stride;
  double* L = a.factor_in_smem ? smem : wsi + (long long)W_NVEC * n;
  double* sv = a.factor_in_smem ? smem + P.sky_size : smem;
  double* bs = sv + S_B * m;
  double* y = sv + S_Y * m;
  double* r1 = sv + S_R1 * m;
  double* q = sv + S_Q * m;
  double* v = sv + S_V * m;
  double* dg = sv + S_DG * m;
  double* dgi = sv + S_DGI * m;
  int* fail_flag = (int*)(sv + S_NVEC * m);  // one double slot; cx.red follows it

  double* x = wsi + (long long)W_X * n;
  double* t = wsi + (long long)W_T * n;
  double* D = wsi + (long long)W_D * n;
  double* r2 = wsi + (long long)W_R2 * n;
  double* p = wsi + (long long)W_P * n;
  double* w = wsi + (long long)W_W * n;
  double* dw = wsi + (long long)W_DW * n;
  double* dx = wsi + (long long)W_DX * n;
  double* dxdt = wsi + (long long)W_DXDT * n;
  const double* __restrict__ c = a.c + (long long)inst * a.c_stride;

  // ---- load b, initial point (_initialization :647-661) -----------------------------------------
  CTA_FOR(i, m) {
    bs[i] = a.b_batch ? a.b_batch[(long long)inst * a.b_stride + P.perm[i]] : P.b[i];
  }
  double tau = 1.0, kappa = 1.0;
  const long long slen = 2LL * n + m + 2;
  if (a.init) {
    const double* s0 = a.init + (long long)inst * slen;
    CTA_FOR(k, n) {
      x[k] = s0[k];
      t[k] = s0[n + m + k];
    }
    CTA_FOR(i, m) y[i] = s0[n + P.perm[i]];
    tau = s0[2 * n + m];
    kappa = s0[2 * n + m + 1];
  } else {
    CTA_FOR(k, n) {
      x[k] = 1.0;
      t[k] = 1.0;
    }
    CTA_FOR(i, m) y[i] = 0.0;
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
    // ||b - A 1||
    double accp = 0.0;
    WARP_ROWS(i, m) {
      double s = 0.0;
      LANE_FOR(e, P.csr_ptr[i], P.csr_ptr[i + 1]) s += P.csr_val[e];
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
  if (nonfinite) {
    // the reference raises LinAlgError for non-finite c (:1332-1341); here: status word, x untouched
    status = ST_NONFINITE;
  }

  while (status != ST_NONFINITE) {
    // ---- residuals at the current iterate (:765-770; identical to _indicators' r_p, r_d :887-894) --
    double xt_, cx_, nd2;
    {
      double acc[3] = {0.0, 0.0, 0.0};
      CTA_FOR(k, n) {
        const double xk = x[k], tk = t[k], ck = c[k];
        const double aty = col_dot(P, k, y);
        const double r = -(aty + tk - ck * tau);  // :766
        r2[k] = r;
        D[k] = xk / tk;  // :772
        acc[0] = fma(xk, tk, acc[0]);
        acc[1] = fma(ck, xk, acc[1]);
        acc[2] = fma(r, r, acc[2]);
      }
      double accm[2] = {0.0, 0.0};
      WARP_ROWS(i, m) {
        double s = 0.0;
        LANE_FOR(e, P.csr_ptr[i], P.csr_ptr[i + 1]) s = fma(P.csr_val[e], x[P.csr_col[e]], s);
        s = warp_sum(s);
        if (cx.lane == 0) {
          const double r = -(s - bs[i] * tau);  // :765
          r1[i] = r;
          accm[0] = fma(r, r, accm[0]);
          accm[1] = fma(bs[i], y[i], accm[1]);
        }
      }
      double red[5] = {acc[0], acc[1], acc[2], accm[0], accm[1]};
      cta_sum<5>(cx, red);
      xt_ = red[0];
      cx_ = red[1];
      nd2 = red[2];
      const double np2 = red[3], by_ = red[4];
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

      // ---- M = A D A^T + delta I, factor (:772-777) ------------------------------------------------
      assemble_normal_matrix(cx, P, D, a.damping, L);
      solved = envelope_cholesky(cx, P, L, dg, dgi, fail_flag);
      if (!solved) continue;  // deltas = 0 (:855-857): the iterate is unchanged, tests run at loop top

      // ---- (p,q) = sym_solve(c, b) (:800, :663-681), once per iteration ---------------------------
      CTA_FOR(k, n) dw[k] = D[k] * c[k];
      CTA_SYNC();
      spmv_rows(cx, P, dw, q);
      CTA_SYNC();
      CTA_FOR(i, m) q[i] = bs[i] + q[i];
      CTA_SYNC();
      envelope_solve(cx, P, L, dg, dgi, q);
      double cp_, bq_;
      {
        double acc2[3] = {0.0, 0.0, 0.0};
        CTA_FOR(k, n) {
          const double pk = D[k] * (col_dot(P, k, q) - c[k]);
          p[k] = pk;
          acc2[0] = fma(c[k], pk, acc2[0]);
          if (!(fabs(pk) < 1e300)) acc2[2] += 1.0;
        }
        CTA_FOR(i, m) {
          acc2[1] = fma(bs[i], q[i], acc2[1]);
          if (!(fabs(q[i]) < 1e300)) acc2[2] += 1.0;
        }
        cta_sum<3>(cx, acc2);
        cp_ = acc2[0];
        bq_ = acc2[1];
        if (acc2[2] > 0.0) {  // NaN in p or q (:805-806) -> unsolved
          solved = false;
          continue;
        }
      }

      // ---- predictor (i=0) and corrector (i=1) passes (:783-847) -----------------------------------
      double gamma = 0.0, dtau = 0.0, dkappa = 0.0, minr = 0.0;
      for (int pass = 0; pass < 2; ++pass) {
        const double eta = 1.0 - gamma;
        const double gm = gamma * mu;
        double rtk = -(tau * kappa - gm);          // :789
        if (pass == 1) rtk -= dtau * dkappa;       // :795
        CTA_FOR(k, n) {
          double rxt = -(x[k] * t[k] - gm);        // :788
          if (pass == 1) rxt -= dxdt[k];           // :793
          const double wk = eta * r2[k] - (1.0 / x[k]) * rxt;  // :801
          w[k] = wk;
          dw[k] = D[k] * wk;
        }
        CTA_SYNC();
        spmv_rows(cx, P, dw, v);
        CTA_SYNC();
        CTA_FOR(i, m) v[i] = eta * r1[i] + v[i];   // r = r2 + A(Dinv r1), :666
        CTA_SYNC();
        envelope_solve(cx, P, L, dg, dgi, v);
        double acc3[2] = {0.0, 0.0};
        CTA_FOR(k, n) {
          const double uk = D[k] * (col_dot(P, k, v) - w[k]);  // :678
          dx[k] = uk;
          acc3[0] = fma(c[k], uk, acc3[0]);
        }
        CTA_FOR(i, m) acc3[1] = fma(bs[i], v[i], acc3[1]);
        cta_sum<2>(cx, acc3);
        const double cu_ = acc3[0], bv_ = acc3[1];
        dtau = ((eta * r3 + 1.0 / tau * rtk - (-cu_ + bv_)) / (1.0 / tau * kappa + (-cp_ + bq_)));  // :837-838
        dkappa = (rtk - kappa * dtau) / tau;  // :842
        double mn[1] = {INFINITY};
        CTA_FOR(k, n) {
          const double xk = x[k], tk = t[k];
          double rxt = -(xk * tk - gm);
          if (pass == 1) rxt -= dxdt[k];
          const double dxk = dx[k] + p[k] * dtau;           // :839
          const double dtk = (1.0 / xk) * (rxt - tk * dxk);  // :841
          dx[k] = dxk;
          if (pass == 0) dxdt[k] = dxk * dtk;
          if (dxk < 0.0) mn[0] = fmin(mn[0], xk / -dxk);     // :871-875
          if (dtk < 0.0) mn[0] = fmin(mn[0], tk / -dtk);
        }
        cta_min<1>(cx, mn);
        minr = mn[0];
        // _get_step (:870-878) with alpha0 = 1 (:843); an empty candidate set contributes 1
        double alpha = fmin(1.0, minr);
        if (dtau < 0.0) alpha = fmin(alpha, tau / -dtau);
        if (dkappa < 0.0) alpha = fmin(alpha, kappa / -dkappa);
        const double oma = 1.0 - alpha;
        gamma = oma * oma * fmin(0.1, oma);                   // :844
        if (pass == 1) {
          // ---- step (:1402-1403).  alpha0*min(x/-dx) == min(alpha0*x/-dx) bit for bit
          double al = fmin(1.0, a.alpha0 * minr);
          if (dtau < 0.0) al = fmin(al, a.alpha0 * tau / -dtau);        // evaluation order of :874,876
          if (dkappa < 0.0) al = fmin(al, a.alpha0 * kappa / -dkappa);
          CTA_FOR(k, n) {
            const double xk = x[k], tk = t[k];
            const double rxt = -(xk * tk - gm) - dxdt[k];
            const double dxk = dx[k];
            const double dtk = (1.0 / xk) * (rxt - tk * dxk);
            x[k] = xk + al * dxk;
            t[k] = tk + al * dtk;
          }
          CTA_FOR(i, m) y[i] = y[i] + al * (v[i] + q[i] * dtau);  // d_y = v + q d_tau (:840)
          tau = tau + al * dtau;
          kappa = kappa + al * dkappa;
          CTA_SYNC();
        }
      }
    }
  }

  // ---- epilogue (:1441-1447): x_hat = x / tau, save the iterate ------------------------------------
  double* st = a.state + (long long)inst * slen;
  double* xh = a.xhat + (long long)inst * n;
  CTA_FOR(k, n) {
    const double xk = x[k];
    xh[k] = xk / tau;
    st[k] = xk;
    st[n + m + k] = t[k];
  }
  CTA_FOR(i, m) st[n + P.perm[i]] = y[i];
  if (cx.tid == 0) {
    st[2 * n + m] = tau;
    st[2 * n + m + 1] = kappa;
    a.status[inst] = status;
    a.iters[inst] = iters;
  }
}

struct BwdParams {
  int B, mode, nvar;
  double damping;
  const double* state;
  const double* c;
  long long c_stride;
  const double* b_batch;
  long long b_stride;
  const double* gx;  // B x nvar
  double* gc;        // B x nvar
  int* status;
  double* ws;
  long long ws_stride;
  int factor_in_smem;
};

// K5: dL/dc at the saved iterate, VJP form:
//   whole: w = M^-1 (tau A(X2 g)), ug = X2 (A^T w - tau g)/mu, q = M^-1 (mu b + A(X2 c)), p = X2 (A^T q - c)/mu,
//          s = (x.g + c.ug - b.w)/(b.q - c.p + mu/tau^2), dc = ug + p s        (ip_model_whole.py:1515-1556)
//   alter: dc = -X2 (g - A^T M^-1 A (X2 g)) / mu  with x,t normalised by tau   (ip_model_alter.py:1524-1620)
// with M = A diag(x^2) A^T + damping I in both.
IO_DEV void backward_instance(const CtaCtx& cx, const ProblemDev& P, const BwdParams& a, int inst, double* smem) {
  const int m = P.m, n = P.n, nvar = a.nvar;
  double* wsi = a.ws + (long long)inst * a.ws_stride;
  double* L = a.factor_in_smem ? smem : wsi + (long long)W_NVEC * n;
  double* sv = a.factor_in_smem ? smem + P.sky_size : smem;
  double* bs = sv + S_B * m;
  double* q = sv + S_Q * m;
  double* v = sv + S_V * m;
  double* dg = sv + S_DG * m;
  double* dgi = sv + S_DGI * m;
  int* fail_flag = (int*)(sv + S_NVEC * m);
  double* x2 = wsi + (long long)W_D * n;
  double* tmp = wsi + (long long)W_DW * n;
  double* ug = wsi + (long long)W_DX * n;
  const long long slen = 2LL * n + m + 2;
  const double* st = a.state + (long long)inst * slen;
  const double* __restrict__ c = a.c + (long long)inst * a.c_stride;
  const double* __restrict__ g = a.gx + (long long)inst * nvar;
  double* gc = a.gc + (long long)inst * nvar;
  const double* xs = st;
  const double* ts = st + n + m;
  const double tau = st[2 * n + m], kappa = st[2 * n + m + 1];
  const bool whole = a.mode == 0;
  const double scale = whole ? 1.0 : 1.0 / tau;  // alter saves x/tau, t/tau (ip_model_alter.py:1459-1461)

  CTA_FOR(i, m) bs[i] = a.b_batch ? a.b_batch[(long long)inst * a.b_stride + P.perm[i]] : (P.b ? P.b[i] : 0.0);
  double xt[2] = {0.0, 0.0};
  CTA_FOR(k, n) {
    const double xk = whole ? xs[k] : xs[k] / tau;
    const double tk = whole ? ts[k] : ts[k] / tau;
    x2[k] = xk * xk;
    xt[0] = fma(xk, tk, xt[0]);
    if (k < nvar) xt[1] = fma(xk, g[k], xt[1]);
  }
  cta_sum<2>(cx, xt);
  (void)scale;
  const double mu = whole ? (xt[0] + tau * kappa) / (double)(n + 1) : xt[0] / (double)n;  // :1517 | alter :1525
  const double xg = xt[1];

  assemble_normal_matrix(cx, P, x2, a.damping, L);  // :1525-1527,1535
  const bool ok = envelope_cholesky(cx, P, L, dg, dgi, fail_flag);
  if (!ok) {
    // the reference raises LinAlgError (:1546-1547); here: zero gradient + status
    CTA_FOR(k, nvar) gc[k] = 0.0;
    if (cx.tid == 0) a.status[inst] = ST_BREAKDOWN;
    return;
  }
  // w = M^-1 (tau * A (X2 g_hat))
  CTA_FOR(k, n) tmp[k] = (k < nvar) ? x2[k] * g[k] : 0.0;
  CTA_SYNC();
  spmv_rows(cx, P, tmp, v);
  CTA_SYNC();
  if (whole) {
    CTA_FOR(i, m) v[i] = tau * v[i];
    CTA_SYNC();
  }
  envelope_solve(cx, P, L, dg, dgi, v);
  if (!whole) {
    CTA_FOR(k, nvar) gc[k] = -x2[k] * (g[k] - col_dot(P, k, v)) / mu;
    if (cx.tid == 0) a.status[inst] = ST_OK;
    return;
  }
  // q = M^-1 (mu b + A (X2 c))   (:1530,1542)
  CTA_FOR(k, n) tmp[k] = x2[k] * c[k];
  CTA_SYNC();
  spmv_rows(cx, P, tmp, q);
  CTA_SYNC();
  CTA_FOR(i, m) q[i] = mu * bs[i] + q[i];
  CTA_SYNC();
  envelope_solve(cx, P, L, dg, dgi, q);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  CTA_FOR(k, n) {
    const double gk = (k < nvar) ? g[k] : 0.0;
    const double ugk = x2[k] * (col_dot(P, k, v) - tau * gk) / mu;  // :1549 applied to g
    const double pk = x2[k] * (col_dot(P, k, q) - c[k]) / mu;       // :1548
    ug[k] = ugk;
    tmp[k] = pk;
    acc[0] = fma(c[k], ugk, acc[0]);
    acc[1] = fma(c[k], pk, acc[1]);
  }
  CTA_FOR(i, m) {
    acc[2] = fma(bs[i], v[i], acc[2]);
    acc[3] = fma(bs[i], q[i], acc[3]);
  }
  cta_sum<4>(cx, acc);
  const double s = (xg + acc[0] - acc[2]) / (acc[3] - acc[1] + mu / (tau * tau));  // :1552-1553
  CTA_FOR(k, nvar) gc[k] = ug[k] + tmp[k] * s;  // :1554-1556
  if (cx.tid == 0) a.status[inst] = ST_OK;
}

}  // namespace intopt
