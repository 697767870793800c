// Level-scheduled factorisation and substitutions (included inside namespace intopt by intopt_solver.cuh).
//
// The blocked Cholesky and the two substitution sweeps are dependent chains: one 8x8 diagonal block after the other,
// with a CTA barrier per step.  Block columns whose diagonal blocks do not depend on each other (the three machine bands
// of the Energy LP; generally: independent subtrees of the elimination tree) are eliminated TOGETHER here: a group of up
// to LV_MAX columns per step, the diagonal blocks of a group factored side by side in the four lane octets of one warp.
// The chain has n_groups steps instead of mb (Energy: 10 instead of 22).  Every block row / block column that a step
// updates is a "target" with an explicit, ordered list of source blocks (built on the host, intopt_structure.hpp): a
// target that several columns of the group feed (the arrowhead border) is updated by ONE work item that walks its
// sources in column order, so every sum keeps a fixed order and results stay bitwise reproducible — and the kernels
// chase no indices on their critical path.
//
// Same arithmetic as block_cholesky / block_solve (ip_model_whole.py:742-745 cho_factor / cho_solve): right-looking
// blocked Cholesky with inverted diagonal blocks, blocked forward / backward substitution.

IO_FORCEINLINE int lv_ncols(const Meta& M, int g) { return M.grp_ptr[g + 1] - M.grp_ptr[g]; }
// column slot c of group g: block column index and offset of its diagonal block
IO_FORCEINLINE int lv_col(const Meta& M, int g, int c) { return M.lv[LV_COL_REC][(4 * g + c) * 2]; }
IO_FORCEINLINE int lv_diag_off(const Meta& M, int g, int c) { return M.lv[LV_COL_REC][(4 * g + c) * 2 + 1]; }
// target record:  index | nsrc << 11 | first source << 14 ;  source:  block offset | column slot << 16
IO_FORCEINLINE int lv_t_index(int rec) { return rec & 0x7ff; }
IO_FORCEINLINE int lv_t_nsrc(int rec) { return (rec >> 11) & 7; }
IO_FORCEINLINE int lv_t_first(int rec) { return rec >> 14; }
IO_FORCEINLINE int lv_s_off(int src) { return src & 0xffff; }
IO_FORCEINLINE int lv_s_col(int src) { return src >> 16; }
IO_FORCEINLINE const double* lv_block(const Meta& M, const double* L, int I, int K) {
  return L + M.boff[I] + (K - M.fb[I]) * 64;
}
IO_FORCEINLINE double* lv_block(const Meta& M, double* L, int I, int K) { return L + M.boff[I] + (K - M.fb[I]) * 64; }
// block (I, K) lies inside the skyline below the diagonal
IO_FORCEINLINE bool lv_has(const Meta& M, int I, int K) { return K >= 0 && K < I && M.fb[I] <= K; }
// which of the (up to four) columns holds flat work item idx, given the cumulative counts tab[0..4]
IO_FORCEINLINE int lv_which(const int* tab, int idx) { return (idx >= tab[1]) + (idx >= tab[2]) + (idx >= tab[3]); }

// Factor + invert the diagonal blocks of group `g` (one lane octet per column).  Before that, each diagonal block (K',K')
// receives the updates of the previous group's panels: C -= sum_K L_(K',K) L_(K',K)^T over the sources of target K' in
// the previous group's next-pivot list (gp < 0: none).  One warp.  `scr`: KB + KB*KB doubles (only the single-threaded
// test build uses it: the CUDA build stages L_KK in the block's own storage, which ends up holding the inverse anyway).
IO_DEV void lv_diag_group(const CtaCtx& cx, const Meta& M, double* L, double* scr, int* fail, int g, int gp) {
  const int nc = lv_ncols(M, g);
  // the next-pivot targets of group gp are exactly the columns of group g that receive updates, in column order
  const int t0 = gp >= 0 ? M.lv[LV_FL_PTR][gp] : 0, nt = gp >= 0 ? M.lv[LV_FL_PTR][gp + 1] - t0 : 0;
#ifdef INTOPT_HOST_EMU
  for (int o = 0; o < nc; ++o) {
    const int Kn = lv_col(M, g, o);
    double* blk = L + lv_diag_off(M, g, o);
    for (int t = 0; t < nt; ++t) {
      const int rec = M.lv[LV_FL][t0 + t];
      if (lv_t_index(rec) != Kn) continue;
      for (int s = 0; s < lv_t_nsrc(rec); ++s) {
        const double* panel = L + lv_s_off(M.lv[LV_FSRC][lv_t_first(rec) + s]);
        for (int i = 0; i < KB; ++i)
          for (int k = 0; k <= i; ++k) {
            double acc = 0.0;
            for (int j = 0; j < KB; ++j) acc = fma(panel[bidx(i, j)], panel[bidx(k, j)], acc);
            blk[bidx(i, k)] -= acc;
          }
      }
    }
    diag_block_factor_invert(cx, blk, scr, fail, nullptr);
  }
#else
  const unsigned full = 0xffffffffu;
  const int o = cx.lane >> 3, i = cx.lane & 7;
  const bool valid = o < nc;
  const int Kn = valid ? lv_col(M, g, o) : -1;
  double* blk = L + (valid ? lv_diag_off(M, g, o) : 0);
  double a[KB], rinv[KB];
  if (valid) {
    load_row(blk, i, a);
  } else {
#pragma unroll
    for (int k = 0; k < KB; ++k) a[k] = (k == i) ? 1.0 : 0.0;
  }
  // this octet's target record (at most one per column; columns without sources have none)
  int rec = 0;
  for (int t = 0; t < nt; ++t) {
    const int r2 = M.lv[LV_FL][t0 + t];
    if (lv_t_index(r2) == Kn) rec = r2;
  }
  const int ns = lv_t_nsrc(rec), s0 = lv_t_first(rec);
  for (int s = 0; s < ns; ++s) {
    const double* panel = L + lv_s_off(M.lv[LV_FSRC][s0 + s]);
    double own[KB];
    load_row(panel, i, own);
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      double pk[KB];
      load_row(panel, k, pk);
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < KB; ++j) acc = fma(own[j], pk[j], acc);
      a[k] -= acc;
    }
  }
  __syncwarp();
  bool bad = false;
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    double piv = __shfl_sync(full, a[j], j, 8);
    if (!(piv > 0.0) || !(piv < 1e300)) {
      bad = valid;
      piv = 1.0;
    }
    const double r = rsqrt(piv);
    rinv[j] = r;
    a[j] = (i == j) ? piv * r : a[j] * r;
#pragma unroll
    for (int k = j + 1; k < KB; ++k) {
      const double lkj = __shfl_sync(full, a[j], k, 8);
      a[k] = fma(-a[j], lkj, a[k]);
    }
  }
  if (bad) *fail = 1;
  (void)scr;
  if (valid) store_row(blk, i, a);  // publish L_KK (entries above the diagonal are junk and never read)
  __syncwarp();
  double x[KB];
#pragma unroll
  for (int r = 0; r < KB; ++r) {
    double acc = 0.0;
    if (valid) {
#pragma unroll
      for (int k = 0; k < KB; ++k)
        if (k < r) acc = fma(blk[bidx(r, k)], (k >= i) ? x[k] : 0.0, acc);
    }
    x[r] = (r == i) ? rinv[r] : ((r > i) ? -rinv[r] * acc : 0.0);
  }
  __syncwarp();
  if (valid) {
#pragma unroll
    for (int r = 0; r < KB; ++r)
      if (r >= i) blk[bidx(r, i)] = x[r];
  }
  __syncwarp();
#endif
}

// one 4x4 tile of  dst(I,J) -= L_IK L_JK^T  (tile row tr, tile column tc)
IO_FORCEINLINE void lv_tile_acc(const double* A, const double* Bm, int tr, int tc, double (&acc)[4][4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    D2 av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      av[i] = reinterpret_cast<const D2*>(A + (4 * tr + i) * KB)[c ^ bswz(4 * tr + i)];
      bv[i] = reinterpret_cast<const D2*>(Bm + (4 * tc + i) * KB)[c ^ bswz(4 * tc + i)];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i].y, bv[j].y, fma(av[i].x, bv[j].x, acc[i][j]));
  }
}
IO_FORCEINLINE void lv_tile_sub(double* dst, int tr, int tc, const double (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    D2* drow = reinterpret_cast<D2*>(dst + (4 * tr + i) * KB);
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

IO_DEV bool block_cholesky_lv(const CtaCtx& cx, const ProblemDev& P, const Meta& M, double* L, double* scr, int* fail_flag,
                              PhaseTimer* pt = nullptr) {
  const int ng = P.n_groups;
  if (cx.tid == 0) *fail_flag = 0;
  CTA_SYNC();
  if (cx.warp == 0) lv_diag_group(cx, M, L, scr, fail_flag, 0, -1);
  CTA_SYNC();
  const int o_tid = cx.nwarps > 1 ? cx.tid - cx.warpsize : 0;
  const int o_n = cx.nwarps > 1 ? cx.nthreads - cx.warpsize : 1;
  for (int g = 0; g < ng; ++g) {
    const int* tab = M.grp_tab + 16 * g;
    // (a) panel rows of every column of the group: L_IK = C_IK Linv_KK^T; the panel blocks are the group's sources
    const int p0 = M.lv[LV_FSRC_PTR][g], np = M.lv[LV_FSRC_PTR][g + 1] - p0;
    CTA_FOR(idx, np * KB) {
      const int src = M.lv[LV_FSRC][p0 + (idx >> 3)];
      const double* Linv = L + lv_diag_off(M, g, lv_s_col(src));
      double* blk = L + lv_s_off(src);
      double cin[KB], out[KB];
      load_row(blk, idx & 7, cin);
#pragma unroll
      for (int cc = 0; cc < KB; ++cc) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < KB; ++k)
          if (k <= cc) s = fma(cin[k], Linv[bidx(cc, k)], s);
        out[cc] = s;
      }
      store_row(blk, idx & 7, out);
    }
    CTA_SYNC();
    if (pt) { PHASE_MARK(*pt, 11); }
    // (b) warp 0: the next group's diagonal blocks (their updates by this group's panels, factor, invert);
    //     the other warps: every other trailing update of this group
    if (cx.warp == 0 && g + 1 < ng) lv_diag_group(cx, M, L, scr, fail_flag, g + 1, g);
    if (pt) { PHASE_MARK(*pt, 13); }
    if (cx.nwarps == 1 || cx.warp > 0) {
      const int* tt = tab + 5;  // tile counts: tt[0..4] per column, tt[5] incl. the merged pairs of the shared rows
      const int r0 = M.rg_ptr[g], nc = lv_ncols(M, g);
      for (int task = o_tid; task < tt[5]; task += o_n) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        if (task < tt[4]) {
          // a pair of active rows of ONE column; pairs of two shared rows belong to the merged tasks below
          const int c = lv_which(tt, task);
          const int t = task - tt[c];
          const int pi = t >> 2, tr = (t >> 1) & 1, tc = t & 1;
          const int ia = tri_row(pi), ib = pi - ia * (ia + 1) / 2;
          const int K = lv_col(M, g, c), a0 = M.act_ptr[K];
          const int ea = M.act_blk[a0 + ia], eb = M.act_blk[a0 + ib];
          const int I = ea & 0xffff, J = eb & 0xffff;
          if ((ea >> 16) && (eb >> 16)) continue;
          if (I == J && M.col_group[I] == g + 1) continue;  // diagonal block of the next group: warp 0 has it
          lv_tile_acc(lv_block(M, (const double*)L, I, K), lv_block(M, (const double*)L, J, K), tr, tc, acc);
          lv_tile_sub(lv_block(M, L, I, J), tr, tc, acc);
        } else {
          const int t = task - tt[4];
          const int pi = t >> 2, tr = (t >> 1) & 1, tc = t & 1;
          const int ia = tri_row(pi), ib = pi - ia * (ia + 1) / 2;
          const int I = M.rg_blk[r0 + ia], J = M.rg_blk[r0 + ib];
          if (I == J && M.col_group[I] == g + 1) continue;
          for (int c = 0; c < nc; ++c) {
            const int K = lv_col(M, g, c);
            if (lv_has(M, I, K) && lv_has(M, J, K))
              lv_tile_acc(lv_block(M, (const double*)L, I, K), lv_block(M, (const double*)L, J, K), tr, tc, acc);
          }
          lv_tile_sub(lv_block(M, L, I, J), tr, tc, acc);
        }
      }
    }
    CTA_SYNC();
    if (pt) { PHASE_MARK(*pt, 14); }
  }
  return *fail_flag == 0;
}

// Pivots of a group: y[(c * NR + q) * 8 + i] = row i of Linv_KK (or Linv_KK^T) times r_q[K*8 ..], K = column slot c of the
// group.  One lane per (column, row), both right-hand sides in the lane (two independent chains).
template <int NR, bool TRANS>
IO_FORCEINLINE void lv_pivot_group(const CtaCtx& cx, const Meta& M, const double* L, double* const (&r)[NR], int g,
                                   double* yout) {
  const int nc = lv_ncols(M, g);
  LANE_FOR(idx, 0, nc * KB) {
    const int c = idx >> 3, i = idx & 7;
    const int K = lv_col(M, g, c);
    const double* Linv = L + lv_diag_off(M, g, c);
    double li[KB];
    if (!TRANS) {
      load_row(Linv, i, li);
    } else {
#pragma unroll
      for (int j = 0; j < KB; ++j) li[j] = Linv[bidx(j, i)];
    }
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      double rk[KB];
      load_vec8(rhs_sel<NR>(r, q) + K * KB, rk);
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (TRANS ? (j >= i) : (j <= i)) s = fma(li[j], rk[j], s);
      yout[(c * NR + q) * KB + i] = s;
    }
  }
}

// r_(target)[rr] -= sum over the target's sources of (row rr of the source block) . y_(source column), q-th right-hand side
template <int NR>
IO_FORCEINLINE void lv_fwd_item(const Meta& M, const double* L, double* const (&r)[NR], const double* ycur, int rec, int rr) {
  const int I = lv_t_index(rec), ns = lv_t_nsrc(rec), s0 = lv_t_first(rec);
  double sacc[NR];
#pragma unroll
  for (int q = 0; q < NR; ++q) sacc[q] = 0.0;
  for (int s = 0; s < ns; ++s) {
    const int src = M.lv[LV_FSRC][s0 + s];
    double row[KB];
    load_row(L + lv_s_off(src), rr, row);
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      double yk[KB];
      load_vec8(ycur + (lv_s_col(src) * NR + q) * KB, yk);
#pragma unroll
      for (int k = 0; k < KB; ++k) sacc[q] = fma(row[k], yk[k], sacc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < NR; ++q) rhs_sel<NR>(r, q)[I * KB + rr] -= sacc[q];
}

// r_(target)[2 ch, 2 ch + 1] -= sum over the sources of (columns 2 ch, 2 ch + 1 of the source block)^T . z_(source row)
template <int NR>
IO_FORCEINLINE void lv_bwd_item(const Meta& M, const double* L, double* const (&r)[NR], const double* zcur, int rec, int ch) {
  const int Kc = lv_t_index(rec), ns = lv_t_nsrc(rec), s0 = lv_t_first(rec);
  double sx[NR], sy[NR];
#pragma unroll
  for (int q = 0; q < NR; ++q) sx[q] = sy[q] = 0.0;
  for (int s = 0; s < ns; ++s) {
    const int src = M.lv[LV_BSRC][s0 + s];
    const double* blk = L + lv_s_off(src);
    D2 t2[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) t2[i] = reinterpret_cast<const D2*>(blk + i * KB)[ch ^ bswz(i)];
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      double zi[KB];
      load_vec8(zcur + (lv_s_col(src) * NR + q) * KB, zi);
#pragma unroll
      for (int i = 0; i < KB; ++i) {
        sx[q] = fma(t2[i].x, zi[i], sx[q]);
        sy[q] = fma(t2[i].y, zi[i], sy[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    D2* dst = reinterpret_cast<D2*>(rhs_sel<NR>(r, q) + Kc * KB) + ch;
    D2 t = *dst;
    t.x -= sx[q];
    t.y -= sy[q];
    *dst = t;
  }
}

template <int NR>
IO_DEV void block_solve_lv(const CtaCtx& cx, const ProblemDev& P, const Meta& M, const double* L, double* const (&r)[NR],
                           double* ybuf) {
  const int ng = P.n_groups;
  constexpr int YB = LV_MAX * NR * KB;
  const int o_tid = cx.nwarps > 1 ? cx.tid - cx.warpsize : 0;
  const int o_n = cx.nwarps > 1 ? cx.nthreads - cx.warpsize : 1;
  // ---- forward: L y = r --------------------------------------------------------------------------------
  if (cx.warp == 0) lv_pivot_group<NR, false>(cx, M, L, r, 0, ybuf);
  CTA_SYNC();
  for (int g = 0; g < ng; ++g) {
    const int nc = lv_ncols(M, g);
    const double* ycur = ybuf + (g & 1) * YB;
    double* ynext = ybuf + ((g + 1) & 1) * YB;
    if (cx.warp == 0) {
      LANE_FOR(idx, 0, nc * NR * KB) {  // y_K is final: store it
        const int c = idx / (NR * KB), rem = idx - c * (NR * KB);
        rhs_sel<NR>(r, rem >> 3)[lv_col(M, g, c) * KB + (rem & 7)] = ycur[idx];
      }
      if (g + 1 < ng) {
        const int t0 = M.lv[LV_FL_PTR][g], nt = M.lv[LV_FL_PTR][g + 1] - t0;
        LANE_FOR(idx, 0, nt * KB) lv_fwd_item<NR>(M, L, r, ycur, M.lv[LV_FL][t0 + (idx >> 3)], idx & 7);  // next pivots' rows
        WARP_SYNC();
        lv_pivot_group<NR, false>(cx, M, L, r, g + 1, ynext);
      }
    }
    if (cx.nwarps == 1 || cx.warp > 0) {
      const int t0 = M.lv[LV_FO_PTR][g], nt = M.lv[LV_FO_PTR][g + 1] - t0;
      for (int idx = o_tid; idx < nt * KB; idx += o_n) lv_fwd_item<NR>(M, L, r, ycur, M.lv[LV_FO][t0 + (idx >> 3)], idx & 7);
    }
    CTA_SYNC();
  }
  // ---- backward: L^T z = y ------------------------------------------------------------------------------
  if (cx.warp == 0) lv_pivot_group<NR, true>(cx, M, L, r, ng - 1, ybuf + ((ng - 1) & 1) * YB);
  CTA_SYNC();
  for (int g = ng - 1; g >= 0; --g) {
    const int nc = lv_ncols(M, g);
    const double* zcur = ybuf + (g & 1) * YB;
    double* znext = ybuf + ((g + 1) & 1) * YB;  // (buffer of parity (g-1) & 1)
    if (cx.warp == 0) {
      LANE_FOR(idx, 0, nc * NR * KB) {  // z_I is final
        const int c = idx / (NR * KB), rem = idx - c * (NR * KB);
        rhs_sel<NR>(r, rem >> 3)[lv_col(M, g, c) * KB + (rem & 7)] = zcur[idx];
      }
      if (g > 0) {
        const int t0 = M.lv[LV_BL_PTR][g], nt = M.lv[LV_BL_PTR][g + 1] - t0;
        LANE_FOR(idx, 0, nt * 4) lv_bwd_item<NR>(M, L, r, zcur, M.lv[LV_BL][t0 + (idx >> 2)], idx & 3);  // next pivots
        WARP_SYNC();
        lv_pivot_group<NR, true>(cx, M, L, r, g - 1, znext);
      }
    }
    if (cx.nwarps == 1 || cx.warp > 0) {
      const int t0 = M.lv[LV_BO_PTR][g], nt = M.lv[LV_BO_PTR][g + 1] - t0;
      for (int idx = o_tid; idx < nt * 4; idx += o_n) lv_bwd_item<NR>(M, L, r, zcur, M.lv[LV_BO][t0 + (idx >> 2)], idx & 3);
    }
    CTA_SYNC();
  }
}
