// Host-side structure analysis for the batched interior-point solver.
//
// Everything that depends only on the sparsity pattern of the standard-form A is computed once
// per problem here (the reference redoes its whole host prep per call, Interior/
// intopt_energy_mlp.py:422):
//   * a row ordering that minimises the envelope of M = A D A^T (natural / reverse Cuthill-McKee /
//     dense-rows-last, the smallest wins);
//   * the 8x8 block-skyline layout of chol(M) and the static schedule of the blocked right-looking
//     factorisation (panel rows and trailing block pairs per block column);
//   * warp-interleaved (ELL-style) gather lists that assemble M entry by entry, replacing the dense
//     `A.dot(Dinv.reshape(-1,1)*A.T)` of Interior/ip_model_whole.py:773 and :1525-1527;
//   * A by rows (CSR, one warp per row) and by columns (32-column interleaved groups).
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <vector>

namespace intopt {

constexpr int NB = 8;           // block edge of the block-skyline factor
constexpr int NBB = NB * NB;    // doubles per block
// index of element k in the solver's padded gather buffer; MUST match intopt::gpad() of intopt_solver.cuh
inline int gpad_host(int k) { return k + (k >> 5); }
constexpr int RS_MAXLEN = 15;   // longest run piece of the run-length row kernel (<= 16, odd)

struct HostStructure {
  int m = 0, n = 0, nnz = 0, mb = 0, m8 = 0;
  std::vector<int> perm;  // internal row i == original row perm[i]
  // A by rows (internal order): one warp per row
  std::vector<int> csr_ptr, csr_col;
  std::vector<double> csr_val;
  bool a_unit = true;  // every non-zero of A equals 1: value streams are skipped
  // Per-instance values (same pattern, different numbers per LP: the knapsack cost row): position in the CALLER's
  // CSR value array of every stored copy of an entry of A; -1 marks padding.  Built only when requested.
  bool inst_values = false;
  std::vector<int> csr_src;           // parallel to csr_col / csr_val
  std::vector<int> csc_src;           // parallel to csc_val
  std::vector<int> asm_sa, asm_sb;    // parallel to asm_k / asm_w: the term's weight is val[sa] * val[sb]
  // A by columns, 32 consecutive columns per group, group length padded to a multiple of 8.  Values: entry t of
  // column 32g+l sits at (csc_gptr[g] + t) * 32 + l.  Row indices: chunk-major, the eight indices t = 8c..8c+7 of
  // column 32g+l are the 16 bytes at ((csc_gptr[g]/8 + c) * 32 + l) * 8, so a warp reads one chunk of its 32
  // columns with a single coalesced 128-bit load per lane.  Padding: row = m8 (a guaranteed-zero slot of every
  // m-vector) / value 0
  std::vector<int> csc_gptr, csc_glen;
  std::vector<uint16_t> csc_row;
  std::vector<double> csc_val;
  // run-length forms for unit-weight A (consecutive indices collapse to start | len << 16):
  //  * rows: whole rows packed into 32-lane slots, one run per lane (rs_row = row id, 0xffff = padding lane);
  std::vector<uint32_t> rs_hdr;
  std::vector<uint16_t> rs_row;
  int rs_slots = 0;
  // scalar envelope (first non-zero column of each row of M in internal order)
  std::vector<int> sky_first;
  int sky_size = 0;
  // block skyline: block row I holds block columns fb[I]..I, 64 doubles each (row-major), at boff[I]
  std::vector<int> fb, boff;
  int bsize = 0;
  std::vector<int> dummy_pos;  // diagonal positions of the padding rows m..m8-1 (set to 1)
  // factorisation schedule
  std::vector<int> act_ptr, act_blk;               // block rows I > K with fb[I] <= K
  std::vector<int> pair_ptr, pair_dst, pair_a, pair_b;  // trailing updates of step K: dst -= a * b^T
  std::vector<unsigned char> next_diag_first;      // pair_ptr[K] is the update of block (K+1,K+1)
  // assembly gather lists, entries sorted by length (descending), 32 entries per group, interleaved like csc
  int n_asm = 0;  // padded to a multiple of 32
  std::vector<int> asm_gptr, asm_glen, asm_pos;  // asm_pos < 0: padding entry
  std::vector<unsigned char> asm_diag;
  std::vector<int> asm_k;  // padded index (gpad) into the gather buffer; gpad(n) = the zero slot
  std::vector<uint16_t> asm_k16;  // the same as 16-bit indices when n < 65535 (then asm_k is not uploaded)
  std::vector<double> asm_w;
  bool asm_unit = true;
  // run-length form of the unit-weight lists (consecutive gather indices collapse to start | len << 16);
  // group g holds R = asm_rptr[g+1]-asm_rptr[g] run slots per lane (R % 4 == 0), lane-major, padded with len = 0
  std::vector<int> asm_rptr;
  std::vector<uint32_t> asm_run;
  long long asm_runs = 0;
  long long asm_terms = 0, asm_padded = 0, chol_macs = 0, solve_macs = 0;
  int ordering = 0;
};

namespace detail {

// adjacency of the graph of A A^T (rows are vertices), as sorted neighbour lists without self loops
inline std::vector<std::vector<int>> row_graph(int m, int n, const int* rp, const int* ci) {
  std::vector<std::vector<int>> colrows(n);
  for (int i = 0; i < m; ++i)
    for (int e = rp[i]; e < rp[i + 1]; ++e) colrows[ci[e]].push_back(i);
  std::vector<std::vector<int>> adj(m);
  std::vector<int> mark(m, -1);
  for (int i = 0; i < m; ++i) {
    mark[i] = i;
    for (int e = rp[i]; e < rp[i + 1]; ++e)
      for (int r : colrows[ci[e]])
        if (mark[r] != i) {
          mark[r] = i;
          adj[i].push_back(r);
        }
    std::sort(adj[i].begin(), adj[i].end());
  }
  return adj;
}

// doubles of block-skyline storage for a given ordering (the quantity that decides residency)
inline long long block_envelope_of(const std::vector<std::vector<int>>& adj, const std::vector<int>& perm) {
  int m = (int)perm.size();
  std::vector<int> inv(m);
  for (int i = 0; i < m; ++i) inv[perm[i]] = i;
  int mb = (m + NB - 1) / NB;
  std::vector<int> fbk(mb);
  for (int I = 0; I < mb; ++I) fbk[I] = I;
  for (int i = 0; i < m; ++i) {
    int f = i;
    for (int r : adj[perm[i]]) f = std::min(f, inv[r]);
    fbk[i / NB] = std::min(fbk[i / NB], f / NB);
  }
  long long blocks = 0;
  for (int I = 0; I < mb; ++I) blocks += I - fbk[I] + 1;
  return blocks * NBB;
}

inline std::vector<int> rcm(const std::vector<std::vector<int>>& adj) {
  int m = (int)adj.size();
  std::vector<int> order;
  order.reserve(m);
  std::vector<char> seen(m, 0);
  auto deg = [&](int v) { return (int)adj[v].size(); };
  auto bfs_levels = [&](int root, std::vector<int>& lvl_nodes) {
    std::vector<int> dist(m, -1);
    std::queue<int> q;
    q.push(root);
    dist[root] = 0;
    int far = 0;
    while (!q.empty()) {
      int v = q.front();
      q.pop();
      far = std::max(far, dist[v]);
      for (int w : adj[v])
        if (dist[w] < 0 && !seen[w]) {
          dist[w] = dist[v] + 1;
          q.push(w);
        }
    }
    lvl_nodes.clear();
    for (int v = 0; v < m; ++v)
      if (dist[v] == far) lvl_nodes.push_back(v);
    return far;
  };
  for (int start = 0; start < m; ++start) {
    if (seen[start]) continue;
    int root = start;  // pseudo-peripheral root of this component (George & Liu)
    std::vector<int> last;
    int ecc = bfs_levels(root, last);
    for (int it = 0; it < 8; ++it) {
      int best = last[0];
      for (int v : last)
        if (deg(v) < deg(best)) best = v;
      std::vector<int> last2;
      int ecc2 = bfs_levels(best, last2);
      if (ecc2 <= ecc) break;
      root = best;
      ecc = ecc2;
      last.swap(last2);
    }
    std::queue<int> q;  // Cuthill-McKee from root
    q.push(root);
    seen[root] = 1;
    while (!q.empty()) {
      int v = q.front();
      q.pop();
      order.push_back(v);
      std::vector<int> nb;
      for (int w : adj[v])
        if (!seen[w]) {
          seen[w] = 1;
          nb.push_back(w);
        }
      std::sort(nb.begin(), nb.end(), [&](int a, int b) { return deg(a) != deg(b) ? deg(a) < deg(b) : a < b; });
      for (int w : nb) q.push(w);
    }
  }
  std::reverse(order.begin(), order.end());
  return order;
}

// natural order, but rows whose degree in the row graph exceeds `thresh` go last (arrowhead border)
inline std::vector<int> dense_rows_last(const std::vector<std::vector<int>>& adj, int thresh) {
  int m = (int)adj.size();
  std::vector<int> a, b;
  for (int i = 0; i < m; ++i) ((int)adj[i].size() > thresh ? b : a).push_back(i);
  a.insert(a.end(), b.begin(), b.end());
  return a;
}

}  // namespace detail

// Returns 0 on success, negative on malformed input.  `force_order`: -1 = smallest block envelope.
constexpr long long kMaxAssemblyTerms = 1LL << 28;  // 268 M pair terms (dense 512 x 1024 has 134 M)

inline int analyze(int m, int n, const int* rp, const int* ci, const double* val, HostStructure& S,
                   int force_order = -1, bool inst_values = false) {
  if (m <= 0 || n <= 0 || !rp || !ci || !val) return -1;
  if (rp[0] != 0) return -2;
  for (int i = 0; i < m; ++i)
    if (rp[i + 1] < rp[i]) return -2;
  const int nnz = rp[m];
  for (int e = 0; e < nnz; ++e)
    if (ci[e] < 0 || ci[e] >= n) return -3;
  if (m + 2 * NB >= 65535) return -4;  // row indices travel as uint16
  {
    // The normal matrix is formed from gather lists with one term per pair of entries in a column: refuse problems
    // whose lists would not fit (large dense matrices) instead of exhausting host and device memory.
    std::vector<int> cnt(n, 0);
    for (int e = 0; e < nnz; ++e) cnt[ci[e]]++;
    long long terms = 0;
    for (int k = 0; k < n; ++k) terms += (long long)cnt[k] * (cnt[k] + 1) / 2;
    if (terms > kMaxAssemblyTerms) return -5;
  }
  S = HostStructure();
  S.inst_values = inst_values;
  S.m = m;
  S.n = n;
  S.nnz = nnz;
  S.mb = (m + NB - 1) / NB;
  S.m8 = S.mb * NB;
  const int mb = S.mb;

  // ---- row ordering ------------------------------------------------------------------------------
  auto adj = detail::row_graph(m, n, rp, ci);
  std::vector<std::vector<int>> cand(3);
  cand[0].resize(m);
  std::iota(cand[0].begin(), cand[0].end(), 0);
  cand[1] = detail::rcm(adj);
  cand[2] = detail::dense_rows_last(adj, std::max(8, m / 2));
  int best = 0;
  long long best_env = -1;
  for (int k = 0; k < 3; ++k) {
    if (force_order >= 0 && k != force_order) continue;
    long long env = detail::block_envelope_of(adj, cand[k]);
    if (best_env < 0 || env < best_env) {
      best_env = env;
      best = k;
    }
  }
  S.ordering = best;
  S.perm = cand[best];
  std::vector<int> inv(m);
  for (int i = 0; i < m; ++i) inv[S.perm[i]] = i;

  // ---- A by rows (internal order, columns ascending) ------------------------------------------------
  S.csr_ptr.assign(m + 1, 0);
  S.csr_col.resize(nnz);
  S.csr_val.resize(nnz);
  S.csr_src.resize(nnz);
  for (int i = 0; i < m; ++i) S.csr_ptr[i + 1] = S.csr_ptr[i] + (rp[S.perm[i] + 1] - rp[S.perm[i]]);
  for (int i = 0; i < m; ++i) {
    int src = rp[S.perm[i]], dst = S.csr_ptr[i], len = rp[S.perm[i] + 1] - src;
    std::vector<int> idx(len);
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return ci[src + a] < ci[src + b]; });
    for (int t = 0; t < len; ++t) {
      S.csr_col[dst + t] = ci[src + idx[t]];
      S.csr_val[dst + t] = val[src + idx[t]];
      S.csr_src[dst + t] = src + idx[t];
    }
  }
  S.a_unit = !inst_values;  // per-instance values always take the weighted code paths
  for (double v : S.csr_val)
    if (v != 1.0) S.a_unit = false;

  // ---- A by columns: plain CSC first, then the interleaved groups ------------------------------------
  std::vector<int> cptr(n + 1, 0), crow(nnz), csrc(nnz);
  std::vector<double> cval(nnz);
  for (int e = 0; e < nnz; ++e) cptr[S.csr_col[e] + 1]++;
  for (int k = 0; k < n; ++k) cptr[k + 1] += cptr[k];
  {
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int i = 0; i < m; ++i)
      for (int e = S.csr_ptr[i]; e < S.csr_ptr[i + 1]; ++e) {
        int k = S.csr_col[e], d = fill[k]++;
        crow[d] = i;  // ascending in i
        cval[d] = S.csr_val[e];
        csrc[d] = S.csr_src[e];
      }
  }
  const int ngrp = (n + 31) / 32;
  S.csc_gptr.assign(ngrp + 1, 0);
  S.csc_glen.assign(ngrp, 0);
  for (int g = 0; g < ngrp; ++g) {
    int len = 0;
    for (int k = 32 * g; k < std::min(n, 32 * g + 32); ++k) len = std::max(len, cptr[k + 1] - cptr[k]);
    len = (len + 7) & ~7;  // whole 16-byte index chunks
    S.csc_glen[g] = len;
    S.csc_gptr[g + 1] = S.csc_gptr[g] + len;
  }
  S.csc_row.assign((size_t)S.csc_gptr[ngrp] * 32, (uint16_t)S.m8);
  S.csc_val.assign((size_t)S.csc_gptr[ngrp] * 32, 0.0);
  if (inst_values) S.csc_src.assign((size_t)S.csc_gptr[ngrp] * 32, -1);
  for (int k = 0; k < n; ++k) {
    int g = k >> 5, l = k & 31;
    for (int t = 0; t < cptr[k + 1] - cptr[k]; ++t) {
      // indices: chunk-major (8 per lane and chunk, one 16-byte load); values: term-major (8-byte loads)
      size_t dr = ((((size_t)S.csc_gptr[g] >> 3) + (size_t)(t >> 3)) * 32 + l) * 8 + (size_t)(t & 7);
      size_t dv = ((size_t)S.csc_gptr[g] + t) * 32 + l;
      S.csc_row[dr] = (uint16_t)crow[cptr[k] + t];
      S.csc_val[dv] = cval[cptr[k] + t];
      if (inst_values) S.csc_src[dv] = csrc[cptr[k] + t];
    }
  }

  // ---- run-length forms of A (unit weights only) -----------------------------------------------------
  if (S.a_unit && gpad_host(n) < 65535) {
    auto runs_of = [](const int* idx, int len) {
      std::vector<uint32_t> r;
      int t = 0;
      while (t < len) {
        int l = 1;
        while (t + l < len && idx[t + l] == idx[t] + l && l < 65535) ++l;
        r.push_back((uint32_t)idx[t] | ((uint32_t)l << 16));
        t += l;
      }
      return r;
    };
    // rows -> slots
    bool ok = true;
    std::vector<uint32_t> hdr;
    std::vector<uint16_t> rid;
    int used = 32;  // lanes used in the current slot (start a new slot first)
    for (int i = 0; i < m && ok; ++i) {
      auto r0 = runs_of(S.csr_col.data() + S.csr_ptr[i], S.csr_ptr[i + 1] - S.csr_ptr[i]);
      // long runs are cut into pieces of at most RS_MAXLEN so every lane of a slot does the same bounded amount of
      // work (the segmented scan adds the pieces up again); the odd piece stride spreads the pieces of one run
      // over different shared-memory banks
      std::vector<uint32_t> r;
      for (uint32_t h : r0) {
        int st = (int)(h & 0xffffu), len = (int)(h >> 16);
        while (len > 0) {
          // padded coordinates: a piece that crosses a multiple of 32 also sums the (always zero) pad slot there
          const int piece = std::min(len, RS_MAXLEN);
          const int p0 = gpad_host(st), p1 = gpad_host(st + piece - 1);
          r.push_back((uint32_t)p0 | ((uint32_t)(p1 - p0 + 1) << 16));
          st += piece;
          len -= piece;
        }
      }
      if (r.size() > 32) ok = false;
      if (used + (int)r.size() > 32) {
        while (used < 32 && !hdr.empty()) {
          hdr.push_back(0u);
          rid.push_back(0xffff);
          ++used;
        }
        used = 0;
      }
      for (uint32_t h : r) {
        hdr.push_back(h);
        rid.push_back((uint16_t)i);
      }
      if (r.empty()) {  // a row without entries still has to produce out[i] = 0
        hdr.push_back(0u);
        rid.push_back((uint16_t)i);
        used += 1;
      }
      used += (int)r.size();
    }
    while (hdr.size() % 32) {
      hdr.push_back(0u);
      rid.push_back(0xffff);
    }
    // rotation of every lane's visiting order: (start + rot) mod 16 == lane mod 16 (see run_total16)
    for (size_t i = 0; i < hdr.size(); ++i) {
      const int st = (int)(hdr[i] & 0xffffu), lane = (int)(i & 31);
      hdr[i] |= (uint32_t)(((lane - st) % 16 + 16) % 16) << 24;
    }
    if (ok) {
      S.rs_hdr = hdr;
      S.rs_row = rid;
      S.rs_slots = (int)hdr.size() / 32;
    }
  }

  // ---- scalar envelope, block skyline -------------------------------------------------------------------
  S.sky_first.assign(m, 0);
  S.sky_size = 0;
  for (int i = 0; i < m; ++i) {
    int f = i;
    for (int r : adj[S.perm[i]]) f = std::min(f, inv[r]);
    S.sky_first[i] = f;
    S.sky_size += i - f + 1;
  }
  S.fb.resize(mb);
  for (int I = 0; I < mb; ++I) S.fb[I] = I;
  for (int i = 0; i < m; ++i) S.fb[i / NB] = std::min(S.fb[i / NB], S.sky_first[i] / NB);
  S.boff.assign(mb + 1, 0);
  for (int I = 0; I < mb; ++I) S.boff[I + 1] = S.boff[I] + (I - S.fb[I] + 1) * NBB;
  S.bsize = S.boff[mb];
  auto pos_of = [&](int i, int j) {  // entry (i,j), j <= i, inside the block skyline
    int I = i / NB, J = j / NB;
    const int r = i % NB, k = j % NB;  // XOR-swizzled chunk layout, see bidx() in intopt_solver.cuh
    return S.boff[I] + (J - S.fb[I]) * NBB + r * NB + ((((k >> 1) ^ ((r >> 1) & 3))) << 1) + (k & 1);
  };
  for (int i = m; i < S.m8; ++i) S.dummy_pos.push_back(pos_of(i, i));

  // ---- factorisation schedule ------------------------------------------------------------------------------
  S.act_ptr.assign(mb + 1, 0);
  S.pair_ptr.assign(mb + 1, 0);
  S.next_diag_first.assign(mb, 0);
  S.chol_macs = 0;
  S.solve_macs = 0;
  for (int K = 0; K < mb; ++K) {
    std::vector<int> act;
    for (int I = K + 1; I < mb; ++I)
      if (S.fb[I] <= K) act.push_back(I);
    for (int I : act) S.act_blk.push_back(I);
    S.act_ptr[K + 1] = (int)S.act_blk.size();
    // pairs (I >= J), the update of the next diagonal block first (look-ahead)
    std::vector<std::pair<int, int>> pairs;
    if (!act.empty() && act[0] == K + 1) {
      pairs.push_back({K + 1, K + 1});
      S.next_diag_first[K] = 1;
    }
    for (size_t a = 0; a < act.size(); ++a)
      for (size_t b2 = 0; b2 <= a; ++b2) {
        if (act[a] == K + 1 && act[b2] == K + 1) continue;
        pairs.push_back({act[a], act[b2]});
      }
    for (auto& pr : pairs) {
      int I = pr.first, J = pr.second;
      S.pair_dst.push_back(S.boff[I] + (J - S.fb[I]) * NBB);
      S.pair_a.push_back(S.boff[I] + (K - S.fb[I]) * NBB);
      S.pair_b.push_back(S.boff[J] + (K - S.fb[J]) * NBB);
    }
    S.pair_ptr[K + 1] = (int)S.pair_dst.size();
    S.chol_macs += (long long)pairs.size() * NB * NB * NB + (long long)act.size() * NB * 36 + 120 + 120;
    S.solve_macs += 2 * ((long long)act.size() * NBB + 36);
  }

  // ---- assembly gather lists -------------------------------------------------------------------------------
  S.asm_terms = 0;
  for (int k = 0; k < n; ++k) {
    long long c = cptr[k + 1] - cptr[k];
    S.asm_terms += c * (c + 1) / 2;
  }
  // entry ids over the envelope positions that are structurally non-zero (plus every diagonal)
  std::vector<int> id_of(S.bsize, -1);
  struct Ent {
    int pos;
    unsigned char diag;
    std::vector<int> k, sa, sb;
    std::vector<double> w;
    std::vector<uint32_t> runs;  // start | len << 16, longest first
  };
  std::vector<Ent> ents;
  {
    std::vector<char> present(S.bsize, 0);
    for (int k = 0; k < n; ++k)
      for (int a = cptr[k]; a < cptr[k + 1]; ++a)
        for (int b2 = cptr[k]; b2 <= a; ++b2) present[pos_of(crow[a], crow[b2])] = 1;
    for (int i = 0; i < m; ++i)
      for (int j = S.sky_first[i]; j <= i; ++j) {
        const int pos = pos_of(i, j);
        if (present[pos] || i == j) {
          id_of[pos] = (int)ents.size();
          ents.push_back({pos, (unsigned char)(i == j), {}, {}, {}, {}, {}});
        }
      }
  }
  S.asm_unit = true;
  for (int k = 0; k < n; ++k)  // ascending k => every term list is ascending in k
    for (int a = cptr[k]; a < cptr[k + 1]; ++a)
      for (int b2 = cptr[k]; b2 <= a; ++b2) {
        Ent& en = ents[id_of[pos_of(crow[a], crow[b2])]];
        const double w = cval[a] * cval[b2];
        en.k.push_back(k);
        en.w.push_back(w);
        if (inst_values) {
          en.sa.push_back(csrc[a]);
          en.sb.push_back(csrc[b2]);
        }
        if (w != 1.0 || inst_values) S.asm_unit = false;
      }
  const bool rle = S.asm_unit && gpad_host(n) < 65535;
  if (rle)
    for (Ent& en : ents) {
      size_t t = 0;
      while (t < en.k.size()) {
        size_t len = 1;
        while (t + len < en.k.size() && en.k[t + len] == en.k[t] + (int)len && len < 65535) ++len;
        // prefix-sum difference PP[gpad(a + len)] - PP[gpad(a)]: header = padded start | padded distance << 16
        const int a = en.k[t], pa = gpad_host(a), pb = gpad_host(a + (int)len);
        en.runs.push_back((uint32_t)pa | ((uint32_t)(pb - pa) << 16));
        t += len;
      }
      std::stable_sort(en.runs.begin(), en.runs.end(), [](uint32_t a, uint32_t b) { return (a >> 16) > (b >> 16); });
      S.asm_runs += (long long)en.runs.size();
    }
  // order: lanes of a warp should do the same amount of work in the same shape.  Run-length form: sort by the
  // sequence of run lengths (lexicographic, descending) so that run slot u has similar length on every lane;
  // plain form: sort by term count.
  std::vector<int> order(ents.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (rle) {
      const auto &ra = ents[a].runs, &rb = ents[b].runs;
      // a run costs two gathers whatever its length (prefix sums), so the work of a lane is its run COUNT: groups
      // of equal count first (no padding slots inside a group), then the old tie-break by run lengths
      if (((ra.size() + 3) >> 2) != ((rb.size() + 3) >> 2)) return ra.size() > rb.size();
      const size_t c = std::min(ra.size(), rb.size());
      for (size_t u = 0; u < c; ++u)
        if ((ra[u] >> 16) != (rb[u] >> 16)) return (ra[u] >> 16) > (rb[u] >> 16);
      return ra.size() > rb.size();
    }
    return ents[a].k.size() > ents[b].k.size();
  });
  const int ne = (int)ents.size();
  const int nag = (ne + 31) / 32;
  S.n_asm = nag * 32;
  S.asm_pos.assign(S.n_asm, -1);
  S.asm_diag.assign(S.n_asm, 0);
  S.asm_gptr.assign(nag + 1, 0);
  S.asm_rptr.assign(nag + 1, 0);
  for (int g = 0; g < nag; ++g) {
    size_t len = 0, R = 0;
    for (int e = 32 * g; e < std::min(ne, 32 * g + 32); ++e) {
      len = std::max(len, ents[order[e]].k.size());
      R = std::max(R, ents[order[e]].runs.size());
    }
    S.asm_gptr[g + 1] = S.asm_gptr[g] + (rle ? 0 : (int)len);  // the per-term lists are dropped when RLE applies
    S.asm_rptr[g + 1] = S.asm_rptr[g] + (int)((R + 3) & ~(size_t)3);  // multiple of 4: 16-byte header loads
  }
  S.asm_padded = rle ? (long long)S.asm_rptr[nag] * 32 : (long long)S.asm_gptr[nag] * 32;
  S.asm_k.assign((size_t)S.asm_gptr[nag] * 32, gpad_host(n));  // gpad(n) = the zero slot of the gather buffer
  S.asm_w.assign((size_t)S.asm_gptr[nag] * 32, 0.0);
  if (inst_values) {
    S.asm_sa.assign((size_t)S.asm_gptr[nag] * 32, -1);
    S.asm_sb.assign((size_t)S.asm_gptr[nag] * 32, -1);
  }
  S.asm_run.assign((size_t)S.asm_rptr[nag] * 32, 0u);
  for (int e = 0; e < ne; ++e) {
    const Ent& en = ents[order[e]];
    S.asm_pos[e] = en.pos;
    S.asm_diag[e] = en.diag;
    const int g = e / 32, l = e % 32;
    if (rle) {
      // lane-major inside the group: lane l owns R consecutive slots (R = slots per lane of this group)
      const size_t R = (size_t)(S.asm_rptr[g + 1] - S.asm_rptr[g]);
      for (size_t r = 0; r < en.runs.size(); ++r) S.asm_run[(size_t)S.asm_rptr[g] * 32 + (size_t)l * R + r] = en.runs[r];
    } else {
      for (size_t t = 0; t < en.k.size(); ++t) {
        const size_t d = ((size_t)S.asm_gptr[g] + t) * 32 + l;
        S.asm_k[d] = gpad_host(en.k[t]);
        S.asm_w[d] = en.w[t];
        if (inst_values) {
          S.asm_sa[d] = en.sa[t];
          S.asm_sb[d] = en.sb[t];
        }
      }
    }
  }
  if (!rle && gpad_host(n) < 65535) {
    S.asm_k16.resize(S.asm_k.size());
    for (size_t i = 0; i < S.asm_k.size(); ++i) S.asm_k16[i] = (uint16_t)S.asm_k[i];
  }
  // the row lists gather from the padded buffer as well (every structure above was built from the plain indices)
  for (int& cidx : S.csr_col) cidx = gpad_host(cidx);
  return 0;
}

}  // namespace intopt
