// Host-side structure analysis for the batched interior-point solver.
//
// Everything that depends only on the sparsity pattern of the standard-form A is computed once
// per problem here (the reference redoes its whole host prep per call, Interior/
// intopt_energy_mlp.py:422): row ordering, the envelope (skyline) layout of chol(A D A^T), and the
// gather lists that assemble A.diag(d).A^T entry by entry (replaces the dense
// `A.dot(Dinv.reshape(-1,1)*A.T)` of Interior/ip_model_whole.py:773 and :1525-1527).
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <vector>

namespace intopt {

struct HostStructure {
  int m = 0, n = 0, nnz = 0;
  std::vector<int> perm;  // internal row i == original row perm[i]
  std::vector<int> csr_ptr, csr_col;
  std::vector<double> csr_val;  // rows in internal order
  std::vector<int> csc_ptr, csc_row;
  std::vector<double> csc_val;  // row indices internal
  std::vector<int> sky_first, sky_off;  // row i stores columns first[i]..i at off[i]..off[i+1]-1
  int sky_size = 0;
  // assembly: entry e (e < n_asm) lives at skyline position asm_pos[e] and equals
  //   sum_{t in [asm_ptr[e], asm_ptr[e+1])} asm_w[t] * d[asm_k[t]]      (+ damping on the diagonal)
  std::vector<int> asm_ptr, asm_pos, asm_k;
  std::vector<unsigned char> asm_diag;
  std::vector<double> asm_w;
  bool asm_unit = true;  // every asm_w == 1 (0/1 matrices): the kernel skips the weight stream
  long long asm_terms = 0, chol_macs = 0;
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

inline long long envelope_of(const std::vector<std::vector<int>>& adj, const std::vector<int>& perm) {
  int m = (int)perm.size();
  std::vector<int> inv(m);
  for (int i = 0; i < m; ++i) inv[perm[i]] = i;
  long long env = 0;
  for (int i = 0; i < m; ++i) {
    int f = i;
    for (int r : adj[perm[i]]) f = std::min(f, inv[r]);
    env += i - f + 1;
  }
  return env;
}

inline std::vector<int> rcm(const std::vector<std::vector<int>>& adj) {
  int m = (int)adj.size();
  std::vector<int> order;
  order.reserve(m);
  std::vector<char> seen(m, 0);
  auto deg = [&](int v) { return (int)adj[v].size(); };
  auto bfs_levels = [&](int root, std::vector<int>& lvl_nodes) {
    // returns eccentricity; lvl_nodes = last level
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
    // pseudo-peripheral root of this component (George & Liu)
    int root = start;
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
    // Cuthill-McKee from root
    std::queue<int> q;
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

// Returns 0 on success, negative on malformed input.  `force_order`: -1 choose the smallest envelope.
inline int analyze(int m, int n, const int* rp, const int* ci, const double* val, HostStructure& S,
                   int force_order = -1) {
  if (m <= 0 || n <= 0 || !rp || !ci || !val) return -1;
  if (rp[0] != 0) return -2;
  for (int i = 0; i < m; ++i)
    if (rp[i + 1] < rp[i]) return -2;
  const int nnz = rp[m];
  for (int e = 0; e < nnz; ++e)
    if (ci[e] < 0 || ci[e] >= n) return -3;
  S.m = m;
  S.n = n;
  S.nnz = nnz;

  // ---- choose the row ordering with the smallest envelope --------------------------------
  auto adj = detail::row_graph(m, n, rp, ci);
  std::vector<std::vector<int>> cand(3);
  cand[0].resize(m);
  std::iota(cand[0].begin(), cand[0].end(), 0);
  cand[1] = detail::rcm(adj);
  {
    // border threshold: rows coupled to more than half of the rows
    cand[2] = detail::dense_rows_last(adj, std::max(8, m / 2));
  }
  int best = 0;
  long long best_env = -1;
  for (int k = 0; k < 3; ++k) {
    if (force_order >= 0 && k != force_order) continue;
    long long env = detail::envelope_of(adj, cand[k]);
    if (best_env < 0 || env < best_env) {
      best_env = env;
      best = k;
    }
  }
  S.ordering = best;
  S.perm = cand[best];
  std::vector<int> inv(m);
  for (int i = 0; i < m; ++i) inv[S.perm[i]] = i;

  // ---- permuted CSR and CSC ---------------------------------------------------------------
  S.csr_ptr.assign(m + 1, 0);
  S.csr_col.resize(nnz);
  S.csr_val.resize(nnz);
  for (int i = 0; i < m; ++i) S.csr_ptr[i + 1] = S.csr_ptr[i] + (rp[S.perm[i] + 1] - rp[S.perm[i]]);
  for (int i = 0; i < m; ++i) {
    int src = rp[S.perm[i]], dst = S.csr_ptr[i], len = rp[S.perm[i] + 1] - src;
    std::vector<int> idx(len);
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return ci[src + a] < ci[src + b]; });
    for (int t = 0; t < len; ++t) {
      S.csr_col[dst + t] = ci[src + idx[t]];
      S.csr_val[dst + t] = val[src + idx[t]];
    }
  }
  S.csc_ptr.assign(n + 1, 0);
  for (int e = 0; e < nnz; ++e) S.csc_ptr[S.csr_col[e] + 1]++;
  for (int k = 0; k < n; ++k) S.csc_ptr[k + 1] += S.csc_ptr[k];
  S.csc_row.resize(nnz);
  S.csc_val.resize(nnz);
  {
    std::vector<int> fill(S.csc_ptr.begin(), S.csc_ptr.end() - 1);
    for (int i = 0; i < m; ++i)
      for (int e = S.csr_ptr[i]; e < S.csr_ptr[i + 1]; ++e) {
        int k = S.csr_col[e], d = fill[k]++;
        S.csc_row[d] = i;  // ascending in i because rows are visited in order
        S.csc_val[d] = S.csr_val[e];
      }
  }

  // ---- envelope layout -----------------------------------------------------------------------
  S.sky_first.assign(m, 0);
  for (int i = 0; i < m; ++i) {
    int f = i;
    for (int r : adj[S.perm[i]]) f = std::min(f, inv[r]);
    S.sky_first[i] = f;
  }
  S.sky_off.assign(m + 1, 0);
  for (int i = 0; i < m; ++i) S.sky_off[i + 1] = S.sky_off[i] + (i - S.sky_first[i] + 1);
  S.sky_size = S.sky_off[m];

  // ---- assembly gather lists: one list per structurally non-zero lower entry (i >= j) -------
  // count terms per (i,j): walk the columns; rows inside a column are ascending
  S.asm_terms = 0;
  for (int k = 0; k < n; ++k) {
    long long c = S.csc_ptr[k + 1] - S.csc_ptr[k];
    S.asm_terms += c * (c + 1) / 2;
  }
  // position -> entry id map over the skyline (dense in the envelope)
  std::vector<int> cnt(S.sky_size, 0);
  for (int k = 0; k < n; ++k)
    for (int a = S.csc_ptr[k]; a < S.csc_ptr[k + 1]; ++a)
      for (int b2 = S.csc_ptr[k]; b2 <= a; ++b2) {
        int i = S.csc_row[a], j = S.csc_row[b2];
        cnt[S.sky_off[i] + (j - S.sky_first[i])]++;
      }
  std::vector<int> entry_of(S.sky_size, -1);
  S.asm_pos.clear();
  S.asm_ptr.assign(1, 0);
  S.asm_diag.clear();
  for (int i = 0; i < m; ++i)
    for (int j = S.sky_first[i]; j <= i; ++j) {
      int pos = S.sky_off[i] + (j - S.sky_first[i]);
      if (cnt[pos] > 0 || i == j) {  // the diagonal is always assembled (it carries the damping)
        entry_of[pos] = (int)S.asm_pos.size();
        S.asm_pos.push_back(pos);
        S.asm_diag.push_back(i == j ? 1 : 0);
        S.asm_ptr.push_back(S.asm_ptr.back() + cnt[pos]);
      }
    }
  S.asm_k.resize((size_t)S.asm_terms);
  S.asm_w.resize((size_t)S.asm_terms);
  S.asm_unit = true;
  {
    std::vector<int> fill(S.asm_ptr.begin(), S.asm_ptr.end() - 1);
    for (int k = 0; k < n; ++k)
      for (int a = S.csc_ptr[k]; a < S.csc_ptr[k + 1]; ++a)
        for (int b2 = S.csc_ptr[k]; b2 <= a; ++b2) {
          int i = S.csc_row[a], j = S.csc_row[b2];
          int e = entry_of[S.sky_off[i] + (j - S.sky_first[i])];
          int d = fill[e]++;
          double w = S.csc_val[a] * S.csc_val[b2];
          S.asm_k[d] = k;
          S.asm_w[d] = w;
          if (w != 1.0) S.asm_unit = false;
        }
  }

  // ---- MAC count of the envelope Cholesky -------------------------------------------------------
  S.chol_macs = 0;
  {
    // for column j, rows i > j with first[i] <= j contribute (j - max(first[i], first[j])) MACs
    for (int i = 0; i < m; ++i)
      for (int j = S.sky_first[i]; j <= i; ++j) S.chol_macs += j - std::max(S.sky_first[i], S.sky_first[j]);
  }
  return 0;
}

}  // namespace intopt
