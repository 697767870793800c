// TEST-ONLY: compiles the solver's device code (intopt_solver.cuh) as plain single-threaded C++
// (-DINTOPT_HOST_EMU) so the arithmetic and indexing can be unit-tested on a CPU-only machine.
// Never loaded by the neuripsintopt_b200 package; the product path is the CUDA build only.
#define INTOPT_HOST_EMU 1
#include <cstring>
#include <vector>

#include "../../neuripsintopt_b200/csrc/intopt_bind.hpp"

extern "C" int emu_solve(int m, int n, const int* rp, const int* ci, const double* val, const double* b,
                         int force_order, int B, const double* c, long long c_stride, const double* b_batch,
                         double alpha0, double tol, double thr, double damping_fwd, int max_iter,
                         const double* init, double* xhat, double* state, int* status, int* iters,
                         int bwd_mode, double damping_bwd, int nvar, const double* gx, double* gc,
                         int* bstatus, int* envelope_out, const double* a_vals, int plain, double beta) {
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rp, ci, val, S, force_order, a_vals != nullptr);
  if (rc) return rc;
  if (envelope_out) *envelope_out = S.bsize;
  std::vector<double> bperm;
  if (b) {
    bperm.resize(m);
    for (int i = 0; i < m; ++i) bperm[i] = b[S.perm[i]];
  }
  intopt::ProblemDev P{};
  intopt::bind_problem(S, P, [](const auto& v) { return v.data(); });
  std::vector<double> smem((size_t)S.bsize + 2 * intopt::gather_len(n) + intopt::smem_fixed_doubles(S) +
                           (intopt::meta_ints(S) + 1) / 2);
  std::vector<double> ws((size_t)intopt::W_NVEC * n);
  const intopt::SmemLayout L = intopt::carve<true>(P, smem.data(), nullptr, 1, 1);
  CtaCtx cx;
  cx.red = L.red;
  intopt::Meta M = intopt::stage_meta(cx, P, L.meta);
  intopt::FwdParams a{};
  a.B = B; a.c = c; a.c_stride = c_stride; a.b_batch = b_batch; a.b_stride = m;
  a.b_problem = b ? bperm.data() : nullptr; a.init = init; a.xhat = xhat; a.state = state;
  std::vector<double> vws((size_t)rp[m] + S.csc_val.size() + S.asm_w.size() + 2);
  a.a_vals = a_vals; a.a_stride = rp[m]; a.vals_ws = vws.data(); a.vals_ws_stride = 0;
  if (a_vals) intopt::instance_value_streams(P, M, vws.data());
  a.status = status; a.iters = iters; a.ws = ws.data(); a.ws_stride = 0;
  a.factor_in_smem = 1; a.gather_in_smem = 1;
  a.alpha0 = alpha0; a.tol = tol; a.thr = thr; a.damping = damping_fwd; a.max_iter = max_iter;
  a.plain = plain; a.beta = beta;
  // cold-start factor (as the CUDA host code does): built once with the instance code path, copied in at iteration 1
  std::vector<double> cold((size_t)S.bsize);
  int cold_ok = 0;
  if (!init && !a_vals) {
    intopt::cold_factor_build(cx, P, M, damping_fwd, ws.data(), L, cold.data(), &cold_ok);
    a.cold_factor = cold.data();
    a.cold_ok = &cold_ok;
  }
  const bool use_cold = !init && !a_vals && cold_ok;
  for (int i = 0; i < B; ++i) intopt::forward_instance(cx, P, M, a, i, ws.data(), L, use_cold, 0u);
  if (gx && gc) {
    intopt::BwdParams g{};
    g.B = B; g.mode = bwd_mode; g.nvar = nvar; g.damping = damping_bwd; g.state = state; g.c = c;
    g.c_stride = c_stride; g.b_batch = b_batch; g.b_stride = m; g.b_problem = b ? bperm.data() : nullptr;
    g.gx = gx; g.gc = gc; g.status = bstatus; g.ws = ws.data(); g.ws_stride = 0;
    g.a_vals = a_vals; g.a_stride = rp[m]; g.vals_ws = vws.data(); g.vals_ws_stride = 0;
    g.factor_in_smem = 1; g.gather_in_smem = 1;
    for (int i = 0; i < B; ++i) intopt::backward_instance(cx, P, M, g, i, ws.data(), L);
  }
  return 0;
}
