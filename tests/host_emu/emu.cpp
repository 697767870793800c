// TEST-ONLY: compiles the solver's device code (intopt_solver.cuh) as plain single-threaded C++
// (-DINTOPT_HOST_EMU) so the arithmetic and indexing can be unit-tested on a CPU-only machine.
// Never loaded by the neuripsintopt_b200 package; the product path is the CUDA build only.
#define INTOPT_HOST_EMU 1
#include <cstring>
#include <vector>

#include "../../neuripsintopt_b200/csrc/intopt_solver.cuh"
#include "../../neuripsintopt_b200/csrc/intopt_structure.hpp"

namespace {
intopt::ProblemDev view(const intopt::HostStructure& S, const std::vector<double>& b) {
  intopt::ProblemDev D{};
  D.m = S.m; D.n = S.n; D.nnz = S.nnz;
  D.csr_ptr = S.csr_ptr.data(); D.csr_col = S.csr_col.data(); D.csr_val = S.csr_val.data();
  D.csc_ptr = S.csc_ptr.data(); D.csc_row = S.csc_row.data(); D.csc_val = S.csc_val.data();
  D.b = b.empty() ? nullptr : b.data();
  D.perm = S.perm.data();
  D.sky_first = S.sky_first.data(); D.sky_off = S.sky_off.data(); D.sky_size = S.sky_size;
  D.n_asm = (int)S.asm_pos.size();
  D.asm_ptr = S.asm_ptr.data(); D.asm_pos = S.asm_pos.data(); D.asm_diag = S.asm_diag.data();
  D.asm_k = S.asm_k.data(); D.asm_w = S.asm_w.data(); D.asm_unit = S.asm_unit ? 1 : 0;
  return D;
}
}  // namespace

extern "C" int emu_solve(int m, int n, const int* rp, const int* ci, const double* val, const double* b,
                         int force_order, int B, const double* c, long long c_stride, const double* b_batch,
                         double alpha0, double tol, double thr, double damping_fwd, int max_iter,
                         const double* init, double* xhat, double* state, int* status, int* iters,
                         int bwd_mode, double damping_bwd, int nvar, const double* gx, double* gc,
                         int* bstatus, int* envelope_out) {
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rp, ci, val, S, force_order);
  if (rc) return rc;
  if (envelope_out) *envelope_out = S.sky_size;
  std::vector<double> bperm;
  if (b) {
    bperm.resize(m);
    for (int i = 0; i < m; ++i) bperm[i] = b[S.perm[i]];
  }
  intopt::ProblemDev P = view(S, bperm);
  std::vector<double> smem((size_t)S.sky_size + intopt::S_NVEC * m + 1 + 8 * 32);
  std::vector<double> ws((size_t)intopt::W_NVEC * n * B);
  CtaCtx cx;
  cx.red = smem.data() + S.sky_size + intopt::S_NVEC * m + 1;
  intopt::FwdParams a{};
  a.B = B; a.c = c; a.c_stride = c_stride; a.b_batch = b_batch; a.b_stride = m; a.init = init; a.xhat = xhat; a.state = state;
  a.status = status; a.iters = iters; a.ws = ws.data(); a.ws_stride = (long long)intopt::W_NVEC * n;
  a.factor_in_smem = 1; a.alpha0 = alpha0; a.tol = tol; a.thr = thr; a.damping = damping_fwd; a.max_iter = max_iter;
  for (int i = 0; i < B; ++i) intopt::forward_instance(cx, P, a, i, smem.data());
  if (gx && gc) {
    intopt::BwdParams g{};
    g.B = B; g.mode = bwd_mode; g.nvar = nvar; g.damping = damping_bwd; g.state = state; g.c = c;
    g.c_stride = c_stride; g.b_batch = b_batch; g.b_stride = m; g.gx = gx; g.gc = gc; g.status = bstatus; g.ws = ws.data();
    g.ws_stride = a.ws_stride; g.factor_in_smem = 1;
    for (int i = 0; i < B; ++i) intopt::backward_instance(cx, P, g, i, smem.data());
  }
  return 0;
}
