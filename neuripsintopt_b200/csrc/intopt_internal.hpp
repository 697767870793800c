// Internals shared by the translation units of libintopt_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

#include "../../include/intopt_b200.h"
#include "intopt_solver.cuh"

struct IntoptDense;  // lock-step dense path (intopt_dense.cu), null for problems on the persistent-CTA path

// Cold-start factor cache (see FwdParams::cold_factor): one entry per forward damping value seen.
struct ColdEntry {
  double delta;
  double* factor;  // bsize doubles
  int* ok;         // device flag: 1 if the factorisation succeeded
  cudaEvent_t ready;
  bool done;
};

struct IntoptProblem {
  intopt::ProblemDev dev{};
  IntoptInfo info{};
  const double* b_dev = nullptr;
  int device = 0;
  int sms = 0;
  int slots = 0;  // persistent CTAs per launch
  std::vector<void*> allocs;
  // lazily filled, guarded by `mu`; the problem data itself is immutable after create
  mutable std::mutex mu;
  mutable std::vector<ColdEntry> cold;
  mutable std::vector<void*> cold_allocs;
  IntoptDense* dense = nullptr;
  bool has_b = false;  // (dense path) the problem carries its own right-hand side
};


constexpr size_t kCounterBytes = 256;

inline size_t vals_ws_doubles(const IntoptProblem* p) {
  return (size_t)p->dev.nnz + (size_t)p->dev.n_csc_pad + (size_t)p->dev.n_asm_pad;
}
inline size_t ws_per_slot_doubles(const IntoptProblem* p) {
  size_t per = (size_t)intopt::ws_nvec_doubles(p->info.n);
  if (!p->info.factor_in_smem) per += (size_t)p->info.envelope;
  if (!p->info.gather_in_smem) per += 2 * (size_t)intopt::gather_len(p->info.n);
  per = (per + 1) & ~(size_t)1;
  if (p->dev.inst_values) per += vals_ws_doubles(p);  // this instance's copies of the value streams of A
  return (per + 1) & ~(size_t)1;
}


#ifdef INTOPT_PROFILE
unsigned long long* intopt_profile_phase_buckets();  // intopt_profile.cu
#endif

// ---- lock-step dense path (intopt_dense.cu) ------------------------------------------------------------
// Chosen at intopt_problem_create for dense problems whose factor does not fit one CTA's shared memory (or by
// INTOPT_PATH=dense / cta).  All entry points return 0 or a negative code and put the message into *err.
#include <string>
namespace intopt_dense {
bool wanted(int m, int n, long long nnz, int flags);
int create(IntoptDense** out, int m, int n, const int* rowptr, const int* colidx, const double* val, const double* b,
           std::string* err);
void destroy(IntoptDense* d);
size_t workspace_bytes(const IntoptDense* d, int B);
int forward(const IntoptDense* d, int B, double* c, long long c_stride, const double* b_batch, long long b_stride,
            const IntoptOpts* o, const double* init, const intopt::UserMap& um, double* xhat, double* state, int* status,
            int* iters, void* ws, size_t ws_bytes, cudaStream_t stream, std::string* err);
int backward(const IntoptDense* d, int B, int mode, double damping, int nvar, const double* state, const double* c,
             long long c_stride, const double* b_batch, long long b_stride, const intopt::UserMap& um, double* gx,
             double* gc, int* status, void* ws, size_t ws_bytes, cudaStream_t stream, std::string* err);
int normal_solve(const IntoptDense* d, int B, const double* dvec, double delta, double* rhs, int* fail, void* ws,
                 size_t ws_bytes, cudaStream_t stream, std::string* err);
int measure_dmma_peak(double* gflops_out, int repeats);
}  // namespace intopt_dense
