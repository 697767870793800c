// C ABI (include/intopt_b200.h) over the sm_100a solver kernels.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/intopt_b200.h"
#include "intopt_bind.hpp"
#include "intopt_kernels.cuh"
#include "intopt_internal.hpp"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU_TRY(expr)                                                                             \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(-100 - (int)e_, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
  } while (0)

constexpr int kMaxThreads = 512;
constexpr int kMaxDynSmem = 232448 - 256;  // 227 KB opt-in per CTA on sm_100, minus the kernels' static smem

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

// Decide what lives in shared memory.  Preference: factor + gather buffer; then factor only; then neither.

void plan_smem(const intopt::HostStructure& S, int* factor_in_smem, int* gather_in_smem, int* smem_bytes) {
  const int mi = intopt::meta_ints(S);
  const size_t fixed = intopt::smem_fixed_doubles(S) + (mi <= kMaxMetaInts ? (size_t)(mi + 1) / 2 : 0);
  const size_t fac = (size_t)S.bsize, gat = 2 * (size_t)intopt::gather_len(S.n);
  *factor_in_smem = 1;
  *gather_in_smem = 1;
  if ((fixed + fac + gat) * 8 > (size_t)kMaxDynSmem) {
    *gather_in_smem = 0;
    if ((fixed + fac) * 8 > (size_t)kMaxDynSmem) {
      *factor_in_smem = 0;
      if ((fixed + gat) * 8 <= (size_t)kMaxDynSmem) *gather_in_smem = 1;
    }
  }
  *smem_bytes = (int)((fixed + (*factor_in_smem ? fac : 0) + (*gather_in_smem ? gat : 0)) * 8);
}

void fill_info(const intopt::HostStructure& S, IntoptInfo* info) {
  info->m = S.m;
  info->n = S.n;
  info->nnz = S.nnz;
  info->envelope = S.bsize;
  info->dense_lower = (int)((long long)S.m * (S.m + 1) / 2);
  info->asm_terms = S.asm_terms;
  info->chol_macs = S.chol_macs;
  int g = 0;
  plan_smem(S, &info->factor_in_smem, &g, &info->smem_bytes);
  info->gather_in_smem = g;
  // CTA size: INTOPT_THREADS overrides.  Default 256 (two LPs resident per SM for the experiment shapes); small
  // problems take 128 threads so that four LPs share an SM; problems whose shared memory admits a single resident
  // CTA take 512 threads (the only way to put more warps on that SM).
  int thr = env_int("INTOPT_THREADS", 0);
  if (thr <= 0) {
    thr = 256;
    if (S.m <= 64 && info->smem_bytes <= (kMaxDynSmem + 1024) / 4 - 1024) thr = 128;
    if (info->smem_bytes > (kMaxDynSmem + 1024) / 2 - 1024) thr = 512;
  }
  info->threads = thr;
  info->ordering = S.ordering;
  info->scalar_envelope = S.sky_size;
  info->asm_padded = S.asm_padded;
  info->solve_macs = S.solve_macs;
}

// The lock-step dense path keeps no block-skyline structure: its "envelope" is the padded square it factors in place.
void fill_info_dense(int m, int n, int nnz, IntoptInfo* info) {
  std::memset(info, 0, sizeof(*info));
  const long long m64 = (m + 63) & ~63;
  info->m = m;
  info->n = n;
  info->nnz = nnz;
  info->envelope = (int)std::min<long long>(m64 * m64, 2147483647LL);
  info->dense_lower = (int)((long long)m * (m + 1) / 2);
  info->asm_terms = (long long)n * m * (m + 1) / 2;
  info->chol_macs = (long long)m * m * m / 6;
  info->solve_macs = (long long)m * m;
  info->threads = 512;
  info->path = 1;
}

int check_csr(int m, int n, const int* rp, const int* ci, const double* val) {
  if (m <= 0 || n <= 0 || !rp || !ci || !val) return -1;
  if (rp[0] != 0) return -2;
  for (int i = 0; i < m; ++i)
    if (rp[i + 1] < rp[i]) return -2;
  for (int e = 0; e < rp[m]; ++e)
    if (ci[e] < 0 || ci[e] >= n) return -3;
  return 0;
}

template <class T>
cudaError_t upload(const std::vector<T>& h, const T** d, std::vector<void*>& allocs) {
  *d = nullptr;
  if (h.empty()) return cudaSuccess;
  T* p = nullptr;
  cudaError_t e = cudaMalloc((void**)&p, h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  allocs.push_back(p);
  *d = p;
  return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

// Host-glue helper: dst[b, j] = idx[j] >= 0 ? scale[j] * src[b, idx[j]] : constv[j]   (scale NULL = 1, constv NULL = 0),
// converting between float and double on the way.  One launch replaces the cast + zero-fill + index / index_put
// chains that map the user's cost / gradient vectors to the standard form and the solution back.
template <class TS, class TD>
__global__ void __launch_bounds__(256) intopt_gather_convert_kernel(const TS* __restrict__ src, long long src_stride, int B,
                                                                    int n_out, const int* __restrict__ idx,
                                                                    const double* __restrict__ scale,
                                                                    const double* __restrict__ constv, TD* __restrict__ dst,
                                                                    long long dst_stride) {
  const long long total = (long long)B * n_out;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / n_out), j = (int)(e - (long long)b * n_out);
    const int k = idx ? idx[j] : j;
    double v;
    if (k >= 0) {
      v = (double)src[(long long)b * src_stride + k];
      if (scale) v *= scale[j];
    } else {
      v = constv ? constv[j] : 0.0;
    }
    dst[(long long)b * dst_stride + j] = (TD)v;
  }
}

// Register-resident DFMA chains: the measured FP64 roofline denominator.
__global__ void __launch_bounds__(256) intopt_fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c);
      a1 = fma(a1, m, c);
      a2 = fma(a2, m, c);
      a3 = fma(a3, m, c);
      a4 = fma(a4, m, c);
      a5 = fma(a5, m, c);
      a6 = fma(a6, m, c);
      a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

static std::vector<const void*> all_kernels() {
  std::vector<const void*> v;
  intopt_k::list_forward(v);
  intopt_k::list_forward_plain(v);
  intopt_k::list_tmem(v);
  intopt_k::list_backward(v);
  return v;
}
static const void* pick_kernel(bool forward, int threads, bool allsmem, bool plain = false) {
  return forward ? intopt_k::pick_forward(threads, allsmem, plain) : intopt_k::pick_backward(threads, allsmem);
}
static const void* pick_cold_kernel(int threads, bool allsmem) { return intopt_k::pick_cold(threads, allsmem); }
// factor, gather region AND schedule arrays in shared memory: the kernels compiled for exactly that
static bool all_in_smem(const IntoptProblem* p) {
  return p->info.factor_in_smem && p->info.gather_in_smem && p->dev.meta_ints <= kMaxMetaInts;
}
// The tensor-memory kernels (element-wise n-vectors in TMEM instead of the L2-resident workspace): 128-thread CTAs, at most
// four per SM (128 columns each), CTAs of up to 256 threads, at most two per SM (256 columns each), or larger CTAs, one per
// SM (all 512 columns); every thread's share of an n-vector must fit its columns.
static bool tmem_variant(const IntoptProblem* p) {
  const int th = p->info.threads;
  // (a CTA takes 128 columns per started group of four warps: smaller CTAs than 128 threads would over-subscribe them)
  // (measured: +2.7 % at n = 128 on 128-thread CTAs, -6 % at n = 32, where most lanes of the collective accesses are idle)
  const bool fits = th > 96 && (th <= 128 ? (p->info.ctas_per_sm <= 4 && p->info.n >= th)
                                          : th <= 256 ? p->info.ctas_per_sm <= 2 : p->info.ctas_per_sm == 1);
  return fits && (p->info.n + th - 1) / th <= intopt::NV_EMAX && env_int("INTOPT_TMEM", 1) != 0;
}

namespace {
// Returns the cold-start factor for `delta` (computing it on `stream` the first time) or nulls if unavailable.
int cold_factor_for(const IntoptProblem* p, double delta, cudaStream_t stream, const double** factor, const int** ok) {
  *factor = nullptr;
  *ok = nullptr;
  if (env_int("INTOPT_NO_COLD", 0)) return 0;
  std::lock_guard<std::mutex> lock(p->mu);
  for (ColdEntry& e : p->cold) {
    if (std::memcmp(&e.delta, &delta, sizeof(double)) != 0) continue;
    if (!e.done) {
      if (cudaEventQuery(e.ready) == cudaSuccess)
        e.done = true;
      else
        CU_TRY(cudaStreamWaitEvent(stream, e.ready, 0));
    }
    *factor = e.factor;
    *ok = e.ok;
    return 0;
  }
  if (p->cold.size() >= 8) return 0;  // unusual: many different dampings on one problem -> plain path
  ColdEntry e{};
  e.delta = delta;
  e.done = false;
  double* ws = nullptr;
  CU_TRY(cudaMalloc((void**)&e.factor, sizeof(double) * (size_t)p->info.envelope));
  p->cold_allocs.push_back(e.factor);
  CU_TRY(cudaMalloc((void**)&e.ok, sizeof(int)));
  p->cold_allocs.push_back(e.ok);
  CU_TRY(cudaMalloc((void**)&ws, sizeof(double) * ws_per_slot_doubles(p)));
  p->cold_allocs.push_back(ws);
  CU_TRY(cudaMemsetAsync(e.ok, 0, sizeof(int), stream));
  CU_TRY(cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming));
  {
    int fis = p->info.factor_in_smem, gis = p->info.gather_in_smem;
    void* args[7] = {(void*)&p->dev, (void*)&delta, (void*)&ws, (void*)&fis, (void*)&gis, (void*)&e.factor, (void*)&e.ok};
    CU_TRY(cudaLaunchKernel(pick_cold_kernel(p->info.threads, all_in_smem(p)), dim3(1), dim3(p->info.threads), args,
                            (size_t)p->info.smem_bytes, stream));
  }
  CU_TRY(cudaEventRecord(e.ready, stream));
  p->cold.push_back(e);
  *factor = e.factor;
  *ok = e.ok;
  return 0;
}
}  // namespace

extern "C" {

int intopt_version(void) { return INTOPT_B200_VERSION; }
const char* intopt_last_error(void) { return g_err.c_str(); }

int intopt_analyze(int m, int n, const int* rowptr, const int* colidx, const double* val, IntoptInfo* info,
                   int* row_perm) {
  if (!info) return fail(-1, "info is NULL");
  if (int rc0 = check_csr(m, n, rowptr, colidx, val)) return fail(rc0, "malformed CSR input");
  if (intopt_dense::wanted(m, n, rowptr[m], 0)) {
    fill_info_dense(m, n, rowptr[m], info);
    if (row_perm)
      for (int i = 0; i < m; ++i) row_perm[i] = i;
    return 0;
  }
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rowptr, colidx, val, S);
  if (rc == -5) return fail(rc, "problem too large for the gather-list assembly (sum over columns of nnz*(nnz+1)/2 > 2^28)");
  if (rc) return fail(rc, "malformed CSR input");
  fill_info(S, info);
  if (row_perm) std::memcpy(row_perm, S.perm.data(), sizeof(int) * m);
  return 0;
}

int intopt_problem_create(IntoptProblem** out, int m, int n, const int* rowptr, const int* colidx,
                          const double* val, const double* b, int flags) {
  if (!out) return fail(-1, "out is NULL");
  *out = nullptr;
  if (int rc0 = check_csr(m, n, rowptr, colidx, val)) return fail(rc0, "malformed CSR input");
  if (intopt_dense::wanted(m, n, rowptr[m], flags)) {
    int ndev0 = 0;
    if (cudaGetDeviceCount(&ndev0) != cudaSuccess || ndev0 <= 0)
      return fail(-10, "no CUDA device: intopt_b200 has no CPU fallback");
    IntoptProblem* p = new (std::nothrow) IntoptProblem();
    if (!p) return fail(-2, "out of host memory");
    fill_info_dense(m, n, rowptr[m], &p->info);
    std::string msg;
    int rcd = intopt_dense::create(&p->dense, m, n, rowptr, colidx, val, b, &msg);
    if (rcd) {
      delete p;
      return fail(rcd, msg);
    }
    CU_TRY(cudaGetDevice(&p->device));
    CU_TRY(cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, p->device));
    p->info.ctas_per_sm = 1;
    p->has_b = b != nullptr;
    *out = p;
    return 0;
  }
  intopt::HostStructure S;
  int rc = intopt::analyze(m, n, rowptr, colidx, val, S, -1, (flags & INTOPT_FLAG_INSTANCE_VALUES) != 0);
  if (rc == -5) return fail(rc, "problem too large for the gather-list assembly (sum over columns of nnz*(nnz+1)/2 > 2^28)");
  if (rc) return fail(rc, "malformed CSR input");
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev <= 0) return fail(-10, "no CUDA device: intopt_b200 has no CPU fallback");
  IntoptProblem* p = new (std::nothrow) IntoptProblem();
  if (!p) return fail(-2, "out of host memory");
  cudaError_t err = cudaSuccess;
  auto up = [&](const auto& v) {
    using T = typename std::decay<decltype(v)>::type::value_type;
    const T* d = nullptr;
    cudaError_t e = upload(v, &d, p->allocs);
    if (e != cudaSuccess) err = e;
    return d;
  };
  fill_info(S, &p->info);
  intopt::bind_problem(S, p->dev, up);
  if (b) {
    std::vector<double> bperm(m);
    for (int i = 0; i < m; ++i) bperm[i] = b[S.perm[i]];
    p->b_dev = up(bperm);
  }
  if (err != cudaSuccess) {
    intopt_problem_destroy(p);
    return fail(-100 - (int)err, std::string("device upload: ") + cudaGetErrorString(err));
  }
  CU_TRY(cudaGetDevice(&p->device));
  CU_TRY(cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, p->device));
  if (p->info.threads < 32 || p->info.threads > kMaxThreads || p->info.threads % 32) {
    intopt_problem_destroy(p);
    return fail(-12, "INTOPT_THREADS must be a multiple of 32 in [32, 512]");
  }
  for (const void* k : all_kernels())
    CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  int occ = 0;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &occ, pick_kernel(true, p->info.threads, all_in_smem(p)), p->info.threads,
      (size_t)p->info.smem_bytes));
  if (occ < 1) {
    intopt_problem_destroy(p);
    return fail(-11, "solver kernel does not fit on an SM (shared memory / registers)");
  }
  occ = std::min(occ, env_int("INTOPT_MAX_CTAS_PER_SM", 8));
  p->info.ctas_per_sm = occ;
  p->info.tmem = tmem_variant(p) ? 1 : 0;
  p->slots = p->sms * occ;
  CU_TRY(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

int intopt_problem_destroy(IntoptProblem* p) {
  if (!p) return 0;
  intopt_dense::destroy(p->dense);
  for (void* a : p->allocs)
    if (a) cudaFree(a);
  for (void* a : p->cold_allocs)
    if (a) cudaFree(a);
  for (ColdEntry& e : p->cold) cudaEventDestroy(e.ready);
  delete p;
  return 0;
}

int intopt_problem_info(const IntoptProblem* p, IntoptInfo* info) {
  if (!p || !info) return fail(-1, "NULL argument");
  *info = p->info;
  return 0;
}

size_t intopt_workspace_bytes(const IntoptProblem* p, int B) {
  if (!p || B <= 0) return 0;
  if (p->dense) return intopt_dense::workspace_bytes(p->dense, B);
  const size_t slots = (size_t)std::min(B, p->slots);
  return kCounterBytes + ws_per_slot_doubles(p) * sizeof(double) * slots;
}

int intopt_forward(const IntoptProblem* p, int B, const double* c, long long c_stride, const double* b_batch,
                   long long b_stride, const IntoptOpts* o, const double* init, double* xhat, double* state,
                   int* status, int* iters, void* ws, size_t ws_bytes, void* stream) {
  return intopt_forward_values(p, B, c, c_stride, b_batch, b_stride, nullptr, 0, o, init, xhat, state, status, iters, ws,
                               ws_bytes, stream);
}

int intopt_forward_values(const IntoptProblem* p, int B, const double* c, long long c_stride, const double* b_batch,
                          long long b_stride, const double* a_vals, long long a_stride, const IntoptOpts* o,
                          const double* init, double* xhat, double* state, int* status, int* iters, void* ws,
                          size_t ws_bytes, void* stream) {
  return intopt_forward_mapped(p, B, const_cast<double*>(c), c_stride, b_batch, b_stride, a_vals, a_stride, o, init, nullptr,
                               xhat, state, status, iters, ws, ws_bytes, stream);
}

static void fill_user_map(intopt::UserMap& u, const IntoptUserMap* m) {
  u = intopt::UserMap{};
  if (!m) return;
  u.in = m->in_dev;
  u.in_f32 = m->in_is_f32;
  u.in_stride = m->in_stride;
  u.in_idx = m->in_idx_dev;
  u.in_scale = m->in_scale_dev;
  u.out = m->out_dev;
  u.out_f32 = m->out_is_f32;
  u.out_stride = m->out_stride;
  u.n_out = m->n_out;
  u.out_idx = m->out_idx_dev;
  u.out_const = m->out_const_dev;
}

int intopt_forward_mapped(const IntoptProblem* p, int B, double* c, long long c_stride, const double* b_batch,
                          long long b_stride, const double* a_vals, long long a_stride, const IntoptOpts* o,
                          const double* init, const IntoptUserMap* um, double* xhat, double* state, int* status,
                          int* iters, void* ws, size_t ws_bytes, void* stream) {
  if (!p || !c || !o || !state || !status || !iters || !ws) return fail(-1, "NULL argument");
  if (!xhat && !(um && um->out_dev)) return fail(-1, "neither xhat_dev nor a mapped output was given");
  if (um && um->in_dev && (!um->in_idx_dev || c_stride < p->info.n)) return fail(-8, "mapped input needs in_idx_dev and one writable c row per instance");
  if (um && um->out_dev && !um->out_idx_dev) return fail(-8, "mapped output needs out_idx_dev");
  if (a_vals && !p->dev.inst_values) return fail(-7, "per-instance values need a problem created with INTOPT_FLAG_INSTANCE_VALUES");
  if (!a_vals && p->dev.inst_values) return fail(-7, "this problem was created for per-instance values: pass a_vals_dev");
  if (B <= 0) return 0;
  if (p->dense) {
    if (!p->has_b && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
    intopt::UserMap umap;
    fill_user_map(umap, um);
    std::string msg;
    const int rc = intopt_dense::forward(p->dense, B, c, c_stride, b_batch, b_stride, o, init, umap, xhat, state, status,
                                         iters, ws, ws_bytes, (cudaStream_t)stream, &msg);
    return rc ? fail(rc, msg) : 0;
  }
  if (!p->b_dev && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
  if (ws_bytes < intopt_workspace_bytes(p, B)) return fail(-4, "workspace too small");
  intopt::FwdParams a{};
  a.B = B;
  a.c = c;
  a.c_stride = c_stride;
  a.b_batch = b_batch;
  a.b_stride = b_stride;
  a.b_problem = p->b_dev;
  a.init = init;
  fill_user_map(a.um, um);
  a.a_vals = a_vals;
  a.a_stride = a_stride;
  a.xhat = xhat;
  a.state = state;
  a.status = status;
  a.iters = iters;
  a.counter = (int*)ws;
  a.ws = (double*)((char*)ws + kCounterBytes);
  a.factor_in_smem = p->info.factor_in_smem;
  a.gather_in_smem = p->info.gather_in_smem;
  a.ws_stride = (long long)ws_per_slot_doubles(p);
  // the value streams sit at the end of every slot
  a.vals_ws = a.ws + (a.ws_stride - (long long)((vals_ws_doubles(p) + 1) & ~(size_t)1));
  a.vals_ws_stride = a.ws_stride;
  a.alpha0 = o->alpha0;
  a.tol = o->tol;
  a.thr = o->thr;
  a.damping = o->damping_fwd;
  a.max_iter = o->max_iter;
  a.plain = o->plain != 0;
  a.beta = o->beta;
#ifdef INTOPT_PROFILE
  a.phase_cycles = intopt_profile_phase_buckets();  // null unless intopt_debug_phase_cycles(1, ...) switched it on
  a.debug = env_int("INTOPT_DEBUG", 0);
#endif
  if (!init && !a_vals) {  // (the cold-start factor is a constant of the problem only if A is)
    int rc = cold_factor_for(p, a.damping, (cudaStream_t)stream, &a.cold_factor, &a.cold_ok);
    if (rc) return rc;
  }
  const int grid = std::min(B, p->slots);
  CU_TRY(cudaMemsetAsync(ws, 0, kCounterBytes, (cudaStream_t)stream));
  {
    void* args[2] = {(void*)&p->dev, (void*)&a};
    const bool allsmem = all_in_smem(p);
    const int th = p->info.threads;
    const void* kern = (!a.plain && tmem_variant(p)) ? intopt_k::pick_forward_tmem(th, allsmem) : pick_kernel(true, th, allsmem, a.plain != 0);
    CU_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(th), args, (size_t)p->info.smem_bytes, (cudaStream_t)stream));
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

int intopt_backward(const IntoptProblem* p, int B, int mode, double damping, int nvar, const double* state,
                    const double* c, long long c_stride, const double* b_batch, long long b_stride,
                    const double* gx, double* gc, int* status, void* ws, size_t ws_bytes, void* stream) {
  return intopt_backward_values(p, B, mode, damping, nvar, state, c, c_stride, b_batch, b_stride, nullptr, 0, gx, gc, status,
                                ws, ws_bytes, stream);
}

int intopt_backward_values(const IntoptProblem* p, int B, int mode, double damping, int nvar, const double* state,
                           const double* c, long long c_stride, const double* b_batch, long long b_stride,
                           const double* a_vals, long long a_stride, const double* gx, double* gc, int* status, void* ws,
                           size_t ws_bytes, void* stream) {
  return intopt_backward_mapped(p, B, mode, damping, nvar, state, c, c_stride, b_batch, b_stride, a_vals, a_stride, nullptr,
                                const_cast<double*>(gx), gc, status, ws, ws_bytes, stream);
}

int intopt_backward_mapped(const IntoptProblem* p, int B, int mode, double damping, int nvar, const double* state,
                           const double* c, long long c_stride, const double* b_batch, long long b_stride,
                           const double* a_vals, long long a_stride, const IntoptUserMap* um, double* gx, double* gc,
                           int* status, void* ws, size_t ws_bytes, void* stream) {
  if (!p || !c || !state || !gx || !gc || !status || !ws) return fail(-1, "NULL argument");
  if (um && ((um->in_dev && !um->in_idx_dev) || (um->out_dev && !um->out_idx_dev))) return fail(-8, "mapped input / output needs its index array");
  if ((a_vals != nullptr) != (p->dev.inst_values != 0)) return fail(-7, "a_vals_dev must be given exactly for problems created with INTOPT_FLAG_INSTANCE_VALUES");
  if (B <= 0) return 0;
  if (mode != 0 && mode != 1) return fail(-5, "mode must be 0 (whole) or 1 (alter)");
  if (nvar < 0 || nvar > p->info.n) return fail(-6, "nvar out of range");
  if (p->dense) {
    if (mode == 0 && !p->has_b && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
    intopt::UserMap umap;
    fill_user_map(umap, um);
    std::string msg;
    const int rc = intopt_dense::backward(p->dense, B, mode, damping, nvar, state, c, c_stride, b_batch, b_stride, umap, gx,
                                          gc, status, ws, ws_bytes, (cudaStream_t)stream, &msg);
    return rc ? fail(rc, msg) : 0;
  }
  if (mode == 0 && !p->b_dev && !b_batch) return fail(-3, "problem has no b and no b_batch was given");
  if (ws_bytes < intopt_workspace_bytes(p, B)) return fail(-4, "workspace too small");
  intopt::BwdParams a{};
  a.B = B;
  a.mode = mode;
  a.nvar = nvar;
  a.damping = damping;
  a.state = state;
  a.c = c;
  a.c_stride = c_stride;
  a.b_batch = b_batch;
  a.b_stride = b_stride;
  a.b_problem = p->b_dev;
  a.gx = gx;
  fill_user_map(a.um, um);
  a.a_vals = a_vals;
  a.a_stride = a_stride;
  a.gc = gc;
  a.status = status;
  a.counter = (int*)ws;
  a.ws = (double*)((char*)ws + kCounterBytes);
  a.factor_in_smem = p->info.factor_in_smem;
  a.gather_in_smem = p->info.gather_in_smem;
  a.ws_stride = (long long)ws_per_slot_doubles(p);
  // the value streams sit at the end of every slot
  a.vals_ws = a.ws + (a.ws_stride - (long long)((vals_ws_doubles(p) + 1) & ~(size_t)1));
  a.vals_ws_stride = a.ws_stride;
  const int grid = std::min(B, p->slots);
  CU_TRY(cudaMemsetAsync(ws, 0, kCounterBytes, (cudaStream_t)stream));
  {
    void* args[2] = {(void*)&p->dev, (void*)&a};
    const bool allsmem = all_in_smem(p);
    const void* kern = tmem_variant(p) ? intopt_k::pick_backward_tmem(p->info.threads, allsmem) : pick_kernel(false, p->info.threads, allsmem);
    CU_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(p->info.threads), args, (size_t)p->info.smem_bytes, (cudaStream_t)stream));
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

int intopt_gather_convert(const void* src, int src_is_f32, long long src_stride, int B, int n_out, const int* idx,
                          const double* scale, const double* constv, void* dst, int dst_is_f32, long long dst_stride,
                          void* stream) {
  if (!src || !dst || B < 0 || n_out < 0) return fail(-1, "bad argument");
  const long long total = (long long)B * n_out;
  if (total == 0) return 0;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_is_f32 && dst_is_f32)
    intopt_gather_convert_kernel<float, float><<<blocks, threads, 0, st>>>((const float*)src, src_stride, B, n_out, idx, scale, constv, (float*)dst, dst_stride);
  else if (src_is_f32)
    intopt_gather_convert_kernel<float, double><<<blocks, threads, 0, st>>>((const float*)src, src_stride, B, n_out, idx, scale, constv, (double*)dst, dst_stride);
  else if (dst_is_f32)
    intopt_gather_convert_kernel<double, float><<<blocks, threads, 0, st>>>((const double*)src, src_stride, B, n_out, idx, scale, constv, (float*)dst, dst_stride);
  else
    intopt_gather_convert_kernel<double, double><<<blocks, threads, 0, st>>>((const double*)src, src_stride, B, n_out, idx, scale, constv, (double*)dst, dst_stride);
  CU_TRY(cudaGetLastError());
  return 0;
}

int intopt_dense_normal_solve(const IntoptProblem* p, int B, const double* d, double delta, double* rhs, int* fail_flags,
                              void* ws, size_t ws_bytes, void* stream) {
  if (!p || !d || !rhs || !ws) return fail(-1, "NULL argument");
  if (!p->dense) return fail(-13, "this problem runs on the persistent-CTA path (IntoptInfo.path == 0)");
  if (B <= 0) return 0;
  std::string msg;
  const int rc = intopt_dense::normal_solve(p->dense, B, d, delta, rhs, fail_flags, ws, ws_bytes, (cudaStream_t)stream, &msg);
  return rc ? fail(rc, msg) : 0;
}

int intopt_measure_dmma_peak(double* gflops_out, int repeats) {
  if (!gflops_out) return fail(-1, "NULL argument");
  const int rc = intopt_dense::measure_dmma_peak(gflops_out, repeats);
  return rc ? fail(rc, "DMMA peak measurement failed") : 0;
}

int intopt_measure_fp64_peak(double* gflops_out, int repeats) {
  if (!gflops_out) return fail(-1, "NULL argument");
  int dev = 0, sms = 0;
  CU_TRY(cudaGetDevice(&dev));
  CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  CU_TRY(cudaMalloc((void**)&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < (repeats > 0 ? repeats : 5) + 2; ++r) {
    CU_TRY(cudaEventRecord(e0));
    intopt_fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + r);
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    double gf = flops / (ms * 1e-3) * 1e-9;
    if (r >= 2 && gf > best) best = gf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *gflops_out = best;
  return 0;
}

}  // extern "C"
