// Probe (not part of the product): TMEM as a per-thread scratch store for fp64 values on sm_100a.
//   * two 256-thread CTAs per SM, each allocating 256 TMEM columns (tcgen05.alloc), i.e. all 512 columns of the SM;
//   * thread (warp w, lane l) owns TMEM lane 32 (w % 4) + l and the column window [128 (w / 4), +128): 64 doubles;
//   * tcgen05.st / tcgen05.ld (.32x32b.x2: one double per thread) round trip, checked, and timed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st(uint32_t taddr, double v) {
  const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void tm_ld_issue(uint32_t taddr, uint32_t& lo, uint32_t& hi) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ double tm_value(uint32_t& lo, uint32_t& hi) {
  asm volatile("" : "+r"(lo), "+r"(hi));  // (after tm_wait_ld: pins the use behind the wait)
  return __hiloint2double((int)hi, (int)lo);
}

__global__ void __launch_bounds__(256, 2) probe(double* out, long long* cycles, int reps) {
  __shared__ uint32_t s_taddr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_taddr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_taddr;
  const uint32_t mine = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * (warp >> 2));
  // store 63 doubles, read them back
  for (int j = 0; j < 63; ++j) tm_st(mine + 2 * j, 1000.0 * blockIdx.x + tid + 0.001 * j);
  tm_wait_st();
  double bad = 0.0;
  for (int j = 0; j < 63; ++j) {
    uint32_t lo, hi;
    tm_ld_issue(mine + 2 * j, lo, hi);
    tm_wait_ld();
    const double v = tm_value(lo, hi);
    if (v != 1000.0 * blockIdx.x + tid + 0.001 * j) bad += 1.0;
  }
  // latency: dependent chain ld -> add -> st -> ld ...
  double acc = 0.0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t lo, hi;
    tm_ld_issue(mine + 2 * (r & 31), lo, hi);
    tm_wait_ld();
    acc += tm_value(lo, hi);
  }
  long long t1 = clock64();
  // throughput: 7 loads in flight, one wait (what an n-loop does), all 8 warps at once
  __syncthreads();
  long long t2 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t lo[7], hi[7];
#pragma unroll
    for (int u = 0; u < 7; ++u) tm_ld_issue(mine + 2 * ((7 * r + u) % 63), lo[u], hi[u]);
    tm_wait_ld();
#pragma unroll
    for (int u = 0; u < 7; ++u) acc += tm_value(lo[u], hi[u]);
  }
  long long t3 = clock64();
  // store throughput
  __syncthreads();
  long long t4 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int u = 0; u < 7; ++u) tm_st(mine + 2 * ((7 * r + u) % 63), acc + u);
    tm_wait_st();
  }
  long long t5 = clock64();
  out[blockIdx.x * blockDim.x + tid] = bad + 1e-300 * acc;
  if (tid == 0) {
    cycles[3 * blockIdx.x + 0] = (t1 - t0) / reps;
    cycles[3 * blockIdx.x + 1] = (t3 - t2) / reps;
    cycles[3 * blockIdx.x + 2] = (t5 - t4) / reps;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory");
}

int main() {
  const int grid = 296, reps = 2000;
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaMalloc(&cyc, sizeof(long long) * grid * 3);
  probe<<<grid, 256>>>(out, cyc, reps);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  static double h[296 * 256];
  static long long hc[296 * 3];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  double bad = 0;
  for (int i = 0; i < grid * 256; ++i) bad += h[i] > 0.5 ? 1 : 0;
  printf("mismatching threads: %.0f of %d\n", bad, grid * 256);
  printf("CTA 0:   dependent ld+wait %lld cycles;  7 ld + wait (8 warps x 2 CTAs/SM) %lld cycles;  7 st + wait %lld cycles\n", hc[0], hc[1], hc[2]);
  printf("CTA 200: dependent ld+wait %lld cycles;  7 ld + wait %lld cycles;  7 st + wait %lld cycles\n", hc[600], hc[601], hc[602]);
  return 0;
}
