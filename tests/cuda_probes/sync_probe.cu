// Probe (not part of the product): cost of CTA-level synchronisation primitives on sm_100a, 256 threads, one or two CTAs
// per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_probe sync_probe.cu && ./sync_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}\n" ::"r"(
          (uint32_t)__cvta_generic_to_shared(b)),
      "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(256, 2) probe(int which, int reps, unsigned long long* out) {
  __shared__ uint64_t mb[2];
  __shared__ volatile int flag[16];
  __shared__ double buf[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&mb[0], 256);
    mbar_init(&mb[1], 32);
  }
  if (tid < 16) flag[tid] = 0;
  if (tid < 64) buf[tid] = 1.0;
  __syncthreads();
  double z = 1.0;
  const long long t0 = clock64();
  for (int r = 1; r <= reps; ++r) {
    switch (which) {
      case 0: __syncthreads(); break;
      case 1: asm volatile("bar.sync 1, 256;" ::: "memory"); break;
      case 2:  // split barrier: everyone arrives at once, waits later (two named barriers alternate)
        if (r & 1) {
          asm volatile("bar.arrive 1, 256;" ::: "memory");
          asm volatile("bar.sync 2, 256;" ::: "memory");
        } else {
          asm volatile("bar.arrive 2, 256;" ::: "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        break;
      case 3:  // mbarrier, everyone arrives and waits
        mbar_arrive(&mb[0]);
        mbar_wait(&mb[0], (r - 1) & 1);
        break;
      case 4:  // flag hand-off: warp 0 publishes, the others spin (one hop)
        if (warp == 0) {
          if (lane == 0) flag[0] = r;
        } else {
          while (flag[0] < r) {}
        }
        break;
      case 5:  // two hops through flags: warp 0 publishes, others see it and publish theirs, warp 0 collects
        if (warp == 0) {
          if (lane == 0) flag[0] = r;
          if (lane >= 1 && lane < 8) while (flag[lane] < r) {}
          __syncwarp();
        } else {
          while (flag[0] < r) {}
          if (lane == 0) flag[warp] = r;
        }
        break;
      case 6:  // a shared-memory round trip between two syncthreads (store, barrier, dependent load)
        buf[tid & 63] = z;
        __syncthreads();
        z = buf[(tid + 1) & 63] + 1.0;
        break;
      case 7:  // warp-level: syncwarp only
        __syncwarp();
        break;
      case 8:  // 8 dependent shared-memory loads
#pragma unroll
        for (int i = 0; i < 8; ++i) z = buf[((int)z + i) & 63];
        break;
      case 9:  // bar.sync of two warps only (64 threads), the others idle
        if (warp < 2) asm volatile("bar.sync 2, 64;" ::: "memory");
        break;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (tid == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) / reps + (z == 12345.0);
}

int main() {
  unsigned long long* out;
  cudaMallocManaged(&out, 1024 * sizeof(*out));
  const char* names[] = {"__syncthreads", "bar.sync 1,256", "bar.arrive x + bar.sync y (alternating)", "mbarrier arrive+wait (256)",
                         "flag hand-off, one hop", "flags, two hops", "st + syncthreads + dependent ld", "__syncwarp",
                         "8 dependent LDS", "bar.sync of 2 warps"};
  for (int grid : {1, 296}) {
    for (int w = 0; w < 10; ++w) {
      probe<<<grid, 256>>>(w, 2000, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long mx = 0, mn = ~0ull;
      for (int i = 0; i < grid; ++i) { mx = out[i] > mx ? out[i] : mx; mn = out[i] < mn ? out[i] : mn; }
      printf("grid %3d  %-40s %5llu .. %5llu cycles\n", grid, names[w], mn, mx);
      fflush(stdout);
    }
  }
  return 0;
}
