// Micro-benchmark: throughput of shared-memory reductions (red.shared.add.u32 / atom.shared.add.u32)
// on one SM, as used by the replicated accumulators of the scan kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_shared red_shared.cu && ./red_shared
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void sred(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t satom(uint32_t a, uint32_t v) {
  uint32_t o;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
  return o;
}

// mode 0: red, lane-private word (bank = lane), slot varies per step
// mode 1: red, all lanes of a warp on random slots of a 64-slot table, R=32 replicas (bank = lane)
// mode 2: atom (returning) as mode 1
// mode 3: plain LDS/ADD/STS read-modify-write as mode 1 (no atomicity; for comparison)
// mode 4: red, R=8 replicas (4 lanes share a replica index; random slots)
// mode 5: red, every lane the same address
template <int MODE>
__global__ void k(unsigned long long* out, int iters, uint32_t seed) {
  extern __shared__ uint32_t sm[];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 66 * 8 * 32; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  uint32_t x = seed + threadIdx.x * 2654435761u;
  uint32_t sink = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x = x * 1664525u + 1013904223u;
      const uint32_t e = (x >> 20) & 63u;
      uint32_t addr;
      if (MODE == 0) addr = base + (((uint32_t)(it + u) & 63u) * 224u + lane) * 4u;
      else if (MODE == 4) addr = base + (e * 56u + (lane & 7)) * 4u;
      else if (MODE == 5) addr = base + e * 4u * 0u;
      else addr = base + (e * 224u + lane) * 4u;
      if (MODE == 2) sink += satom(addr, x);
      else if (MODE == 3) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        v += x;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
      } else sred(addr, x);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (sink == 0x12345) out[1] = sink;
}

template <int MODE>
void run(const char* name, int threads) {
  unsigned long long* d;
  cudaMalloc(&d, 64);
  const int iters = 2000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 8 * 32 * 4);
  k<MODE><<<1, threads, 66 * 8 * 32 * 4>>>(d, iters, 1);
  k<MODE><<<1, threads, 66 * 8 * 32 * 4>>>(d, iters, 7);
  unsigned long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double ops = (double)iters * 16 * (threads / 32);
  printf("%-46s warps %2d: %.2f cycles per warp-instruction (SM-wide)  err=%d\n", name, threads / 32, h / ops, (int)cudaGetLastError());
  cudaFree(d);
}

int main() {
  for (int threads : {128, 512, 1024}) {
    run<0>("red  lane-private word, same slot per warp", threads);
    run<1>("red  R=32 replicas, random slot per lane", threads);
    run<2>("atom R=32 replicas, random slot per lane", threads);
    run<3>("ld/add/st R=32 (non-atomic), random slot", threads);
    run<4>("red  R=8 replicas, random slot per lane", threads);
    run<5>("red  all lanes one address", threads);
  }
  return 0;
}
