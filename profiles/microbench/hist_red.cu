// Micro-benchmark: what does one histogram update cost on a B200, chip-wide?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_red hist_red.cu && ./hist_red
// All modes run one CTA of 512 threads on every SM at once (the scan kernel's shape); the figure is
// SM cycles per warp instruction (32 updates), averaged over the CTAs.
//   g64/g32   red.global.add.u64 / .u32 to a random counter of `n` (L2-resident)
//   s32       red.shared.add.u32 to a random word of a 14K-counter table (random banks)
//   a32       atom.shared.add.u32 (returning) likewise
//   cN        red.shared::cluster.add.u32 to a random word of a random CTA of a cluster of N
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

constexpr int THREADS = 512;
constexpr uint32_t SWORDS = 14 * 1024;  // 56 KB of counters per CTA

__device__ __forceinline__ uint32_t lcg(uint32_t& x) {
  x = x * 1664525u + 1013904223u;
  return x >> 8;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(unsigned long long* out, unsigned long long* g64, uint32_t n, int iters, uint32_t seed) {
  extern __shared__ uint32_t sm[];
  for (uint32_t i = threadIdx.x; i < SWORDS; i += blockDim.x) sm[i] = 0;
  uint32_t csize = 1, crank = 0;
  if (MODE >= 10) {
    cg::cluster_group cl = cg::this_cluster();
    csize = cl.num_blocks();
    crank = cl.block_rank();
    cl.sync();
  } else {
    __syncthreads();
  }
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  uint32_t x = seed + (threadIdx.x + blockIdx.x * 977u) * 2654435761u;
  uint32_t sink = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const uint32_t r = lcg(x);
      if (MODE == 0) {
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(g64 + (r % n)), "l"(1ull) : "memory");
      } else if (MODE == 1) {
        asm volatile("red.global.add.u32 [%0], %1;" ::"l"((uint32_t*)g64 + (r % n)), "r"(1u) : "memory");
      } else if (MODE == 2) {
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + (r % SWORDS) * 4u), "r"(1u) : "memory");
      } else if (MODE == 3) {
        uint32_t o;
        asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(base + (r % SWORDS) * 4u), "r"(1u) : "memory");
        sink += o;
      } else {
        const uint32_t tgt = (r >> 4) % csize;
        uint32_t local = base + ((r >> 7) % SWORDS) * 4u, remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(tgt));
        asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
      }
    }
  }
  if (MODE >= 10) {
    cg::this_cluster().sync();
  } else {
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (sink == 0x12345) out[0] = sink;
  // checksum of the shared counters so nothing is optimised away (and the totals can be verified)
  unsigned long long s = 0;
  for (uint32_t i = threadIdx.x; i < SWORDS; i += blockDim.x) s += sm[i];
  if (MODE >= 2) atomicAdd(out + 1024, s);
  (void)crank;
}

template <int MODE>
void run(const char* name, uint32_t n, int cluster, int sms) {
  unsigned long long *d, *g;
  cudaMalloc(&d, 2048 * 8);
  cudaMemset(d, 0, 2048 * 8);
  cudaMalloc(&g, (size_t)(n ? n : 1) * 8);
  cudaMemset(g, 0, (size_t)(n ? n : 1) * 8);
  const int iters = 400;
  const size_t smem = SWORDS * 4;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = sms;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.numAttrs = 0;
  if (cluster > 1) {
    if (cluster > 8) cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(cluster);
    int nc = 0;
    cudaOccupancyMaxActiveClusters(&nc, k<MODE>, &cfg);
    grid = nc * cluster;
    printf("  (cluster %d: %d clusters co-resident = %d CTAs)\n", cluster, nc, grid);
    if (grid == 0) return;
  }
  cfg.gridDim = dim3(grid);
  for (int rep = 0; rep < 2; rep++) {
    cudaMemset(d, 0, 2048 * 8);
    cudaLaunchKernelEx(&cfg, k<MODE>, d, g, n ? n : 1u, iters, (uint32_t)(1 + 6 * rep));
    cudaDeviceSynchronize();
  }
  std::vector<unsigned long long> h(2048);
  cudaMemcpy(h.data(), d, 2048 * 8, cudaMemcpyDeviceToHost);
  double sum = 0, mx = 0;
  for (int i = 0; i < grid; i++) {
    sum += (double)h[i];
    mx = mx > (double)h[i] ? mx : (double)h[i];
  }
  const double ops = (double)iters * 16 * (THREADS / 32);
  const double expect = (double)iters * 16 * THREADS * grid;
  printf("%-44s n=%8u CTAs %3d: %6.2f cycles per warp-instruction per SM (max %6.2f)  check %s  err=%d\n", name, n, grid,
         sum / grid / ops, mx / ops, MODE >= 2 ? ((double)h[1024] == expect ? "ok" : "MISMATCH") : "-", (int)cudaGetLastError());
  cudaFree(d);
  cudaFree(g);
}

int main() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  run<0>("red.global.add.u64 random counter", 117234, 1, sms);
  run<0>("red.global.add.u64 random counter", 1002, 1, sms);
  run<0>("red.global.add.u64 random counter", 1u << 24, 1, sms);
  run<1>("red.global.add.u32 random counter", 117234, 1, sms);
  run<1>("red.global.add.u32 random counter", 1002, 1, sms);
  run<2>("red.shared.add.u32 random word", 0, 1, sms);
  run<3>("atom.shared.add.u32 random word (returning)", 0, 1, sms);
  run<10>("red.shared::cluster.add.u32 random CTA+word", 0, 2, sms);
  run<10>("red.shared::cluster.add.u32 random CTA+word", 0, 4, sms);
  run<10>("red.shared::cluster.add.u32 random CTA+word", 0, 8, sms);
  run<10>("red.shared::cluster.add.u32 random CTA+word", 0, 16, sms);
  return 0;
}
