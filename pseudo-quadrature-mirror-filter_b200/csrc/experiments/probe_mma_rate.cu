// Tensor-pipe throughput of SMALL tcgen05.mma instructions (M128 x N x K16, fp16) as used by hankel16.cuh:
// 48 MMAs per tile issued by one thread; how does the time depend on the number of independent TMEM accumulators
// (dependent accumulation chains) and on N?
#include <cuda_runtime.h>
#include <cstdio>
#include "../ptx.cuh"
using namespace pqmf::ptx;

__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}

__global__ void __launch_bounds__(128) rate_kernel(long long* out, int chains, int n1, int n2, int per_tile, int tiles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 40000 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c002c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t a = smem_u32(sm), b = smem_u32(sm + 8192);
  if (tid == 0) {
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      for (int i = 0; i < per_tile / 2; ++i) umma_f16(tm + 32 * (i % chains), desc_sw32(a + 32 * (i % 24)), umma_desc(b + (i % 24) * 1024, 512, 128), umma_idesc_f16(128, n1), true);
      for (int i = 0; i < per_tile / 2; ++i) umma_f16(tm + 32 * (i % chains), desc_sw32(a + 4096 + 32 * (i % 24)), umma_desc(b + (i % 24) * 1024, 512, 128), umma_idesc_f16(128, n2), true);
      umma_commit(&bar);
      mbar_wait(&bar, t & 1);
    }
    out[blockIdx.x] = (clock64() - t0) / tiles;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

void run(int chains, int n1, int n2, int per_tile, int ctas) {
  long long* d; cudaMalloc(&d, 148 * 8 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  rate_kernel<<<148 * ctas, 128, 40000>>>(d, chains, n1, n2, per_tile, 200);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("chains %d  N %2d/%2d  %2d MMAs/tile  ctas/sm %d : %6lld cyc per tile per CTA -> %5.1f cyc/MMA per SM  [%s]\n", chains, n1, n2, per_tile, ctas, h[0],
         (double)h[0] / per_tile / ctas, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int ctas = 1; ctas <= 2; ++ctas) {
    run(1, 32, 16, 48, ctas);
    run(2, 32, 16, 48, ctas);
    run(4, 32, 16, 48, ctas);
    run(8, 32, 16, 48, ctas);
    run(4, 32, 32, 48, ctas);
    run(4, 64, 64, 48, ctas);
    run(4, 128, 128, 48, ctas);
    run(4, 256, 256, 48, ctas);
  }
  return 0;
}
