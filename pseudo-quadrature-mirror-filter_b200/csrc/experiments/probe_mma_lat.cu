// How long does one tile's UMMA group take from issue to mbarrier completion, for the operand layouts the kernels use?
// (tensor-pipe utilisation is ~1 %, so only the LATENCY of the group matters: every warp waits on it once per tile)
#include <cuda_runtime.h>
#include <cstdio>
#include "../ptx.cuh"
using namespace pqmf::ptx;

template <int KIND /*0 f16, 1 tf32*/>
__global__ void __launch_bounds__(128) lat_kernel(long long* out, int sbo, int lbo, int n1, int n2, int ksteps, int reps, int nctas_note) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 56000 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t a = smem_u32(sm), b = smem_u32(sm + 40000);
  long long tot = 0, mx = 0;
  if (tid == 0) {
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int ks = 0; ks < ksteps; ++ks) {
        if (KIND == 0) umma_f16(tm, umma_desc(a + ks * 2 * lbo, lbo, sbo), umma_desc(b + ks * 2 * 512, 512, 128), umma_idesc_f16(128, n1), ks != 0);
        else umma_tf32(tm, umma_desc(a + ks * 2 * lbo, lbo, sbo), umma_desc(b + ks * 2 * 512, 512, 128), umma_idesc_tf32(128, n1), ks != 0);
      }
      for (int ks = 0; ks < ksteps && n2 > 0; ++ks) {
        if (KIND == 0) umma_f16(tm, umma_desc(a + 12000 + ks * 2 * lbo, lbo, sbo), umma_desc(b + 4096 + ks * 2 * 256, 256, 128), umma_idesc_f16(128, n2), true);
        else umma_tf32(tm, umma_desc(a + 12000 + ks * 2 * lbo, lbo, sbo), umma_desc(b + 4096 + ks * 2 * 256, 256, 128), umma_idesc_tf32(128, n2), true);
      }
      umma_commit(&bar);
      mbar_wait(&bar, r & 1);
      const long long dt = clock64() - t0;
      if (r > 2) { tot += dt; if (dt > mx) mx = dt; }
    }
    out[blockIdx.x * 2] = tot / (reps - 3);
    out[blockIdx.x * 2 + 1] = mx;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

template <int KIND>
void run(const char* name, int sbo, int lbo, int n1, int n2, int ksteps, int ctas_per_sm) {
  long long* d; cudaMalloc(&d, 148 * 8 * 2 * sizeof(long long));
  auto kern = lat_kernel<KIND>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 56000);
  kern<<<148 * ctas_per_sm, 128, 56000>>>(d, sbo, lbo, n1, n2, ksteps, 200, ctas_per_sm);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-46s sbo %4d lbo %5d N %2d+%2d k-steps %d  ctas/sm %d : avg %5lld cyc, max %5lld  [%s]\n", name, sbo, lbo, n1, n2, ksteps, ctas_per_sm, h[0], h[1], cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<0>("f16 analysis layout (padded)", 160, 2576, 32, 16, 2, 1);
  run<0>("f16 analysis layout (padded)", 160, 2576, 32, 16, 2, 4);
  run<0>("f16 dense layout", 128, 2048, 32, 16, 2, 1);
  run<0>("f16 dense layout", 128, 2048, 32, 16, 2, 4);
  run<0>("f16 one MMA only", 128, 2048, 32, 0, 1, 1);
  run<0>("f16 synthesis layout (K=16, N=64+32)", 128, 4096, 64, 32, 1, 2);
  run<1>("tf32 analysis layout (padded LBO)", 128, 2064, 16, 0, 4, 1);
  run<1>("tf32 analysis layout (padded LBO)", 128, 2064, 16, 0, 4, 4);
  return 0;
}
