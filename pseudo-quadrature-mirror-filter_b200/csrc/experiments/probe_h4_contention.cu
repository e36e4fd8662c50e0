// What slows the Hankel-4 MMA group down inside the real kernel?  One warp issues the pipelined 27 + 27 MMA groups of
// probe_hankel4.cu while the other seven warps of the CTA generate one kind of background traffic:
//   mode 1: STS.64 stream into a scratch plane (the fp16 conversion's writes)
//   mode 2: tcgen05.ld of the OTHER accumulator buffer (the epilogue's reads)
//   mode 4: LDG.128 stream from global memory (the window prefetch)
//   mode 8: STG.128 stream to global memory (the epilogue's stores)
// Prints cycles per MMA group for each mode and for all of them together.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../ptx.cuh"
using namespace pqmf::ptx;

constexpr int KS = 27;
constexpr int OFF_BANK = 18432, OFF_SCRATCH = OFF_BANK + KS * 4096, SCRATCH = 32768, SMEM = OFF_SCRATCH + SCRATCH + 1024;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(256, 1) contention_kernel(const float4* __restrict__ gin, float4* __restrict__ gout, size_t per_cta, long long* cyc, long long* work,
                                                             int reps, int mode, int n128 = KS, int n64 = KS) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (OFF_SCRATCH + SCRATCH) / 4; i += 256) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c002c00u;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); done = 0; }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode & 16) {   // no tensor work: idle for the time one group takes
        const long long t1 = clock64();
        while (clock64() - t1 < 3240) {}
        continue;
      }
      if (elect_one_sync()) {
        const uint64_t da = desc_sw128(smem_u32(sm)), db = umma_desc(smem_u32(sm + OFF_BANK), 2048, 128);
        for (int s = 0; s < n128; ++s) umma_f16(tm + 128 * (r & 1), da + (uint64_t)(2 * s), db + (uint64_t)(256 * s), umma_idesc_f16(128, 128), s != 0);
        for (int s = 0; s < n64; ++s) umma_f16(tm + 128 * (r & 1), da + (uint64_t)(2 * s), db + (uint64_t)(256 * s), umma_idesc_f16(128, 64), true);
        umma_commit(&bar[r & 1]);
      }
      __syncwarp();
      if (r > 0) mbar_wait(&bar[(r - 1) & 1], ((r - 1) >> 1) & 1);
    }
    if (!(mode & 16)) mbar_wait(&bar[(reps - 1) & 1], ((reps - 1) >> 1) & 1);
    if (tid == 0) { cyc[blockIdx.x] = (clock64() - t0) / reps; done = 1; }
  } else {
    // background traffic until the MMA warp is finished
    const int bt = tid - 32;  // 0..223
    const float4* gi = gin + (size_t)blockIdx.x * per_cta;
    float4* go = gout + (size_t)blockIdx.x * per_cta;
    float4 acc = make_float4(0, 0, 0, 0);
    long long n = 0;
    size_t pos = bt;
    while (!done) {
      if (mode == 0) __nanosleep(500);
      if (mode & 1) {
#pragma unroll
        for (int u = 0; u < 8; ++u) *reinterpret_cast<uint2*>(sm + OFF_SCRATCH + ((bt * 8 + u * 224 * 8) & (SCRATCH - 1))) = make_uint2(n, u);
      }
      if (mode & 2) {
        uint32_t r0[8], r1[8];
        const uint32_t ta = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)((n & 7) * 8);
        tmem_ld8(ta, r0);
        tmem_ld8(ta + 64, r1);
        tmem_ld_wait();
        acc.x += __uint_as_float(r0[0]) + __uint_as_float(r1[3]);
      }
      if (mode & 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 v = __ldcs(gi + ((pos + u * 224) & (per_cta - 1)));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      if (mode & 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) __stcs(go + ((pos + u * 224) & (per_cta - 1)), make_float4(n, u, 0, 0));
      }
      pos += 4 * 224;
      ++n;
    }
    if (acc.x == 1234.5f) go[0] = acc;
    if (bt == 0) work[blockIdx.x] = n;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  const size_t per_cta = (size_t)1 << 18;  // float4 per CTA = 4 MiB -> 592 MiB per buffer, far beyond L2
  float4 *gin, *gout; long long *dc, *dw;
  cudaMalloc(&gin, per_cta * 148 * 16); cudaMalloc(&gout, per_cta * 148 * 16); cudaMalloc(&dc, 148 * 8); cudaMalloc(&dw, 148 * 8);
  cudaMemset(gin, 0, per_cta * 148 * 16);
  cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int modes[] = {0, 1, 17, 2, 18, 4, 20, 8, 24, 15, 31};
  for (int m : modes) {
    contention_kernel<<<148, 256, SMEM>>>(gin, gout, per_cta, dc, dw, 100, m);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0, w = 0; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&w, dw, 8, cudaMemcpyDeviceToHost);
    const double per_group = (double)w / 100.0;
    printf("mode %2d%s: %5lld cycles per MMA group | background iterations per group %.1f", m & 15, (m & 16) ? " (tensor pipe idle)" : "                   ", c, per_group);
    if (m & 1) printf(" | STS %.1f KB", per_group * 8 * 224 * 8 / 1024);
    if (m & 2) printf(" | tmem_ld pairs/thread %.1f", per_group);
    if (m & 4) printf(" | LDG %.1f KB", per_group * 4 * 224 * 16 / 1024);
    if (m & 8) printf(" | STG %.1f KB", per_group * 4 * 224 * 16 / 1024);
    printf(" [%s]\n", cudaGetErrorString(e));
  }
  const int shapes[][2] = {{27, 0}, {0, 27}, {27, 27}, {19, 27}, {23, 27}, {1, 0}, {0, 1}};
  for (auto& sh : shapes) {
    contention_kernel<<<148, 256, SMEM>>>(gin, gout, per_cta, dc, dw, 100, 0, sh[0], sh[1]);
    cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("%2d x N=128 + %2d x N=64 per group: %lld cycles\n", sh[0], sh[1], c);
  }
  return 0;
}
