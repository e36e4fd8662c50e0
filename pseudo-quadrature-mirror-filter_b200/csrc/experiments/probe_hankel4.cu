// Hankel-4: one A row = FOUR frames (64 samples = 128 bytes in fp16, the SWIZZLE_128B row pitch); the bank is replicated
// at 4 frame offsets along N, so one 128-row MMA produces 512 frames x 16 bands and every byte of the signal is read
// from shared memory 4x less often than with one frame per row.
//   D[i, (delta, k)] = sum_kappa A[i, kappa] * B[(delta, k), kappa],  A[i, kappa] = x[64 i + kappa],
//   B[(delta, k), kappa] = h[k, kappa - 16 delta]   ->   D[i, (delta, k)] = y[k, frame 4 i + delta]
// Checks exactness (small integers) and times the 27 + 27 MMA group (N = 128 / 64) issued by one elected lane.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ptx.cuh"
using namespace pqmf::ptx;

constexpr int KT = 384, KP = KT + 48, KS = KP / 16;   // taps, padded K, K-steps (27)
constexpr int ROWS = 128 + 7;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;  // 8 rows x 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) h4_kernel(const float* __restrict__ x, const float* __restrict__ h, float* __restrict__ D, long long* cyc, int reps) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* sx = sm;                   // ROWS x 128 B, swizzled
  unsigned char* sb = sm + 18432;           // [KP/8 chunks][128 rows][16 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < ROWS * 64; e += 128) {
    const uint32_t lin = e * 2;
    const uint32_t phys = lin ^ (((lin >> 7) & 7u) << 4);
    *reinterpret_cast<__half*>(sx + phys) = __float2half_rn(x[e]);
  }
  for (int e = tid; e < 128 * KP; e += 128) {
    const int row = e / KP, kap = e % KP;       // row = delta * 32 + k  (k < 32: 16 bands x {c1, c2} stand-in)
    const int delta = row / 32, k = row % 32, j = kap - 16 * delta;
    const float v = (j >= 0 && j < KT) ? h[k * KT + j] : 0.f;
    *reinterpret_cast<__half*>(sb + (kap / 8) * 2048 + row * 16 + (kap % 8) * 2) = __float2half_rn(v);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    long long tot = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (elect_one_sync()) {
        const uint64_t da = desc_sw128(smem_u32(sx)), db = umma_desc(smem_u32(sb), 2048, 128);
#pragma unroll
        for (int s = 0; s < KS; ++s)
          umma_f16(tm, da + (uint64_t)(8 * (s >> 2) + 2 * (s & 3)), db + (uint64_t)(256 * s), umma_idesc_f16(128, 128), s != 0);
        if (reps > 1) {
#pragma unroll
          for (int s = 0; s < KS; ++s)
            umma_f16(tm, da + (uint64_t)(8 * (s >> 2) + 2 * (s & 3)), db + (uint64_t)(256 * s), umma_idesc_f16(128, 64), true);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, r & 1);
      if (r > 0) tot += clock64() - t0;
    }
    if (tid == 0 && reps > 1) cyc[blockIdx.x] = tot / (reps - 1);
  } else {
    for (int r = 0; r < reps; ++r) mbar_wait(&bar, r & 1);
  }
  tc_fence_after();
  if (reps == 1) {
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[tid * 128 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

__global__ void __launch_bounds__(128) h4_pipe_kernel(long long* cyc, int reps) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* sx = sm;
  unsigned char* sb = sm + 18432;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (18432 + (KP / 8) * 2048) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c002c00u;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one_sync()) {
        const uint64_t da = desc_sw128(smem_u32(sx)), db = umma_desc(smem_u32(sb), 2048, 128);
#pragma unroll
        for (int s = 0; s < KS; ++s)
          umma_f16(tm + 128 * (r & 1), da + (uint64_t)(2 * s), db + (uint64_t)(256 * s), umma_idesc_f16(128, 128), s != 0);
#pragma unroll
        for (int s = 0; s < KS; ++s)
          umma_f16(tm + 128 * (r & 1), da + (uint64_t)(2 * s), db + (uint64_t)(256 * s), umma_idesc_f16(128, 64), true);
        umma_commit(&bar[r & 1]);
      }
      __syncwarp();
      if (r > 0) mbar_wait(&bar[(r - 1) & 1], ((r - 1) >> 1) & 1);   // wait for the PREVIOUS group only
    }
    mbar_wait(&bar[(reps - 1) & 1], ((reps - 1) >> 1) & 1);
    if (tid == 0) cyc[blockIdx.x] = (clock64() - t0) / reps;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
  std::vector<float> x(ROWS * 64), h(32 * KT);
  srand(5);
  for (auto& v : x) v = (float)((rand() % 9) - 4) * 0.25f;
  for (auto& v : h) v = (float)((rand() % 7) - 3) * 0.5f;
  std::vector<double> ref(128 * 128, 0.0);
  for (int i = 0; i < 128; ++i) for (int d = 0; d < 4; ++d) for (int k = 0; k < 32; ++k) {
    double a = 0; for (int j = 0; j < KT; ++j) a += (double)x[64 * i + 16 * d + j] * h[k * KT + j];
    ref[i * 128 + d * 32 + k] = a;
  }
  float *dx, *dh, *dD; long long* dc;
  cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dh, h.size() * 4); cudaMalloc(&dD, 128 * 128 * 4); cudaMalloc(&dc, 148 * 8);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dh, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 18432 + (KP / 8) * 2048 + 1024;
  cudaFuncSetAttribute(h4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  h4_kernel<<<1, 128, smem>>>(dx, dh, dD, dc, 1);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(128 * 128);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double mx = 0;
  for (int i = 0; i < 128 * 128; ++i) { double er = fabs((double)D[i] - ref[i]); if (er != 0) ++bad; if (er > mx) mx = er; }
  printf("hankel-4 exactness (SW128, 27 K-steps, N=128): cuda=%s mismatches=%d max|err|=%g -> %s\n", cudaGetErrorString(e), bad, mx, bad ? "FAIL" : "PASS");
  h4_kernel<<<148, 128, smem>>>(dx, dh, dD, dc, 200);
  e = cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("27 x MMA(128x128x16) + 27 x MMA(128x64x16) per 8192 samples: %lld cycles issue->complete (1 CTA/SM, all SMs) [%s]\n", c, cudaGetErrorString(e));
  printf("  -> %.2f samples/clk/SM; tensor floor would be 27*(64+32) = 2592 cycles = 3.16 samples/clk\n", 8192.0 / c);
  cudaFuncSetAttribute(h4_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  h4_pipe_kernel<<<148, 128, smem>>>(dc, 200);
  e = cudaDeviceSynchronize();
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("pipelined (next group issued before the previous one is waited for): %lld cycles per 8192-sample group -> %.2f samples/clk/SM [%s]\n", c, 8192.0 / c, cudaGetErrorString(e));
  return 0;
}
