// Can tcgen05.mma read a HANKEL operand straight out of a contiguous fp16 signal?
//   A[n, j] = x[16 n + j]   (frame hop 16 samples = 32 bytes in fp16)
// With the K-major SWIZZLE_32B canonical layout the rows of a core matrix are exactly 32 bytes apart, so the A tile of
// K-step s (taps 16 s .. 16 s + 15) is "the signal viewed as 32-byte rows, starting at row n0 + s": only the descriptor's
// start address moves.  The signal is stored with the swizzle applied to its absolute shared-memory address
// (16-byte chunk index ^= address bit 7).  This probe checks D = A * B^T exactly (small-integer data) for K = 16 * KS.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ptx.cuh"
using namespace pqmf::ptx;

__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major) = 16 B
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;  // 8-row group pitch
  d |= (uint64_t)1 << 46;                       // sm_100 descriptor version
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)6 << 61;                       // SWIZZLE_32B
  return d;
}

template <int KS, int N>
__global__ void __launch_bounds__(128) hankel_kernel(const float* __restrict__ x, const float* __restrict__ B, float* __restrict__ D, int mode) {
  // x: signal of 16*(128+KS) samples ; B: [N][16*KS] ; D: [128][N]
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* sx = sm;                       // fp16 signal, swizzled, (128 + KS) rows x 32 B
  unsigned char* sb = sm + 8192;                // B: K-major no-swizzle, chunk stride N*16
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_samples = 16 * (128 + KS);
  for (int e = tid; e < n_samples; e += 128) {
    const uint32_t lin = e * 2;                                   // byte offset of sample e
    const uint32_t abs_addr = smem_u32(sx) + lin;
    const uint32_t phys = (mode & 1) ? (lin ^ (((abs_addr >> 7) & 1u) << 4)) : (lin ^ (((lin >> 7) & 1u) << 4));
    *reinterpret_cast<__half*>(sx + phys) = __float2half_rn(x[e]);
  }
  for (int e = tid; e < N * 16 * KS; e += 128) {
    const int row = e / (16 * KS), k = e % (16 * KS);
    *reinterpret_cast<__half*>(sb + (k / 8) * (N * 16) + row * 16 + (k % 8) * 2) = __float2half_rn(B[e]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    for (int s = 0; s < KS; ++s) {
      const uint32_t a_addr = smem_u32(sx) + 32 * s;
      const uint32_t boff = (mode & 2) ? ((a_addr >> 7) & 7u) : 0u;
      umma_f16(tm, desc_sw32(a_addr, 256, boff), umma_desc(smem_u32(sb) + s * 2 * (N * 16), N * 16, 128), umma_idesc_f16(128, N), s != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 16 && j < N; ++j) D[tid * N + j] = __uint_as_float(r[j]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 32);
}

int main() {
  constexpr int KS = 24, N = 16;
  const int n_samples = 16 * (128 + KS);
  std::vector<float> x(n_samples), B(N * 16 * KS);
  srand(3);
  for (auto& v : x) v = (float)((rand() % 9) - 4) * 0.25f;
  for (auto& v : B) v = (float)((rand() % 7) - 3) * 0.5f;
  std::vector<double> ref(128 * N, 0.0);
  for (int n = 0; n < 128; ++n) for (int k = 0; k < N; ++k) for (int j = 0; j < 16 * KS; ++j) ref[n * N + k] += (double)x[16 * n + j] * B[k * 16 * KS + j];
  float *dx, *dB, *dD;
  cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  auto kern = hankel_kernel<KS, N>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + N * 32 * KS + 2048);
  int fails = 0;
  for (int mode = 0; mode < 4; ++mode) {
    cudaMemset(dD, 0xff, 128 * N * 4);
    kern<<<1, 128, 8192 + N * 32 * KS + 2048>>>(dx, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double mx = 0;
    for (int i = 0; i < 128 * N; ++i) { double er = fabs((double)D[i] - ref[i]); if (!(er == 0)) ++bad; if (er > mx) mx = er; }
    printf("mode %d (store swizzle by %s address, base_offset %s): cuda=%s mismatches=%d max|err|=%g -> %s\n", mode,
           (mode & 1) ? "absolute" : "plane-relative", (mode & 2) ? "(addr>>7)&7" : "0", cudaGetErrorString(e), bad, mx, bad ? "FAIL" : "PASS");
    if (bad && mode == 3) fails = 1;
  }
  return 0;
}
