// MMA issue-order probe (round 2), derived from probe_h4_pair.cu: does interleaving the math-bound N = 128 MMAs with the
// operand-fetch-bound N = 64 MMAs let the tensor pipe overlap them?  And does the N = 64 pass cost the same into columns 64-127?
// Hankel-4 on a CTA PAIR (tcgen05 cta_group::2): M = 256 (each CTA's own 128 signal rows), the bank split across the two
// CTAs (each holds N/2 rows), so every SM reads half of B per MMA.  Checks exactness against a host reference for the
// N = 128 and the N = 64 pass, then times pipelined 27 + 27 MMA groups (one group = two 8192-sample tiles).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ptx.cuh"
using namespace pqmf::ptx;

constexpr int KT = 384, KP = KT + 48, KS = KP / 16;
constexpr int ROWS = 128 + 7;
constexpr int OFF_B128 = 18432;                    // this CTA's 64 rows of the N = 128 bank: [KP/8][64][16 B]
constexpr int OFF_B64 = OFF_B128 + (KP / 8) * 1024; // this CTA's 32 rows of the N = 64 bank:  [KP/8][32][16 B]
constexpr int SMEM = OFF_B64 + (KP / 8) * 512 + 1024;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cluster_ctarank() / cluster_sync_all() now live in ../ptx.cuh
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// x: 2 tiles of 8192 samples (+ tail); h: [64 bank rows][KT]: rows 0-31 stand in for c1 (delta-major), 32-63 for c2
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_kernel(const float* __restrict__ x, const float* __restrict__ h, float* __restrict__ D,
                                                                             long long* cyc, int reps, int n128 = KS, int n64 = KS, int order = 0) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const float* xt = x + 8192 * rank;
  for (int e = tid; e < ROWS * 64; e += 128) {
    const uint32_t lin = e * 2;
    const uint32_t phys = lin ^ (((lin >> 7) & 7u) << 4);
    *reinterpret_cast<__half*>(sm + phys) = __float2half_rn(xt[e]);
  }
  // bank row n of the full N = 128 operand: delta = (n % 64) / 16, band = n % 16, part = n / 64 ; h row = part * 32 + ... keep it
  // simple: full row n uses h[(n % 64)] shifted by 16 * delta' where delta' = (n % 64) / 16, scaled by (1 + part)
  auto bank = [&](int n, int kap) {
    const int part = n / 64, dl = (n % 64) / 16, k = n % 16, j = kap - 16 * dl;
    const float v = (j >= 0 && j < KT) ? h[(part * 16 + k) * KT + j] : 0.f;
    return v;
  };
  for (int e = tid; e < 64 * KP; e += 128) {   // my 64 rows of the N = 128 operand
    const int row = e / KP, kap = e % KP;
    *reinterpret_cast<__half*>(sm + OFF_B128 + (kap / 8) * 1024 + row * 16 + (kap % 8) * 2) = __float2half_rn(bank(64 * rank + row, kap));
  }
  for (int e = tid; e < 32 * KP; e += 128) {   // my 32 rows of the N = 64 operand (= part 0 rows 32 rank .. 32 rank + 31)
    const int row = e / KP, kap = e % KP;
    *reinterpret_cast<__half*>(sm + OFF_B64 + (kap / 8) * 512 + row * 16 + (kap % 8) * 2) = __float2half_rn(bank(32 * rank + row, kap));
  }
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc2(&slot, 256);
  fence_proxy_async(); tc_fence_before(); __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  const uint32_t idesc64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  const long long t0 = clock64();
  if (warp == 0) {
    for (int r = 0; r < reps; ++r) {
      if (rank == 0) {
        if (elect_one_sync()) {
          const uint64_t da = desc_sw128(smem_u32(sm));
          const uint64_t db128 = umma_desc(smem_u32(sm + OFF_B128), 1024, 128), db64 = umma_desc(smem_u32(sm + OFF_B64), 512, 128);
          const uint32_t d0 = tm + 128 * (r & 1), d1 = d0 + ((order & 2) ? 64u : 0u);   // order bit 1: the N = 64 pass accumulates into columns 64-127
          auto astep = [&](int s) { return da + (uint64_t)(8 * (s >> 2) + 2 * (s & 3)); };
          if (order & 1) {   // interleaved, branch-free: N = 128 (step s), N = 64 (step s), ... over the common count; the rest after
            const int nmin = n128 < n64 ? n128 : n64;
            for (int s = 0; s < nmin; ++s) {
              umma2_f16(d0, astep(s), db128 + (uint64_t)(128 * s), idesc128, s != 0);
              umma2_f16(d1, astep(s), db64 + (uint64_t)(64 * s), idesc64, true);
            }
            for (int s = nmin; s < n128; ++s) umma2_f16(d0, astep(s), db128 + (uint64_t)(128 * s), idesc128, s != 0);
            for (int s = nmin; s < n64; ++s) umma2_f16(d1, astep(s), db64 + (uint64_t)(64 * s), idesc64, true);
          } else {
            for (int s = 0; s < n128; ++s) umma2_f16(d0, astep(s), db128 + (uint64_t)(128 * s), idesc128, s != 0);
            for (int s = 0; s < n64; ++s) umma2_f16(d1, astep(s), db64 + (uint64_t)(64 * s), idesc64, true);
          }
          umma2_commit_mc(&bar[r & 1], 3);
        }
        __syncwarp();
      }
      if (r > 0) mbar_wait(&bar[(r - 1) & 1], ((r - 1) >> 1) & 1);
    }
  }
  mbar_wait(&bar[(reps - 1) & 1], ((reps - 1) >> 1) & 1);
  if (tid == 0) cyc[blockIdx.x] = (clock64() - t0) / reps;
  tc_fence_after();
  if (reps == 1) {
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[((size_t)blockIdx.x * 128 + tid) * 128 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before(); __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tm, 256);
}

int main() {
  std::vector<float> x(2 * 8192 + 512), h(32 * KT);
  srand(5);
  for (auto& v : x) v = (float)((rand() % 9) - 4) * 0.25f;
  for (auto& v : h) v = (float)((rand() % 7) - 3) * 0.5f;
  // reference: D_r[i, n] = sum_kap x[8192 r + 64 i + kap] * bank(n, kap) for n < 128 (pass 1) plus, for n < 64, the same again (pass 2)
  auto bank = [&](int n, int kap) {
    const int part = n / 64, dl = (n % 64) / 16, k = n % 16, j = kap - 16 * dl;
    return (j >= 0 && j < KT) ? h[(part * 16 + k) * KT + j] : 0.f;
  };
  std::vector<double> ref(2 * 128 * 128, 0.0);
  for (int r = 0; r < 2; ++r) for (int i = 0; i < 128; ++i) for (int n = 0; n < 128; ++n) {
    double a = 0; for (int kap = 0; kap < KP; ++kap) a += (double)x[8192 * r + 64 * i + kap] * bank(n, kap);
    ref[(r * 128 + i) * 128 + n] = n < 64 ? 2 * a : a;
  }
  float *dx, *dh, *dD; long long* dc;
  cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dh, h.size() * 4); cudaMalloc(&dD, 148 * 128 * 128 * 4); cudaMalloc(&dc, 148 * 8);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dh, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  pair_kernel<<<2, 128, SMEM>>>(dx, dh, dD, dc, 1);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(2 * 128 * 128);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double mx = 0;
  for (int i = 0; i < 2 * 128 * 128; ++i) { double er = fabs((double)D[i] - ref[i]); if (er != 0) { if (bad < 5) printf("  mismatch cta %d row %d col %d: got %g want %g\n", i / 16384, (i / 128) % 128, i % 128, D[i], ref[i]); ++bad; } if (er > mx) mx = er; }
  printf("pair exactness (cta_group::2, M=256, bank split across the pair): cuda=%s mismatches=%d max|err|=%g -> %s\n", cudaGetErrorString(e), bad, mx, bad ? "FAIL" : "PASS");
  if (e != cudaSuccess) return 1;
  pair_kernel<<<148, 128, SMEM>>>(dx, dh, dD, dc, 200);
  e = cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("pipelined pair groups (27 x N=128 + 27 x N=64; each SM of the pair does one tile per group): %lld cycles per group [%s]\n", c, cudaGetErrorString(e));
  printf("  (single-CTA Hankel-4: ~3240 cycles per tile)\n");
  const int shapes[][2] = {{27, 27}, {19, 27}, {23, 27}, {19, 19}, {15, 27}, {27, 0}, {0, 27}};
  for (int order = 0; order < 4; ++order)
    for (auto& sh : shapes) {
      pair_kernel<<<148, 128, SMEM>>>(dx, dh, dD, dc, 200, sh[0], sh[1], order);
      cudaDeviceSynchronize();
      cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      printf("order %d (bit0 interleave, bit1 N64 pass into cols 64-127): %2d x N=128 + %2d x N=64 per group: %lld cycles\n", order, sh[0], sh[1], c);
    }
  return 0;
}
