// Fast synthesis for n_band = 16, L = 512 (transpose of fast16.cuh's analysis):
//   w[n, r]  = sum_k C[k, r] * sigma(k, n) * s[k, n]                       inverse modulation, UMMA [256 x 16] x [16 x 32], 3xTF32
//   out[tau] = sum_{16 n' + r + 32 q - off2' = tau} 16 g[r + 32 q] * w[n, r]   overlap-add polyphase FIR, packed FFMA2
// replacing reference pqmf.py:133-157 (+ reverse_half :13-22, the x M gain :152, the flip/rearrange/crop :154-156) and the
// cached variant :345-354 in ONE pass: load sub-bands -> sign + tf32 split -> tensor-core modulation -> TMEM -> shared W tile
// -> FIR -> interleaved coalesced store.  The x16 gain is folded into the taps (exact: power of two).
//
// Frame pairing: output sample u = tau + off2' = 32 i + phi receives
//     sum_q G[phi + 32 q] * w[n' = 2 (i - q), phi]  +  sum_q G[ro + 32 q] * w[n' = 2 (i - q - d) + 1, ro],   ro = (phi + 16) mod 32,
// d = 1 for phi < 16.  n' = n + shift where shift in {0, 1} makes off2' = off2 + 16 shift a multiple of 32 (off2 = 256 for
// PQMF.inverse, 240 for CachedPQMF.inverse, -16 for streaming), so all three variants share one kernel.
//
// A tile = 256 sub-band frames (two M = 128 UMMAs) = 128 frame pairs -> NI = 8 J outputs rows i (32 samples each);
// J = 14 offline (124 of 128 pairs used: 12 pairs of halo are recomputed per tile), J = 8 for short streaming blocks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fast16.cuh"
#include "ptx.cuh"

namespace pqmf {

constexpr int kF16SynRows = 256;                         // sub-band frames per tile
constexpr int kF16SynLboA = kF16SynRows * 16 + 16;       // bytes between K-chunks of an A plane [256 x 16]
constexpr int kF16SynWStride = 144;                      // bytes per W row (32 floats + 16 pad: conflict-free STS.128 / LDS.64)

struct F16SynthesisParams {
  const float* s;        // [B, 16, F]
  const float* hist;     // [B, 16, 32] or nullptr
  float* out;            // [B, 16 F]
  float* hist_out;       // [B, 16, 32] or nullptr
  const float* tables;   // [ g (512) | C_hi (16*32) | C_lo (16*32) ]
  long F;
  int B;
  int shift;             // n' = n + shift
  int i_off;             // off2' / 32
  int parity;            // global parity of frame 0
  long tiles_per_row;
  long n_tiles;
};

struct F16SynthesisSmem {
  static constexpr int APLANE = 4 * kF16SynLboA;          // one tf32 plane of A = s^T [256 x 16]
  static constexpr int BPLANE = 4 * 512;                  // one plane of B = C^T [32 x 16]: 4 K-chunks x (32 rows x 16 B)
  static constexpr int WBYTES = kF16SynRows * kF16SynWStride;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + 2 * APLANE;
  static constexpr int OFF_W = OFF_B + 2 * BPLANE;
  static constexpr int OFF_BAR = OFF_W + WBYTES;
  static constexpr int BYTES = OFF_BAR + 64;
};

template <int QLO, int QN, int J>
__global__ void __launch_bounds__(kF16Threads, 3) f16_synthesis_kernel(F16SynthesisParams p) {
  using S = F16SynthesisSmem;
  constexpr int NI = 8 * J;                 // output rows (of 32 samples) per tile
  constexpr int PAIRS = NI + QN;            // frame pairs a tile needs
  constexpr int ROWS_USED = 2 * PAIRS;      // <= 256
  static_assert(ROWS_USED <= kF16SynRows, "tile does not fit two UMMAs");
  extern __shared__ __align__(128) unsigned char f16s_smem[];
  unsigned char* smem = f16s_smem;
  unsigned char* aplane = smem + S::OFF_A;
  unsigned char* bplane = smem + S::OFF_B;
  unsigned char* wtile = smem + S::OFF_W;
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pp = lane & 15;
  const int ig = warp * 2 + (lane >> 4);
  const int phi = 2 * pp;
  const int ro = (phi + 16) & 31;

  if (tid == 0) {
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  // B = C^T: rows r (N = 32), K = band k.  tables hold C[k][r].
  for (int e = tid; e < 2 * kF16M * kF16R; e += kF16Threads) {
    const int pl = e / (kF16M * kF16R), rem = e % (kF16M * kF16R);
    const int k = rem / kF16R, r = rem % kF16R;
    *reinterpret_cast<float*>(bplane + pl * S::BPLANE + (k >> 2) * 512 + r * 16 + (k & 3) * 4) = __ldg(p.tables + kF16L + e);
  }
  float2 ge[QN], go[QN + 1];
  {
    const float* g = p.tables;
    const float gain = (float)kF16M;
#pragma unroll
    for (int q = 0; q < QN; ++q)
      ge[q] = make_float2(gain * __ldg(g + phi + 32 * (QLO + q)), gain * __ldg(g + phi + 1 + 32 * (QLO + q)));
    const int d = (pp < 8) ? 1 : 0;
#pragma unroll
    for (int t = 0; t <= QN; ++t) {
      const int q = t - d;
      go[t] = (q >= 0 && q < QN) ? make_float2(gain * __ldg(g + ro + 32 * (QLO + q)), gain * __ldg(g + ro + 1 + 32 * (QLO + q)))
                                 : make_float2(0.f, 0.f);
    }
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = ptx::umma_idesc_tf32(128, 32);

  const unsigned tpr = (unsigned)p.tiles_per_row;
  const unsigned step_b = gridDim.x / tpr, step_c = gridDim.x % tpr;
  unsigned b = blockIdx.x / tpr, c = blockIdx.x % tpr;

  long it = 0;
  for (long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
    const long i0 = (long)p.i_off + (long)NI * c;   // first output row of the tile
    const long mlo = i0 - QLO - QN;                 // first frame pair of the tile
    const float* sb = p.s + (size_t)b * kF16M * p.F;
    const float* hb = p.hist ? p.hist + (size_t)b * kF16M * 32 : nullptr;

    // ---------------- load 2 rows x 16 bands, sign, tf32 split, store as UMMA A operand ----------------
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = tid + 128 * half;
      float v[16];
      const long n = 2 * mlo + row - p.shift;  // true frame index
      if (row < ROWS_USED && n >= 0 && n < p.F) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __ldg(sb + (size_t)k * p.F + n);
      } else if (row < ROWS_USED && hb != nullptr && n < 0 && n >= -32) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __ldg(hb + k * 32 + (32 + n));
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0.f;
      }
      const uint32_t flip = (((n + p.parity) & 1) == 0) ? 0x80000000u : 0u;
      if (ROWS_USED > 128 || half == 0) {
        unsigned char* arow = aplane + row * 16;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          float4 hi, lo;
          float* hp = &hi.x;
          float* lp = &lo.x;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * c4 + e;
            const float val = __uint_as_float(__float_as_uint(v[k]) ^ ((k & 1) ? flip : 0u));
            const float h = __uint_as_float((__float_as_uint(val) + 0x1000u) & 0xffffe000u);
            hp[e] = h;
            lp[e] = val - h;
          }
          *reinterpret_cast<float4*>(arow + c4 * kF16SynLboA) = hi;
          *reinterpret_cast<float4*>(arow + c4 * kF16SynLboA + S::APLANE) = lo;
        }
      }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();  // also: every thread is done with the previous tile's W (FIR) and TMEM reads
    // ---------------- inverse modulation: D_h[128 x 32] = A_hi B_hi + A_lo B_hi + A_hi B_lo, h = 0, 1 ----------------
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t a_hi = ptx::smem_u32(aplane), a_lo = a_hi + S::APLANE;
      const uint32_t b_hi = ptx::smem_u32(bplane), b_lo = b_hi + S::BPLANE;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 1 && ROWS_USED <= 128) break;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a = ((term == 1) ? a_lo : a_hi) + half * 128 * 16;
          const uint32_t bb = (term == 2) ? b_lo : b_hi;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            ptx::umma_tf32(tmem + half * 32, ptx::umma_desc(a + ks * 2 * kF16SynLboA, kF16SynLboA, kF16Sbo),
                           ptx::umma_desc(bb + ks * 2 * 512, 512, kF16Sbo), idesc, (term | ks) != 0);
        }
      }
      ptx::umma_commit(mma_bar);
    }
    // streaming: the CTA that owns the last tile of a row rolls that row's sub-band history while the MMAs run
    if (p.hist_out != nullptr && c + 1 == tpr) {
      const int k = tid >> 3, j0 = (tid & 7) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long cc = p.F - 32 + j0 + e;
        p.hist_out[(size_t)b * 512 + k * 32 + j0 + e] = (cc >= 0) ? sb[(size_t)k * p.F + cc] : hb[k * 32 + 32 + cc];
      }
    }
    ptx::mbar_wait(mma_bar, (uint32_t)(it & 1));
    ptx::tc_fence_after();
    // ---------------- TMEM -> registers -> W tile [256 rows x 32] in shared memory ----------------
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (half == 1 && ROWS_USED <= 128) break;
      if (128 * half + 32 * warp < ROWS_USED) {  // warp-uniform: skip row groups the FIR never reads
        uint32_t r[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32), r);
        ptx::tmem_ld_wait();
        unsigned char* wrow = wtile + (tid + 128 * half) * kF16SynWStride;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          *reinterpret_cast<uint4*>(wrow + c4 * 16) = make_uint4(r[4 * c4], r[4 * c4 + 1], r[4 * c4 + 2], r[4 * c4 + 3]);
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    // ---------------- overlap-add FIR: J outputs x 2 phases per thread ----------------
    float2 acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = make_float2(0.f, 0.f);
    {
      const unsigned char* we_p = wtile + (2 * (J * ig)) * kF16SynWStride + phi * 4;
      const unsigned char* wo_p = wtile + (2 * (J * ig) + 1) * kF16SynWStride + ro * 4;
#pragma unroll
      for (int e = 0; e < J + QN; ++e) {
        const float2 we = *reinterpret_cast<const float2*>(we_p + e * 2 * kF16SynWStride);
        const float2 wo = *reinterpret_cast<const float2*>(wo_p + e * 2 * kF16SynWStride);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int t = j + QN - e;
          if (t >= 0 && t < QN) acc[j] = ptx::ffma2(ge[t], we, acc[j]);
          if (t >= 0 && t <= QN) acc[j] = ptx::ffma2(go[t], wo, acc[j]);
        }
      }
    }
    {
      const long total = 16 * p.F;
      float* ob = p.out + (size_t)b * total;
      const long tau0 = 32 * (i0 + (long)J * ig - p.i_off) + phi;  // tau = 32 i + phi - off2'
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const long tau = tau0 + 32 * j;
        if (tau >= 0 && tau < total) __stcs(reinterpret_cast<float2*>(ob + tau), acc[j]);
      }
    }
    b += step_b;
    c += step_c;
    if (c >= tpr) {
      c -= tpr;
      ++b;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

inline bool fast16_synthesis_ok(const float* s, const float* out, long F) {
  return F > 0 && ((uintptr_t)s % 4) == 0 && ((uintptr_t)out % 8) == 0;
}

template <int QLO, int QN, int J>
int f16_launch_synthesis(F16SynthesisParams p, cudaStream_t st) {
  using S = F16SynthesisSmem;
  constexpr int NI = 8 * J;
  auto kern = f16_synthesis_kernel<QLO, QN, J>;
  static int sm_count[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sm_count[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
    if (e != cudaSuccess) return (int)e;
    sm_count[dev] = n > 0 ? n : 148;
  }
  const long rows_i = (16 * p.F + 31) / 32;  // output rows of 32 samples per batch row
  p.tiles_per_row = (rows_i + NI - 1) / NI;
  p.n_tiles = p.tiles_per_row * p.B;
  if (p.tiles_per_row >= (1L << 31) || p.n_tiles >= (1L << 40)) return -2;
  long grid = (long)sm_count[dev] * 3;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, kF16Threads, S::BYTES, st>>>(p);
  return (int)cudaGetLastError();
}

// s [B,16,F] (+ hist [B,16,32]) -> out [B,16F] (+ hist_out).  off2 = 256 (PQMF), 240 (CachedPQMF offline), -16 (streaming).
inline int fast16_synthesis(const float* s, const float* hist, float* out, float* hist_out, const float* tables, int B, long F, int off2,
                            int parity, unsigned flags, cudaStream_t st) {
  F16SynthesisParams p{};
  p.s = s; p.hist = hist; p.out = out; p.hist_out = hist_out; p.tables = tables;
  p.F = F; p.B = B; p.parity = parity & 1;
  p.shift = ((off2 % 32) + 32) % 32 == 16 ? 1 : 0;
  const int off2p = off2 + 16 * p.shift;
  if (off2p % 32 != 0 || off2p < 0) return -2;
  p.i_off = off2p / 32;
  const F16Taps t = fast16_taps_from_flags(flags);
  const bool short_rows = F <= 160;  // streaming-sized blocks: one J = 8 tile covers 64 output rows + history
  if (t.qn == 12) return short_rows ? f16_launch_synthesis<2, 12, 8>(p, st) : f16_launch_synthesis<2, 12, 14>(p, st);
  return short_rows ? f16_launch_synthesis<0, 16, 8>(p, st) : f16_launch_synthesis<0, 16, 14>(p, st);
}

}  // namespace pqmf
