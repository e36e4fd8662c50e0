// Fast synthesis for n_band = 16, L = 512 (transpose of fast16.cuh's analysis):
//   w[n, r]  = sum_k C[k, r] * sigma(k, n) * s[k, n]                       inverse modulation, UMMA [256 x 16] x [16 x 32]
//   out[tau] = sum_{16 n' + r + 32 q - off2' = tau} 16 g[r + 32 q] * w[n, r]   overlap-add polyphase FIR, packed FFMA2
// replacing reference pqmf.py:133-157 (+ reverse_half :13-22, the x M gain :152, the flip/rearrange/crop :154-156) and the
// cached variant :345-354 in ONE pass: load sub-bands -> sign + fp16 split -> tensor-core modulation -> TMEM -> shared W tile
// -> FIR -> interleaved coalesced store.  The x16 gain is folded into the taps (exact: power of two).
//
// Modulation precision: same two-term fp16 scheme as the analysis (fast16.cuh): s' = h1 + 2^-11 h2, C = c1 + c2,
//   D[:, 0:64] = h1 [c1 | c2]^T ,  D[:, 0:32] += h2 (2^-11 c1)^T ,  w = D[:, 0:32] + D[:, 32:64].
//
// Frame pairing: output sample u = tau + off2' = 32 i + phi receives
//     sum_q G[phi + 32 q] * w[n' = 2 (i - q), phi]  +  sum_q G[ro + 32 q] * w[n' = 2 (i - q - d) + 1, ro],   ro = (phi + 16) mod 32,
// d = 1 for phi < 16.  n' = n + shift where shift in {0, 1} makes off2' = off2 + 16 shift a multiple of 32 (off2 = 256 for
// PQMF.inverse, 240 for CachedPQMF.inverse, -16 for streaming), so all three variants share one kernel.
//
// A tile = 256 sub-band frames (two M = 128 UMMAs) = 128 frame pairs -> NI = 16 J output rows i (32 samples each);
// J = 7 offline (124 of 128 pairs used: 12 pairs of halo are recomputed per tile), J = 4 for short streaming blocks.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fast16.cuh"
#include "ptx.cuh"

namespace pqmf {

constexpr int kF16SynThreads = 256;  // thread <-> sub-band frame row of the tile; FIR: 16 phase pairs x 16 output groups
constexpr int kF16SynRows = 256;     // sub-band frames per tile
constexpr int kF16SynLboA = kF16SynRows * 16;  // bytes between the two K-chunks (8 bands each) of an A plane [256 x 16] fp16
constexpr int kF16SynWStride = 144;            // bytes per W row (32 floats + 16 pad: conflict-free STS.128 / LDS.64)

struct F16SynthesisParams {
  const float* s;        // [B, 16, F]
  const float* hist;     // [B, 16, 32] or nullptr
  float* out;            // [B, 16 F]
  float* hist_out;       // [B, 16, 32] or nullptr
  const float* tables;   // [ g (512) | c1 (16*32) | c2 (16*32) ]
  long F;
  int B;
  int shift;             // n' = n + shift
  int i_off;             // off2' / 32
  int parity;            // global parity of frame 0
  long tiles_per_row;
  long n_tiles;
};

struct F16SynthesisSmem {
  static constexpr int APLANE = 2 * kF16SynLboA;          // one fp16 plane of A = s'^T [256 x 16]
  static constexpr int BCAT = 2 * 1024;                   // B = [c1 | c2]^T: N = 64 rows x K = 16 fp16, 2 chunks x (64 rows x 16 B)
  static constexpr int BRES = 2 * 512;                    // B = 2^-11 c1^T : N = 32 rows x K = 16
  static constexpr int WBYTES = kF16SynRows * kF16SynWStride;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + 2 * APLANE;
  static constexpr int OFF_W = OFF_B + BCAT + BRES;
  static constexpr int OFF_BAR = OFF_W + WBYTES;
  static constexpr int BYTES = OFF_BAR + 64;
};

// Per tile: registers (prefetched sub-bands) -> sign + fp16 split -> A planes -> [afull] -> one lane issues 4 UMMAs ->
// prefetch next tile's sub-bands into registers -> [mma_bar] -> TMEM -> W tile -> __syncthreads -> FIR -> store.
// The only block-wide barrier is the one in front of the FIR (every thread reads W rows written by others).
template <int QLO, int QN, int J>
__global__ void __launch_bounds__(kF16SynThreads, 2) f16_synthesis_kernel(F16SynthesisParams p) {
  using S = F16SynthesisSmem;
  constexpr int NI = 16 * J;                // output rows (of 32 samples) per tile
  constexpr int PAIRS = NI + QN;            // frame pairs a tile needs
  constexpr int ROWS_USED = 2 * PAIRS;      // <= 256
  static_assert(ROWS_USED <= kF16SynRows && ROWS_USED > 128, "tile must span both UMMAs");
  extern __shared__ __align__(128) unsigned char f16s_smem[];
  unsigned char* smem = f16s_smem;
  unsigned char* aplane = smem + S::OFF_A;
  unsigned char* bcat = smem + S::OFF_B;
  unsigned char* bres = bcat + S::BCAT;
  unsigned char* wtile = smem + S::OFF_W;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* mma_bar = afull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pp = lane & 15;
  const int ig = warp * 2 + (lane >> 4);  // output group 0..15
  const int phi = 2 * pp;
  const int ro = (phi + 16) & 31;

  if (tid == 0) {
    ptx::mbar_init(afull, kF16SynThreads);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  // B operands: rows = output column n of D, K = band k.  tables hold c1[k][r] then c2[k][r].
  for (int e = tid; e < 2 * kF16M * kF16R; e += kF16SynThreads) {
    const int pl = e / (kF16M * kF16R), rem = e % (kF16M * kF16R);
    const int k = rem / kF16R, r = rem % kF16R;
    const float c = __ldg(p.tables + kF16L + e);
    const int n = pl * 32 + r;  // row of [c1 | c2]^T
    *reinterpret_cast<uint16_t*>(bcat + (k >> 3) * 1024 + n * 16 + (k & 7) * 2) = f16_bits(c);
    if (pl == 0) *reinterpret_cast<uint16_t*>(bres + (k >> 3) * 512 + r * 16 + (k & 7) * 2) = f16_bits(c * (1.f / 2048.f));
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

  const unsigned tpr = (unsigned)p.tiles_per_row;
  const unsigned step_b = gridDim.x / tpr, step_c = gridDim.x % tpr;
  unsigned b = blockIdx.x / tpr, c = blockIdx.x % tpr;

  // this thread's sub-band frame (row tid of the tile): 16 bands; the sign mask is applied where the values are consumed
  // (not here: touching v[] right after the loads would stall on them and defeat the prefetch)
  float v[16];
  bool flip = false;
  auto load_row = [&](unsigned bb, unsigned cc) {
    const long mlo = (long)p.i_off + (long)NI * cc - QLO - QN;  // first frame pair of the tile
    const long n = 2 * mlo + tid - p.shift;                     // true frame index of row tid
    const float* sb = p.s + (size_t)bb * kF16M * p.F;
    if (tid < ROWS_USED && n >= 0 && n < p.F) {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldcs(sb + (size_t)k * p.F + n);
    } else if (tid < ROWS_USED && p.hist != nullptr && n < 0 && n >= -32) {
      const float* hb = p.hist + (size_t)bb * kF16M * 32;
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldg(hb + k * 32 + (32 + n));
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.f;
    }
    flip = ((n + p.parity) & 1) == 0;
  };
  if ((long)blockIdx.x < p.n_tiles) load_row(b, c);

  long it = 0;
  for (long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
    const long i0 = (long)p.i_off + (long)NI * c;   // first output row of the tile
    // ---------------- two-term fp16 split, stored as UMMA A operand (row tid, 2 K-chunks of 8 bands, 2 planes) ----------------
    {
      if (flip) {
#pragma unroll
        for (int k = 1; k < 16; k += 2) v[k] = -v[k];
      }
      unsigned char* arow = aplane + tid * 16;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint4 h1, h2;
        ptx::split_f16x2(make_float2(v[8 * ch + 0], v[8 * ch + 1]), h1.x, h2.x);
        ptx::split_f16x2(make_float2(v[8 * ch + 2], v[8 * ch + 3]), h1.y, h2.y);
        ptx::split_f16x2(make_float2(v[8 * ch + 4], v[8 * ch + 5]), h1.z, h2.z);
        ptx::split_f16x2(make_float2(v[8 * ch + 6], v[8 * ch + 7]), h1.w, h2.w);
        *reinterpret_cast<uint4*>(arow + ch * kF16SynLboA) = h1;
        *reinterpret_cast<uint4*>(arow + ch * kF16SynLboA + S::APLANE) = h2;
      }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    ptx::mbar_arrive(afull);
    // ---------------- inverse modulation, per 128-row half: D_h[128 x 64] = h1 [c1|c2]^T ; D_h[:, 0:32] += h2 (2^-11 c1)^T ----------------
    if (warp == 0) {
      ptx::mbar_wait(afull, (uint32_t)(it & 1));
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        const uint32_t a1 = ptx::smem_u32(aplane), a2 = a1 + S::APLANE;
        const uint64_t d_cat = ptx::umma_desc(ptx::smem_u32(bcat), 1024, 128), d_res = ptx::umma_desc(ptx::smem_u32(bres), 512, 128);
        constexpr uint32_t idesc64 = ptx::umma_idesc_f16(128, 64), idesc32 = ptx::umma_idesc_f16(128, 32);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          ptx::umma_f16(tmem + half * 64, ptx::umma_desc(a1 + half * 128 * 16, kF16SynLboA, 128), d_cat, idesc64, false);
          ptx::umma_f16(tmem + half * 64, ptx::umma_desc(a2 + half * 128 * 16, kF16SynLboA, 128), d_res, idesc32, true);
        }
        ptx::umma_commit(mma_bar);
      }
    }
    __syncwarp();
    // streaming: the CTA that owns the last tile of a row rolls that row's sub-band history while the MMAs run
    if (p.hist_out != nullptr && c + 1 == tpr) {
      const float* sb = p.s + (size_t)b * kF16M * p.F;
      const float* hb = p.hist + (size_t)b * kF16M * 32;
      for (int e = tid; e < 512; e += kF16SynThreads) {
        const int k = e >> 5;
        const long cc = p.F - 32 + (e & 31);
        p.hist_out[(size_t)b * 512 + e] = (cc >= 0) ? sb[(size_t)k * p.F + cc] : hb[k * 32 + 32 + cc];
      }
    }
    // ---------------- prefetch the next tile's sub-band frame into registers (consumed at the top of the next iteration) ----
    unsigned nb = b + step_b, nc = c + step_c;
    if (nc >= tpr) {
      nc -= tpr;
      ++nb;
    }
    if (tile + (long)gridDim.x < p.n_tiles) load_row(nb, nc);

    ptx::mbar_wait(mma_bar, (uint32_t)(it & 1));
    ptx::tc_fence_after();
    // ---------------- TMEM -> registers (main + correction columns) -> W tile [256 rows x 32] in shared memory ----------------
    if (32 * warp < ROWS_USED) {  // warp-uniform: skip row groups the FIR never reads
      const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
      unsigned char* wrow = wtile + tid * kF16SynWStride;
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t r0[16], r1[16];
        ptx::tmem_ld16(trow + part * 16, r0);
        ptx::tmem_ld16(trow + 32 + part * 16, r1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          float4 w;
          w.x = __uint_as_float(r0[4 * c4 + 0]) + __uint_as_float(r1[4 * c4 + 0]);
          w.y = __uint_as_float(r0[4 * c4 + 1]) + __uint_as_float(r1[4 * c4 + 1]);
          w.z = __uint_as_float(r0[4 * c4 + 2]) + __uint_as_float(r1[4 * c4 + 2]);
          w.w = __uint_as_float(r0[4 * c4 + 3]) + __uint_as_float(r1[4 * c4 + 3]);
          *reinterpret_cast<float4*>(wrow + part * 64 + c4 * 16) = w;
        }
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
    b = nb;
    c = nc;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 128);
}

inline bool fast16_synthesis_ok(const float* s, const float* out, long F) {
  return F > 0 && ((uintptr_t)s % 4) == 0 && ((uintptr_t)out % 8) == 0;
}

template <int QLO, int QN, int J>
int f16_launch_synthesis(F16SynthesisParams p, cudaStream_t st) {
  using S = F16SynthesisSmem;
  constexpr int NI = 16 * J;
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
  long grid = (long)sm_count[dev] * 2;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, kF16SynThreads, S::BYTES, st>>>(p);
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
  const bool short_rows = F <= 160;  // streaming-sized blocks: one J = 4 tile covers 64 output rows + history
  if (t.qn == 12) return short_rows ? f16_launch_synthesis<2, 12, 4>(p, st) : f16_launch_synthesis<2, 12, 7>(p, st);
  return short_rows ? f16_launch_synthesis<0, 16, 4>(p, st) : f16_launch_synthesis<0, 16, 7>(p, st);
}

}  // namespace pqmf
