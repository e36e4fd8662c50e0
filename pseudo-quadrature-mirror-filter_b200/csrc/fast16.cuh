// Fast path for n_band = 16, L = 512: window fold on CUDA cores (packed FFMA2) + cosine modulation on the
// 5th-gen tensor cores (tcgen05.mma kind::f16 on a two-term fp16 split, fp32 accumulators in TMEM), input tiles
// staged by the TMA engine (cp.async.bulk + mbarrier), one fused pass per direction.
//
// Factorisation (reference arithmetic it replaces: pqmf.py:115-130 + :13-22 and :133-157):
//   hk[k, r + 32 q] = g[r + 32 q] * C[k, r],   r in [0, 32), q in [0, 16)
//   analysis : v[n, r] = sum_q g[r + 32 q] * X[16 n + r + 32 q - off]          (fold, 32 FMA / sample, exact fp32)
//              y[k, n] = sigma(k, n) * sum_r C[k, r] * v[n, r]                 (modulation, [128 x 32] x [32 x 16] MMA)
//   synthesis: w[n, r] = sum_k C[k, r] * sigma(k, n) * s[k, n]                 (modulation, [128 x 16] x [16 x 32] MMA)
//              out[tau] = sum_{16 n + r + 32 q - off2 = tau} 16 g[r + 32 q] * w[n, r]   (overlap-add FIR)
// Even / odd frames use disjoint 32-sample phase grids, so the fold is 32 phase sequences z_phi[i] = X[32 i + phi],
// each run through two 16-tap FIRs.  Lanes own phase PAIRS (conflict-free LDS.64, natural float2 operands for
// FFMA2), registers hold the taps and J = 8 outputs per parity.
//
// Modulation precision: v = h1 + 2^-11 h2 and C = c1 + c2 with all four terms fp16; the MMAs compute
//   h1 [c1 | c2]  (one N = 32 instruction stream: columns 0-15 main term, 16-31 the c2 correction)  +  h2 (2^-11 c1)
// with exact fp16 x fp16 products accumulated in fp32, dropping only h2 c2 ~ 2^-23 |v C|: fp32-level accuracy for
// 12 bytes/sample of shared-memory operand traffic (a 3xTF32 split costs 24; this kernel is shared-memory-bandwidth
// bound).
//
// Taps: the prototype (N = 377 at attenuation 100) is centre-padded to 512, so g is identically zero for
// q in {0, 1, 14, 15}; the <QLO = 2, QN = 12> instantiation skips them (25 % fewer FMAs, smaller halo).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace pqmf {

constexpr int kF16Threads = 128;
constexpr int kF16M = 16, kF16L = 512, kF16R = 32;  // bands, bank length, fold width 2M
constexpr int kF16TileFrames = 128;                 // frames per tile = rows of one UMMA
constexpr int kF16J = 8;                            // fold outputs per thread per frame parity
// A operand planes [128 rows x 32 K] fp16, K-major no-swizzle: 4 K-chunks of 8 halves; rows of a core matrix 16 B apart,
// 8-row groups SBO apart, chunks LBO apart.  SBO = 160 and LBO = 16*160 + 16 make the split's STS.32 conflict-free:
// bank = 4*chunk + (pair & 3) + 16*(m-group & 1).
constexpr int kF16SboA = 160;
constexpr int kF16LboA = 16 * kF16SboA + 16;

inline bool fast16_supported(int M, int L) { return M == kF16M && L == kF16L; }

struct F16Taps {
  int qlo, qn;
};
// flags bits [8,12) = QLO, [12,17) = QN as produced by pqmf_build_tables_f32; anything else -> all 16 taps
inline F16Taps fast16_taps_from_flags(unsigned flags) {
  const int qlo = (flags >> 8) & 0xF, qn = (flags >> 12) & 0x1F;
  if (qlo == 2 && qn == 12) return {2, 12};
  return {0, 16};
}
inline unsigned fast16_flags_for_taps(int qlo, int qn) { return ((unsigned)qlo << 8) | ((unsigned)qn << 12); }

__device__ __forceinline__ uint16_t f16_bits(float v) {
  const __half h = __float2half_rn(v);
  return *reinterpret_cast<const uint16_t*>(&h);
}

// =============================================================================================
// analysis
// =============================================================================================
struct F16AnalysisParams {
  const float* x;        // [B, T]
  const float* hist;     // [B, 512] or nullptr
  float* y;              // [B, 16, F]
  float* hist_out;       // [B, 512] or nullptr (streaming: last 512 samples of hist ++ x)
  const float* tables;   // [ g (512) | c1 (16*32) | c2 (16*32) ],  C = c1 + c2, both fp16-representable
  long T, F;
  int B;
  int off;               // 256 offline, 512 streaming
  int parity;            // global parity of frame 0
  long tiles_per_row;
  long n_tiles;
  int num_sms;           // CTAs are launched num_sms at a time: slot of a CTA on its SM = blockIdx.x / num_sms
};

template <int QN, int NXB = 2>
struct F16AnalysisSmem {
  static constexpr int XS = 32 * (64 + QN);     // floats per x window
  static constexpr int NXBUF = NXB;
  static constexpr int APLANE = 4 * kF16LboA;   // one fp16 plane of the A operand [128 x 32]
  static constexpr int BCAT = 4 * 512;          // B = [c1 | c2]: N = 32 rows x K = 32 fp16, 4 K-chunks x (32 rows x 16 B)
  static constexpr int BRES = 4 * 256;          // B = 2^-11 c1 : N = 16 rows x K = 32 fp16
  static constexpr int OFF_X = 0;
  static constexpr int OFF_A = OFF_X + NXBUF * XS * 4;
  static constexpr int OFF_B = OFF_A + 2 * APLANE;
  static constexpr int OFF_BAR = OFF_B + BCAT + BRES;
  static constexpr int BYTES = OFF_BAR + 128;  // 2 NXBUF + 2 mbarriers, the TMEM base slot
};

// Synchronisation is mbarrier-only inside the tile loop (no __syncthreads): warps drift freely and only meet where
// data really flows.
//   xfull[b]  (tx)        TMA -> fold        : x window b has landed
//   xempty[b] (4 warps)   fold -> TMA issuer : all four warps are done reading window b
//   afull     (128 thr)   split -> MMA issuer: the A planes of this tile are complete (and D of tile t-1 was drained)
//   mma_bar   (commit)    MMA -> everyone    : tile's MMAs retired: A planes reusable, D readable
// Warp 0 doubles as MMA issuer and warp 3 as TMA issuer (one elected lane each); they are the only warps that ever
// wait on other warps.  The TMA for tile t + NXBUF is issued as soon as the fold of tile t has released its window.
template <int QLO, int QN>
__global__ void __launch_bounds__(kF16Threads, 4) f16_analysis_kernel(F16AnalysisParams p) {
  using S = F16AnalysisSmem<QN>;
  constexpr int J = kF16J;
  extern __shared__ __align__(128) unsigned char f16_smem[];
  unsigned char* smem = f16_smem;
  float* xs = reinterpret_cast<float*>(smem + S::OFF_X);
  unsigned char* aplane = smem + S::OFF_A;
  unsigned char* bcat = smem + S::OFF_B;
  unsigned char* bres = bcat + S::BCAT;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);  // [NXBUF]
  uint64_t* xempty = xfull + S::NXBUF;                                // [NXBUF]
  uint64_t* afull = xempty + S::NXBUF;
  uint64_t* mma_bar = afull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pp = lane & 15;              // phase pair: phases 2pp, 2pp+1
  const int mg = warp * 2 + (lane >> 4);  // m-group: frame pairs [8 mg, 8 mg + 8)
  const int phi = 2 * pp;

  // ---- one-time setup: barriers, TMEM, B operands in UMMA K-major layout, taps in registers ----
  if (tid == 0) {
    for (int i = 0; i < S::NXBUF; ++i) {
      ptx::mbar_init(&xfull[i], 1);
      ptx::mbar_init(&xempty[i], kF16Threads);
    }
    ptx::mbar_init(afull, kF16Threads);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  for (int e = tid; e < 2 * kF16M * kF16R; e += kF16Threads) {
    const int pl = e / (kF16M * kF16R), rem = e % (kF16M * kF16R);
    const int k = rem / kF16R, r = rem % kF16R;
    const float c = __ldg(p.tables + kF16L + e);
    const int n = pl * 16 + k;  // row of [c1 | c2]
    *reinterpret_cast<uint16_t*>(bcat + (r >> 3) * 512 + n * 16 + (r & 7) * 2) = f16_bits(c);
    if (pl == 0) *reinterpret_cast<uint16_t*>(bres + (r >> 3) * 256 + k * 16 + (r & 7) * 2) = f16_bits(c * (1.f / 2048.f));
  }
  float2 ge[QN], go[QN + 1];
  {
    const float* g = p.tables;
#pragma unroll
    for (int q = 0; q < QN; ++q) ge[q] = make_float2(__ldg(g + phi + 32 * (QLO + q)), __ldg(g + phi + 1 + 32 * (QLO + q)));
    // odd frames: r = (phi + 16) mod 32; for phi < 16 the window starts one 32-sample row later (tap index shifts by one)
    const int ro = (phi + 16) & 31;
    const int shift = (pp < 8) ? 1 : 0;
#pragma unroll
    for (int q = 0; q <= QN; ++q) {
      const int qq = q - shift;
      go[q] = (qq >= 0 && qq < QN) ? make_float2(__ldg(g + ro + 32 * (QLO + qq)), __ldg(g + ro + 1 + 32 * (QLO + qq)))
                                   : make_float2(0.f, 0.f);
    }
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- tile bookkeeping: tile = first + it * stride walked incrementally as (row b, tile-in-row c), no divisions ----
  const unsigned tpr = (unsigned)p.tiles_per_row;
  const unsigned step_b = gridDim.x / tpr, step_c = gridDim.x % tpr;
  auto advance = [&](unsigned& b, unsigned& c) {
    b += step_b;
    c += step_c;
    if (c >= tpr) {
      c -= tpr;
      ++b;
    }
  };
  // TMA WARP ONLY: zero the out-of-range part of an edge window (all lanes), then one lane arms the barrier and
  // issues the TMA bulk copies.  The arrive releases the zero fill to the waiting warps.
  auto stage_tile = [&](unsigned b, unsigned c, int buf) {
    const long n0 = (long)c * kF16TileFrames;
    const long s0 = n0 * 16 + 32 * QLO - p.off;  // first sample of the window, multiple of 16
    float* dst = xs + buf * S::XS;
    const long lo = max(s0, 0L), hi = min(s0 + S::XS, p.T);      // part inside x
    const long hlo = max(s0, -512L), hhi = min(s0 + S::XS, 0L);  // part inside the history
    const bool use_hist = p.hist != nullptr && hhi > hlo;
    if (s0 < 0 || s0 + S::XS > p.T) {
      for (int u = lane; u < S::XS; u += 32) {
        const long s = s0 + u;
        const bool from_x = s >= 0 && s < p.T;
        const bool from_h = use_hist && s >= hlo && s < hhi;
        if (!from_x && !from_h) dst[u] = 0.f;
      }
      __syncwarp();
    }
    if (lane == 0) {
      uint32_t bytes = 0;
      if (hi > lo) bytes += (uint32_t)(hi - lo) * 4;
      if (use_hist) bytes += (uint32_t)(hhi - hlo) * 4;
      ptx::mbar_arrive_expect_tx(&xfull[buf], bytes);
      if (hi > lo) ptx::bulk_g2s(dst + (lo - s0), p.x + (size_t)b * p.T + lo, (uint32_t)(hi - lo) * 4, &xfull[buf]);
      if (use_hist) ptx::bulk_g2s(dst + (hlo - s0), p.hist + (size_t)b * 512 + (512 + hlo), (uint32_t)(hhi - hlo) * 4, &xfull[buf]);
    }
  };
  // epilogue of one tile: D (TMEM) -> registers -> main + correction columns -> sign mask -> coalesced sub-band rows
  auto epilogue = [&](unsigned b, unsigned c, int dbuf) {
    const long n0 = (long)c * kF16TileFrames;
    uint32_t r[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dbuf * 32), r);
    ptx::tmem_ld_wait();
    const long n = n0 + tid;
    if (n < p.F) {
      const uint32_t flip = (((n + p.parity) & 1) == 0) ? 0x80000000u : 0u;
      float* yp = p.y + (size_t)b * kF16M * p.F + n;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float v = __uint_as_float(r[k]) + __uint_as_float(r[16 + k]);
        __stcs(yp + (size_t)k * p.F, __uint_as_float(__float_as_uint(v) ^ ((k & 1) ? flip : 0u)));
      }
    }
    // streaming: the CTA that owns the last tile of a row also rolls that row's history
    if (p.hist_out != nullptr && n0 + kF16TileFrames >= p.F) {
      const long cpos = p.T - 512 + tid * 4;  // position of this float4 in x (negative: still in the old history)
      float4 v;
      if (cpos >= 0) v = *reinterpret_cast<const float4*>(p.x + (size_t)b * p.T + cpos);
      else v = *reinterpret_cast<const float4*>(p.hist + (size_t)b * 512 + (512 + cpos));
      *reinterpret_cast<float4*>(p.hist_out + (size_t)b * 512 + tid * 4) = v;
    }
  };

  const long first = blockIdx.x, stride = gridDim.x;
  unsigned cur_b = blockIdx.x / tpr, cur_c = blockIdx.x % tpr;  // tile being folded
  unsigned nxt_b = cur_b, nxt_c = cur_c;                        // tile being staged (NXBUF iterations ahead)
  constexpr int mma_warp = 0, kTmaWarp = 3;  // (rotating the issuer roles over warps / CTAs measured no gain)
  if (warp == kTmaWarp) {  // prologue: stage the first NXBUF tiles
    for (int i = 0; i < S::NXBUF; ++i) {
      const long tile = first + (long)i * stride;
      if (tile < p.n_tiles) stage_tile(nxt_b, nxt_c, i);
      advance(nxt_b, nxt_c);
    }
  } else {
    for (int i = 0; i < S::NXBUF; ++i) advance(nxt_b, nxt_c);
  }

  unsigned it = 0;
  unsigned prev_b = 0, prev_c = 0;
  for (long tile = first; tile < p.n_tiles; tile += stride, ++it) {
    const int buf = (int)(it % S::NXBUF);
    const uint32_t xphase = (it / S::NXBUF) & 1;
    ptx::mbar_wait(&xfull[buf], xphase);

    // ---------------- fold: 32 FMA / sample on packed fp32 ----------------
    float2 ve[J], vo[J];
#pragma unroll
    for (int j = 0; j < J; ++j) ve[j] = vo[j] = make_float2(0.f, 0.f);
    {
      const float2* zp = reinterpret_cast<const float2*>(xs + buf * S::XS + 32 * (mg * J) + phi);
#pragma unroll
      for (int i = 0; i < J + QN; ++i) {
        const float2 z = zp[16 * i];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int q = i - j;
          if (q >= 0 && q < QN) ve[j] = ptx::ffma2(ge[q], z, ve[j]);
          if (q >= 0 && q <= QN) vo[j] = ptx::ffma2(go[q], z, vo[j]);
        }
      }
    }
    ptx::mbar_arrive(&xempty[buf]);  // this thread no longer reads x window `buf`
    if (warp == kTmaWarp) {
      // x window `buf` is free once all four warps released it: stage the tile NXBUF iterations ahead
      const long next = tile + (long)S::NXBUF * stride;
      if (next < p.n_tiles) {
        ptx::mbar_wait(&xempty[buf], xphase);
        stage_tile(nxt_b, nxt_c, buf);
      }
      __syncwarp();
    }
    advance(nxt_b, nxt_c);  // every warp tracks the staged tile (the issuer role may rotate)
    // the previous tile's MMAs must have finished reading the A planes before they are overwritten
    if (it > 0) ptx::mbar_wait(mma_bar, (uint32_t)((it - 1) & 1));
    // ---------------- two-term fp16 split, stored as the UMMA A operand (h1 plane, h2 plane) ----------------
    {
      const int po = (pp + 8) & 15;  // K position of the odd-frame columns r = (phi + 16) mod 32
      unsigned char* ae = aplane + (pp >> 2) * kF16LboA + (pp & 3) * 4 + (2 * mg) * kF16SboA;
      unsigned char* ao = aplane + (po >> 2) * kF16LboA + (po & 3) * 4 + (2 * mg) * kF16SboA;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int re = 2 * j, rd = 2 * j + 1;  // row within this m-group's 16 rows (even frame, odd frame)
        uint32_t h1, h2;
        ptx::split_f16x2(ve[j], h1, h2);
        *reinterpret_cast<uint32_t*>(ae + (re >> 3) * kF16SboA + (re & 7) * 16) = h1;
        *reinterpret_cast<uint32_t*>(ae + (re >> 3) * kF16SboA + (re & 7) * 16 + S::APLANE) = h2;
        ptx::split_f16x2(vo[j], h1, h2);
        *reinterpret_cast<uint32_t*>(ao + (rd >> 3) * kF16SboA + (rd & 7) * 16) = h1;
        *reinterpret_cast<uint32_t*>(ao + (rd >> 3) * kF16SboA + (rd & 7) * 16 + S::APLANE) = h2;
      }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();  // also orders this thread's tcgen05.ld of the previous epilogue before the next MMAs
    ptx::mbar_arrive(afull);
    if (warp == mma_warp) {
      // ---------------- modulation: D[128 x 32] = h1 [c1 | c2];  D[:, 0:16] += h2 (2^-11 c1) ----------------
      // the whole warp waits (warp-uniform control flow) and one elected lane issues, so the descriptors stay in uniform
      // registers (under `if (lane == 0)` ptxas wraps every tcgen05.mma in a uniformisation loop)
      ptx::mbar_wait(afull, (uint32_t)(it & 1));
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        const uint32_t d = tmem + (uint32_t)((it & 1) * 32);
        const uint64_t da1 = ptx::umma_desc(ptx::smem_u32(aplane), kF16LboA, kF16SboA);
        const uint64_t da2 = ptx::umma_desc(ptx::smem_u32(aplane) + S::APLANE, kF16LboA, kF16SboA);
        const uint64_t db_cat = ptx::umma_desc(ptx::smem_u32(bcat), 512, 128), db_res = ptx::umma_desc(ptx::smem_u32(bres), 256, 128);
        constexpr uint32_t idesc32 = ptx::umma_idesc_f16(128, 32), idesc16 = ptx::umma_idesc_f16(128, 16);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          ptx::umma_f16(d, da1 + (uint64_t)(ks * 2 * kF16LboA / 16), db_cat + (uint64_t)(ks * 2 * 512 / 16), idesc32, ks != 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          ptx::umma_f16(d, da2 + (uint64_t)(ks * 2 * kF16LboA / 16), db_res + (uint64_t)(ks * 2 * 256 / 16), idesc16, true);
        ptx::umma_commit(mma_bar);
      }
      __syncwarp();
    }
    // ---------------- epilogue of the PREVIOUS tile overlaps this tile's MMAs ----------------
    if (it > 0) {
      ptx::tc_fence_after();
      epilogue(prev_b, prev_c, (int)((it - 1) & 1));
    }
    prev_b = cur_b;
    prev_c = cur_c;
    advance(cur_b, cur_c);
  }
  if (it > 0) {
    ptx::mbar_wait(mma_bar, (uint32_t)((it - 1) & 1));
    ptx::tc_fence_after();
    epilogue(prev_b, prev_c, (int)((it - 1) & 1));
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

inline bool fast16_analysis_ok(const float* x, const float* y, long T, long F) {
  return (T % 16) == 0 && F == T / 16 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 4) == 0 && T > 0;
}

template <int QLO, int QN>
int f16_launch_analysis(F16AnalysisParams p, cudaStream_t st) {
  using S = F16AnalysisSmem<QN>;
  auto kern = f16_analysis_kernel<QLO, QN>;
  constexpr int kCtasPerSm = 4;  // matches __launch_bounds__ and the shared-memory footprint
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
  if (p.tiles_per_row >= (1L << 31) || p.n_tiles >= (1L << 40)) return -2;
  long grid = (long)sm_count[dev] * kCtasPerSm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  p.num_sms = sm_count[dev];
  kern<<<(unsigned)grid, kF16Threads, S::BYTES, st>>>(p);
  return (int)cudaGetLastError();
}

// x [B,T] (+ hist [B,512]) -> y [B,16,F] (+ hist_out).  flags carry the tap window chosen by pqmf_build_tables_f32.
inline int fast16_analysis(const float* x, const float* hist, float* y, float* hist_out, const float* tables, int B, long T, long F,
                           int off, int parity, unsigned flags, cudaStream_t st) {
  F16AnalysisParams p{};
  p.x = x; p.hist = hist; p.y = y; p.hist_out = hist_out; p.tables = tables;
  p.T = T; p.F = F; p.B = B; p.off = off; p.parity = parity & 1;
  p.tiles_per_row = (F + kF16TileFrames - 1) / kF16TileFrames;
  p.n_tiles = p.tiles_per_row * B;
  if (hist != nullptr && ((uintptr_t)hist % 16 || (uintptr_t)hist_out % 16)) return -2;
  const F16Taps t = fast16_taps_from_flags(flags);
  if (t.qn == 12) return f16_launch_analysis<2, 12>(p, st);
  return f16_launch_analysis<0, 16>(p, st);
}

}  // namespace pqmf
