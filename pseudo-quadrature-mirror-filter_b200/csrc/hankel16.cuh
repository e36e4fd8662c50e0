// n_band = 16 PQMF as an implicit-Hankel GEMM on the 5th-gen tensor cores (tcgen05 / TMEM), exact in the bank `hk`.
//
// Direct form (reference pqmf.py:115-130 / 160-177 + reverse_half :13-22, and :133-157 / 180-199 / 345-354):
//   analysis : y[k, n]      = sigma(k, n) * sum_j hk[k, j] * X[16 n + j - off]
//   synthesis: out[16 f + p] = 16 * sum_d sum_k sigma(k, n) S[k, n = f + o - d] * hk[k, 16 d + p]
// Both are GEMMs whose A operand is a HANKEL matrix of a contiguous sequence with a hop of 16 elements:
//   analysis : A[n, j]      = X[16 n + j]            (the signal itself)
//   synthesis: A[f, (d, k)] = S^T[f - d, k]          (sub-band frames, 16 bands per frame)
// In fp16 a hop of 16 elements is 32 bytes -- exactly the row pitch of the K-major SWIZZLE_32B UMMA layout.  So the
// operand of K-step s (16 taps / one frame of 16 bands) is "the sequence viewed as 32-byte rows, starting at row s":
// the data is written to shared memory ONCE and only the descriptor start address moves (experiments/probe_hankel.cu).
// No window fold, no im2col copy, no overlap-add FIR: the CUDA cores only convert and move data.
//
// Precision: two-term fp16 split of both operands with exact products in the fp32 accumulator,
//   x = h1 + h2 (h1 = fp16(x), h2 = fp16(x - h1)),   2^s hk = c1 + c2,
//   D[:, 0:32] = sum_s h1_s [c1 | c2]_s^T  (N = 32),   D[:, 0:16] += sum_s h2_s c1_s^T  (N = 16),   result = 2^-s (D[:, :16] + D[:, 16:])
// dropping only h2 c2 (2^-22 relative).  Zero taps of the centre-padded prototype are skipped: K = 32 QN (384 of 512 at
// attenuation 100), i.e. 24 + 24 MMAs of K = 16 per 128-frame tile = 576 tensor-pipe cycles per 2048 samples, below the
// 666 cycles the same tile costs in HBM time at the measured 6.55 TB/s.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "ptx.cuh"

namespace pqmf {

constexpr int kH16Threads = 128;
constexpr int kH16Frames = 128;   // frames per tile = rows of one UMMA
constexpr int kH16ScaleLog2 = 10;  // the bank is stored as fp16 split of 2^10 hk (keeps c2 in the fp16 normal range)

inline bool hankel16_supported(int M, int L) { return M == 16 && L == 512; }

// K-major SWIZZLE_32B shared-memory matrix descriptor: rows 32 B apart, 8-row groups 256 B apart, 16-byte chunk index
// XORed with address bit 7.  (LBO is unused for swizzled K-major layouts; 1 = 16 B as CUTLASS sets it.)
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
// byte offset of element e (fp16) of a sequence stored in that layout (plane base 256-byte aligned)
__device__ __forceinline__ uint32_t sw32_offset(uint32_t byte_lin) { return byte_lin ^ (((byte_lin >> 7) & 1u) << 4); }

// two-term fp16 split of a pair, no scaling of the residual (its absolute rounding error is < 2^-25: negligible)
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& h1_bits, uint32_t& h2_bits) {
  const __half2 h1 = __floats2half2_rn(a, b);
  const float2 h1f = __half22float2(h1);
  const __half2 h2 = __floats2half2_rn(a - h1f.x, b - h1f.y);
  h1_bits = *reinterpret_cast<const uint32_t*>(&h1);
  h2_bits = *reinterpret_cast<const uint32_t*>(&h2);
}

// =============================================================================================
// analysis
// =============================================================================================
struct H16AnalysisParams {
  const float* x;         // [B, T]
  const float* hist;      // [B, 512] or nullptr
  float* y;               // [B, 16, F]
  float* hist_out;        // [B, 512] or nullptr
  const uint16_t* bank;   // fp16 [KT/8][32][8]: UMMA K-major no-swizzle image of [c1 | c2] over the active taps
  long T, F;
  int off;                // 256 offline, 512 streaming
  int parity;
  long tiles_per_row, n_tiles;
};

template <int KT>
struct H16AnalysisSmem {
  static constexpr int KS = KT / 16;                   // K-steps
  static constexpr int ROWS = kH16Frames + KS;         // 32-byte rows of the fp16 window (one spare)
  static constexpr int XS = 16 * ROWS;                 // fp32 samples staged per tile
  static constexpr int PLANE = ((ROWS * 32 + 1023) / 1024) * 1024;
  static constexpr int NXBUF = 2;
  static constexpr int BANK = KT * 32 * 2;             // bytes
  static constexpr int OFF_BANK = 0;
  static constexpr int OFF_P = OFF_BANK + BANK;        // [2 buffers][h1, h2]
  static constexpr int OFF_X = OFF_P + 4 * PLANE;
  static constexpr int OFF_BAR = OFF_X + NXBUF * XS * 4;
  static constexpr int BYTES = OFF_BAR + 128;
};

// mbarriers:  xfull[b] (tx) TMA -> convert | xempty[b] (128) convert -> TMA issuer | pfull (128) convert -> MMA issuer |
//             mma_bar (commit) MMA -> epilogue / plane reuse | bankfull (tx) one-off bank load
template <int JLO, int KT>
__global__ void __launch_bounds__(kH16Threads) h16_analysis_kernel(H16AnalysisParams p) {
  using S = H16AnalysisSmem<KT>;
  constexpr int KS = S::KS;
  extern __shared__ __align__(1024) unsigned char h16_smem[];
  unsigned char* smem = h16_smem;
  unsigned char* bank = smem + S::OFF_BANK;
  unsigned char* planes = smem + S::OFF_P;
  float* xs = reinterpret_cast<float*>(smem + S::OFF_X);
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* xempty = xfull + S::NXBUF;
  uint64_t* pfull = xempty + S::NXBUF;  // [2]: one per plane buffer (a fast thread may arrive for tile t+1 before tile t's phase closes)
  uint64_t* mma_bar = pfull + 2;  // [2]: one per accumulator buffer, so a late waiter can never see the phase wrap
  uint64_t* bankfull = mma_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bankfull + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kMmaWarp = 0, kTmaWarp = 2;

  if (tid == 0) {
    for (int i = 0; i < S::NXBUF; ++i) {
      ptx::mbar_init(&xfull[i], 1);
      ptx::mbar_init(&xempty[i], kH16Threads);
    }
    ptx::mbar_init(&pfull[0], kH16Threads);
    ptx::mbar_init(&pfull[1], kH16Threads);
    ptx::mbar_init(&mma_bar[0], 1);
    ptx::mbar_init(&mma_bar[1], 1);
    ptx::mbar_init(bankfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {  // the bank image is already in UMMA layout: one bulk copy
    ptx::mbar_arrive_expect_tx(bankfull, S::BANK);
    ptx::bulk_g2s(bank, p.bank, S::BANK, bankfull);
  }

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
  // TMA WARP: stage the fp32 window of one tile (zero fill of out-of-range parts by all lanes, bulk copies by one lane)
  auto stage_tile = [&](unsigned b, unsigned c, int buf) {
    const long s0 = (long)c * (kH16Frames * 16) + JLO - p.off;  // first sample of the window, multiple of 16
    float* dst = xs + buf * S::XS;
    const long lo = max(s0, 0L), hi = min(s0 + S::XS, p.T);
    const long hlo = max(s0, -512L), hhi = min(s0 + S::XS, 0L);
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
  // D (TMEM) -> main + correction columns, 2^-s, sign mask -> 16 coalesced sub-band rows
  auto epilogue = [&](unsigned b, unsigned c, int dbuf) {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dbuf * 32);
    const long n = (long)c * kH16Frames + tid;
    const uint32_t flip = (((tid + p.parity) & 1) == 0) ? 0x80000000u : 0u;  // tiles start on even frames
    float* yp = p.y + (size_t)b * 16 * p.F + n;
    const float scale = 1.0f / (float)(1 << kH16ScaleLog2);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r0[8], r1[8];
      ptx::tmem_ld8(taddr + half * 8, r0);
      ptx::tmem_ld8(taddr + 16 + half * 8, r1);
      ptx::tmem_ld_wait();
      if (n < p.F) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = half * 8 + kk;
          const float v = (__uint_as_float(r0[kk]) + __uint_as_float(r1[kk])) * scale;
          __stcs(yp, __uint_as_float(__float_as_uint(v) ^ ((k & 1) ? flip : 0u)));
          yp += p.F;
        }
      }
    }
    if (p.hist_out != nullptr && c + 1 == tpr) {  // streaming: last tile of a row rolls that row's history
      const long cpos = p.T - 512 + tid * 4;
      float4 v;
      if (cpos >= 0) v = *reinterpret_cast<const float4*>(p.x + (size_t)b * p.T + cpos);
      else v = *reinterpret_cast<const float4*>(p.hist + (size_t)b * 512 + (512 + cpos));
      *reinterpret_cast<float4*>(p.hist_out + (size_t)b * 512 + tid * 4) = v;
    }
  };

  unsigned cur_b = blockIdx.x / tpr, cur_c = blockIdx.x % tpr;
  unsigned nxt_b = cur_b, nxt_c = cur_c;
  const unsigned n_iter = (unsigned)((p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  if (warp == kTmaWarp) {
    for (int i = 0; i < S::NXBUF; ++i) {
      if ((unsigned)i < n_iter) stage_tile(nxt_b, nxt_c, i);
      advance(nxt_b, nxt_c);
    }
  }
  const uint32_t bank_addr = ptx::smem_u32(bank);
  const uint32_t plane_addr = ptx::smem_u32(planes);
  constexpr uint32_t idesc32 = ptx::umma_idesc_f16(128, 32), idesc16 = ptx::umma_idesc_f16(128, 16);

  unsigned prev_b = 0, prev_c = 0;
  for (unsigned it = 0; it < n_iter; ++it) {
    const int xb = (int)(it & 1);
    const int pb = (int)(it & 1);
    ptx::mbar_wait(&xfull[xb], (it >> 1) & 1);
    // ---------------- fp32 window -> two fp16 planes in the swizzled 32-byte-row layout ----------------
    // (planes[pb] were last read by the MMAs of tile it-2, whose completion every thread observed before its epilogue)
    {
      const float* src = xs + xb * S::XS;
      unsigned char* p1 = planes + (2 * pb) * S::PLANE;
      unsigned char* p2 = p1 + S::PLANE;
#pragma unroll
      for (int q = tid; q < S::XS / 4; q += kH16Threads) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * q);
        uint2 a, bq;
        split2_f16(v.x, v.y, a.x, bq.x);
        split2_f16(v.z, v.w, a.y, bq.y);
        const uint32_t o = sw32_offset((uint32_t)q * 8u);
        *reinterpret_cast<uint2*>(p1 + o) = a;
        *reinterpret_cast<uint2*>(p2 + o) = bq;
      }
    }
    ptx::mbar_arrive(&xempty[xb]);
    if (warp == kTmaWarp) {
      if (it + S::NXBUF < n_iter) {
        ptx::mbar_wait(&xempty[xb], (it >> 1) & 1);
        stage_tile(nxt_b, nxt_c, xb);
      }
      advance(nxt_b, nxt_c);
      __syncwarp();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    ptx::mbar_arrive(&pfull[it & 1]);
    if (warp == kMmaWarp) {
      // whole warp waits (warp-uniform control flow), one elected lane issues: descriptors stay in uniform registers
      if (it == 0) ptx::mbar_wait(bankfull, 0);
      ptx::mbar_wait(&pfull[it & 1], (it >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        const uint32_t d = tmem + (uint32_t)((it & 1) * 32);
        const uint64_t da1 = umma_desc_sw32(plane_addr + (2 * pb) * S::PLANE);
        const uint64_t da2 = umma_desc_sw32(plane_addr + (2 * pb + 1) * S::PLANE);
        const uint64_t db = ptx::umma_desc(bank_addr, 512, 128);
#pragma unroll
        for (int s = 0; s < KS; ++s)  // +32 B per K-step on A (start-address field counts 16-byte units), +1024 B on B
          ptx::umma_f16(d, da1 + (uint64_t)(2 * s), db + (uint64_t)(64 * s), idesc32, s != 0);
#pragma unroll
        for (int s = 0; s < KS; ++s)
          ptx::umma_f16(d, da2 + (uint64_t)(2 * s), db + (uint64_t)(64 * s), idesc16, true);
        ptx::umma_commit(&mma_bar[it & 1]);
      }
      __syncwarp();
    }
    if (it > 0) {
      ptx::mbar_wait(&mma_bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
      ptx::tc_fence_after();
      epilogue(prev_b, prev_c, (int)((it - 1) & 1));
    }
    prev_b = cur_b;
    prev_c = cur_c;
    advance(cur_b, cur_c);
  }
  ptx::mbar_wait(&mma_bar[(n_iter - 1) & 1], ((n_iter - 1) >> 1) & 1);
  ptx::tc_fence_after();
  epilogue(prev_b, prev_c, (int)((n_iter - 1) & 1));
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

inline bool hankel16_analysis_ok(const float* x, const float* y, long T, long F) {
  return T > 0 && (T % 16) == 0 && F == T / 16 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 4) == 0;
}

template <int JLO, int KT>
int h16_launch_analysis(const H16AnalysisParams& p, cudaStream_t st) {
  using S = H16AnalysisSmem<KT>;
  auto kern = h16_analysis_kernel<JLO, KT>;
  static int sm_count[64] = {0};
  static int ctas_per_sm = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sm_count[dev] == 0) {
    int n = 0, smem_sm = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
    if (e != cudaSuccess) return (int)e;
    int c = smem_sm > 0 ? smem_sm / (S::BYTES + 1024) : 2;
    if (c > 4) c = 4;
    if (c < 1) c = 1;
    ctas_per_sm = c;
    sm_count[dev] = n > 0 ? n : 148;
  }
  if (p.tiles_per_row >= (1L << 31) || p.n_tiles >= (1L << 40)) return -2;
  long grid = (long)sm_count[dev] * ctas_per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, kH16Threads, S::BYTES, st>>>(p);
  return (int)cudaGetLastError();
}

// =============================================================================================
// synthesis
// =============================================================================================
struct H16SynthesisParams {
  const float* s;         // [B, 16, F]
  const float* hist;      // [B, 16, 32] or nullptr
  float* out;             // [B, 16 F]
  float* hist_out;        // [B, 16, 32] or nullptr
  const uint16_t* bank;   // fp16 [KT/8][32][8]: image of [c1 | c2]^T per K-step (one frame lag d, 16 bands), rows = output phase
  long F;
  int o;                  // off2 / 16: 16 (PQMF.inverse), 15 (CachedPQMF.inverse), -1 (streaming)
  int parity;
  long tiles_per_row, n_tiles;
};

template <int KT>
struct H16SynthesisSmem {
  static constexpr int KS = KT / 16;
  static constexpr int ROWS = kH16Frames + KS;   // sub-band frames per tile (one spare row)
  static constexpr int PLANE = ((ROWS * 32 + 1023) / 1024) * 1024;
  static constexpr int BANK = KT * 32 * 2;
  static constexpr int OFF_BANK = 0;
  static constexpr int OFF_P = OFF_BANK + BANK;
  static constexpr int OFF_BAR = OFF_P + 4 * PLANE;
  static constexpr int BYTES = OFF_BAR + 128;
};

template <int JLO, int KT>
__global__ void __launch_bounds__(kH16Threads) h16_synthesis_kernel(H16SynthesisParams p) {
  using S = H16SynthesisSmem<KT>;
  constexpr int KS = S::KS;
  constexpr int DMAX = (JLO + KT) / 16 - 1;      // largest frame lag with a non-zero tap
  constexpr int EXTRA = KS - 1;                  // rows 128 .. 128 + KS - 2 are the halo
  extern __shared__ __align__(1024) unsigned char h16s_smem[];
  unsigned char* smem = h16s_smem;
  unsigned char* bank = smem + S::OFF_BANK;
  unsigned char* planes = smem + S::OFF_P;
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);  // [2]: one per plane buffer
  uint64_t* mma_bar = pfull + 2;  // [2]: one per accumulator buffer, so a late waiter can never see the phase wrap
  uint64_t* bankfull = mma_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bankfull + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kMmaWarp = 1;

  if (tid == 0) {
    ptx::mbar_init(&pfull[0], kH16Threads);
    ptx::mbar_init(&pfull[1], kH16Threads);
    ptx::mbar_init(&mma_bar[0], 1);
    ptx::mbar_init(&mma_bar[1], 1);
    ptx::mbar_init(bankfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(bankfull, S::BANK);
    ptx::bulk_g2s(bank, p.bank, S::BANK, bankfull);
  }

  const unsigned tpr = (unsigned)p.tiles_per_row;
  const unsigned step_b = gridDim.x / tpr, step_c = gridDim.x % tpr;
  unsigned b = blockIdx.x / tpr, c = blockIdx.x % tpr;
  const unsigned n_iter = (unsigned)((p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  // sub-band frames of one tile: row r <-> frame n = 128 c + o - DMAX + r.  Thread t owns row t (and halo row 128 + t).
  float va[16], vb[16];
  bool flip_a = false, flip_b = false;
  auto load_frame = [&](const float* sb, const float* hb, long n, float (&v)[16], bool& flip) {
    if (n >= 0 && n < p.F) {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldcs(sb + (size_t)k * p.F + n);
    } else if (hb != nullptr && n < 0 && n >= -32) {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldg(hb + k * 32 + (32 + n));
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.f;
    }
    flip = ((n + p.parity) & 1) == 0;
  };
  auto load_rows = [&](unsigned bb, unsigned cc) {
    const float* sb = p.s + (size_t)bb * 16 * p.F;
    const float* hb = p.hist ? p.hist + (size_t)bb * 512 : nullptr;
    const long n0 = (long)cc * kH16Frames + p.o - DMAX;
    load_frame(sb, hb, n0 + tid, va, flip_a);
    if (tid < EXTRA) load_frame(sb, hb, n0 + kH16Frames + tid, vb, flip_b);
  };
  // one frame (16 bands) -> sign mask, two-term fp16 split -> one 32-byte row of each plane (swizzled)
  auto store_frame = [&](unsigned char* p1, int row, float (&v)[16], bool flip) {
    if (flip) {
#pragma unroll
      for (int k = 1; k < 16; k += 2) v[k] = -v[k];
    }
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint4 h1, h2;
      split2_f16(v[8 * ch + 0], v[8 * ch + 1], h1.x, h2.x);
      split2_f16(v[8 * ch + 2], v[8 * ch + 3], h1.y, h2.y);
      split2_f16(v[8 * ch + 4], v[8 * ch + 5], h1.z, h2.z);
      split2_f16(v[8 * ch + 6], v[8 * ch + 7], h1.w, h2.w);
      const uint32_t o = sw32_offset((uint32_t)row * 32u + 16u * ch);
      *reinterpret_cast<uint4*>(p1 + o) = h1;
      *reinterpret_cast<uint4*>(p1 + S::PLANE + o) = h2;
    }
  };
  // D (TMEM) -> 16 consecutive output samples of frame f -> four 16-byte stores (a warp writes 2 KB contiguous)
  auto epilogue = [&](unsigned bb, unsigned cc, int dbuf) {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dbuf * 32);
    const long f = (long)cc * kH16Frames + tid;
    float* op = p.out + ((size_t)bb * p.F + f) * 16;
    const float scale = 1.0f / (float)(1 << kH16ScaleLog2);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r0[8], r1[8];
      ptx::tmem_ld8(taddr + half * 8, r0);
      ptx::tmem_ld8(taddr + 16 + half * 8, r1);
      ptx::tmem_ld_wait();
      if (f < p.F) {
        float4 w0, w1;
        w0.x = (__uint_as_float(r0[0]) + __uint_as_float(r1[0])) * scale;
        w0.y = (__uint_as_float(r0[1]) + __uint_as_float(r1[1])) * scale;
        w0.z = (__uint_as_float(r0[2]) + __uint_as_float(r1[2])) * scale;
        w0.w = (__uint_as_float(r0[3]) + __uint_as_float(r1[3])) * scale;
        w1.x = (__uint_as_float(r0[4]) + __uint_as_float(r1[4])) * scale;
        w1.y = (__uint_as_float(r0[5]) + __uint_as_float(r1[5])) * scale;
        w1.z = (__uint_as_float(r0[6]) + __uint_as_float(r1[6])) * scale;
        w1.w = (__uint_as_float(r0[7]) + __uint_as_float(r1[7])) * scale;
        __stcs(reinterpret_cast<float4*>(op + half * 8), w0);
        __stcs(reinterpret_cast<float4*>(op + half * 8 + 4), w1);
      }
    }
    if (p.hist_out != nullptr && cc + 1 == tpr) {  // streaming: last tile of a row rolls that row's sub-band history
      const float* sb = p.s + (size_t)bb * 16 * p.F;
      const float* hb = p.hist + (size_t)bb * 512;
      for (int e = tid; e < 512; e += kH16Threads) {
        const int k = e >> 5;
        const long cpos = p.F - 32 + (e & 31);
        p.hist_out[(size_t)bb * 512 + e] = (cpos >= 0) ? sb[(size_t)k * p.F + cpos] : hb[k * 32 + 32 + cpos];
      }
    }
  };

  load_rows(b, c);
  const uint32_t bank_addr = ptx::smem_u32(bank);
  const uint32_t plane_addr = ptx::smem_u32(planes);
  constexpr uint32_t idesc32 = ptx::umma_idesc_f16(128, 32), idesc16 = ptx::umma_idesc_f16(128, 16);
  unsigned prev_b = 0, prev_c = 0;
  for (unsigned it = 0; it < n_iter; ++it) {
    const int pb = (int)(it & 1);
    unsigned char* p1 = planes + (2 * pb) * S::PLANE;
    store_frame(p1, tid, va, flip_a);
    if (tid < EXTRA) store_frame(p1, kH16Frames + tid, vb, flip_b);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    ptx::mbar_arrive(&pfull[it & 1]);
    if (warp == kMmaWarp) {
      // whole warp waits (warp-uniform control flow), one elected lane issues: descriptors stay in uniform registers
      if (it == 0) ptx::mbar_wait(bankfull, 0);
      ptx::mbar_wait(&pfull[it & 1], (it >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        const uint32_t d = tmem + (uint32_t)((it & 1) * 32);
        const uint64_t da1 = umma_desc_sw32(plane_addr + (2 * pb) * S::PLANE);
        const uint64_t da2 = umma_desc_sw32(plane_addr + (2 * pb + 1) * S::PLANE);
        const uint64_t db = ptx::umma_desc(bank_addr, 512, 128);
#pragma unroll
        for (int s2 = 0; s2 < KS; ++s2)  // +32 B per K-step on A (start-address field counts 16-byte units), +1024 B on B
          ptx::umma_f16(d, da1 + (uint64_t)(2 * s2), db + (uint64_t)(64 * s2), idesc32, s2 != 0);
#pragma unroll
        for (int s2 = 0; s2 < KS; ++s2)
          ptx::umma_f16(d, da2 + (uint64_t)(2 * s2), db + (uint64_t)(64 * s2), idesc16, true);
        ptx::umma_commit(&mma_bar[it & 1]);
      }
      __syncwarp();
    }
    // prefetch the next tile's frames into registers (consumed at the top of the next iteration)
    unsigned nb = b + step_b, nc = c + step_c;
    if (nc >= tpr) {
      nc -= tpr;
      ++nb;
    }
    if (it + 1 < n_iter) load_rows(nb, nc);
    if (it > 0) {
      ptx::mbar_wait(&mma_bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
      ptx::tc_fence_after();
      epilogue(prev_b, prev_c, (int)((it - 1) & 1));
    }
    prev_b = b;
    prev_c = c;
    b = nb;
    c = nc;
  }
  ptx::mbar_wait(&mma_bar[(n_iter - 1) & 1], ((n_iter - 1) >> 1) & 1);
  ptx::tc_fence_after();
  epilogue(prev_b, prev_c, (int)((n_iter - 1) & 1));
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

inline bool hankel16_synthesis_ok(const float* s, const float* out, long F) {
  return F > 0 && ((uintptr_t)s % 4) == 0 && ((uintptr_t)out % 16) == 0;
}

template <int JLO, int KT>
int h16_launch_synthesis(const H16SynthesisParams& p, cudaStream_t st) {
  using S = H16SynthesisSmem<KT>;
  auto kern = h16_synthesis_kernel<JLO, KT>;
  static int sm_count[64] = {0};
  static int ctas_per_sm = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sm_count[dev] == 0) {
    int n = 0, smem_sm = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
    if (e != cudaSuccess) return (int)e;
    int cps = smem_sm > 0 ? smem_sm / (S::BYTES + 1024) : 2;
    if (cps > 4) cps = 4;
    if (cps < 1) cps = 1;
    ctas_per_sm = cps;
    sm_count[dev] = n > 0 ? n : 148;
  }
  if (p.tiles_per_row >= (1L << 31) || p.n_tiles >= (1L << 40)) return -2;
  long grid = (long)sm_count[dev] * ctas_per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<(unsigned)grid, kH16Threads, S::BYTES, st>>>(p);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host: bank images (fp16 bit patterns, UMMA K-major no-swizzle layout [chunk of 8 K][32 rows][8])
// ---------------------------------------------------------------------------------------------
inline void hankel16_build_banks(const float* hk /*[16][512]*/, int jlo, int kt, uint16_t* img_analysis, uint16_t* img_synthesis) {
  auto bits = [](float v) {
    const __half h = __float2half_rn(v);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  };
  const int dmax = (jlo + kt) / 16 - 1;
  const float sa = (float)(1 << kH16ScaleLog2), ss = 16.f * sa;  // synthesis folds the x16 gain of pqmf.py:152
  for (int kc = 0; kc < kt / 8; ++kc)
    for (int nrow = 0; nrow < 32; ++nrow)
      for (int e = 0; e < 8; ++e) {
        const int kk = 8 * kc + e;
        const size_t at = ((size_t)kc * 32 + nrow) * 8 + e;
        {  // analysis: K index = tap, row = band
          const float v = sa * hk[(nrow & 15) * 512 + jlo + kk];
          const float c1 = __half2float(__float2half_rn(v));
          img_analysis[at] = nrow < 16 ? bits(c1) : bits(v - c1);
        }
        {  // synthesis: K index = (frame lag step s, band kb), row = output phase p; lag d = dmax - s
          const int s2 = kk / 16, kb = kk % 16, d = dmax - s2;
          const float v = ss * hk[kb * 512 + 16 * d + (nrow & 15)];
          const float c1 = __half2float(__float2half_rn(v));
          img_synthesis[at] = nrow < 16 ? bits(c1) : bits(v - c1);
        }
      }
}

}  // namespace pqmf
