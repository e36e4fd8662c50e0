// n_band = 16 PQMF, offline fast path: the direct form as an implicit-Hankel GEMM with FOUR frames per operand row.
//
// hankel16.cuh shows that a strided signal can be fed to tcgen05.mma without im2col (one frame = 32 B = the SWIZZLE_32B row
// pitch), but there every 16 taps cost a full shared-memory read of the 128-row A tile (~111 B/sample: smem-bound at ~30 % of
// the HBM roofline).  Here one A row holds FOUR frames (64 samples = 128 B in fp16 = the SWIZZLE_128B row pitch) and the
// bank is replicated at the four frame offsets along N:
//   analysis : D[i, (delta, k)] = sum_kappa X[64 i + kappa] * hk[k, kappa - 16 delta]      = y[k, frame 4 i + delta]
//   synthesis: D[i, (delta, p)] = sum_(e,k) S^T[4 i + o - e, k] * 16 hk[k, 16 (e + delta) + p] = out[16 (4 i + delta) + p]
// so one 128-row MMA (N = 128) yields 512 frames and each signal byte is read from shared memory 4x less often.
// K-step s reads rows starting at byte 128 (s / 4) + 32 (s % 4).
// Precision: the same two-term fp16 split as hankel16.cuh (exact in hk up to 2^-22), columns 0-63 main term, 64-127 the
// c2 correction; the second pass (h2) uses only the c1 half (N = 64).  Edge K-steps whose correction terms are provably
// negligible for the actual bank run the main term only (h4_issue_mmas, hankel4_pick_trim).
// Roles: worker warps convert fp32 -> 2 x fp16 planes (swizzled STS) and drain TMEM; ONE warp only issues the MMAs
// (tcgen05.mma issue blocks the issuing thread while the tensor queue is full).  With PAIR the kernel runs as clusters of two
// CTAs sharing the bank operand (cta_group::2).  DESIGN.md section 5 has the measurements behind each of these choices.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "hankel16.cuh"
#include "ptx.cuh"

namespace pqmf {

constexpr int kH4Workers = 256;                 // convert + drain warps
constexpr int kH4Threads = kH4Workers + 32;     // + one warp that only issues tcgen05.mma (the issue queue blocks the issuing thread)
constexpr int kH4Rows = 128;             // A rows per tile
constexpr int kH4Frames = 4 * kH4Rows;   // 512 frames = 8192 samples per tile

// K-major SWIZZLE_128B descriptor: rows 128 B apart, 8-row groups 1024 B apart, 16-byte chunk index ^= address bits 7-9
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t sw128_offset(uint32_t byte_lin) { return byte_lin ^ (((byte_lin >> 7) & 7u) << 4); }

// PAIR: the kernel runs as a cluster of two CTAs (tcgen05 cta_group::2).  One M = 256 MMA covers the two CTAs' tiles; each
// CTA holds only its half of the bank rows: rank 0 [c1 (64 rows) | c1 rows 0-31 again], rank 1 [c2 (64 rows) | c1 rows 32-63],
// i.e. rows 0-63 are this rank's half of the N = 128 operand and rows 64-95 its half of the N = 64 operand (at the same
// offset in both CTAs, because one descriptor addresses both).  Per SM and K-step the tensor pipe then reads 6 + 5 KB of
// operands instead of 8 + 6 KB (experiments/probe_h4_pair.cu: 2380 vs 2868 cycles per tile for the trimmed schedule).
template <int KT, bool PAIR = false>
struct H4Geometry {
  static constexpr int KS = KT / 16 + 3;                        // K-steps: taps + the three extra frame offsets
  static constexpr int ROWS = kH4Rows + ((KS - 1) >> 2) + 1;    // 128-byte rows of one plane (covers the synthesis pad <= 3 too)
  static constexpr int PLANE = ((ROWS * 128 + 1023) / 1024) * 1024;
  static constexpr int BANK_ROWS = PAIR ? 96 : 128;
  static constexpr int BANK = KS * 2 * BANK_ROWS * 16;          // [2 KS chunks][BANK_ROWS][16 B]
  static constexpr int OFF_BANK = 0;
  static constexpr int OFF_P = OFF_BANK + BANK;                 // [2 buffers][h1, h2]
  static constexpr int OFF_BAR = OFF_P + 4 * PLANE;
  static constexpr int BYTES = OFF_BAR + 128;
};

// one elected lane: the MMAs of a tile, D[:, 0:128] = h1 [c1 | c2]^T, D[:, 0:64] += h2 c1^T.  The first and last `trim`
// K-steps carry only the tails of the prototype: there the two correction terms (h1 c2, h2 c1) are below the error budget
// that pqmf_build_tables_f32 checked against the actual bank, so those steps run h1 c1 alone (N = 64) and no h2 pass.
// `pad` shifts the A windows by whole frames (synthesis alignment).
template <int KS, bool PAIR>
__device__ __forceinline__ void h4_issue_mmas(uint32_t d_tmem, uint32_t plane1_addr, uint32_t plane2_addr, uint32_t bank_addr, int pad, int trim) {
  constexpr uint32_t kBankRows = PAIR ? 96 : 128;
  const uint64_t da1 = umma_desc_sw128(plane1_addr), da2 = umma_desc_sw128(plane2_addr);
  const uint64_t db = ptx::umma_desc(bank_addr, kBankRows * 16, 128);                 // N = 128 operand: bank rows from 0
  const uint64_t db64 = PAIR ? ptx::umma_desc(bank_addr + 64 * 16, kBankRows * 16, 128) : db;  // N = 64 operand
  constexpr uint32_t m = PAIR ? 256 : 128;
  constexpr uint32_t idesc128 = ptx::umma_idesc_f16(m, 128), idesc64 = ptx::umma_idesc_f16(m, 64);
  constexpr uint32_t kStep = 2 * kBankRows;                                          // descriptor units (16 B) per K-step of the bank
  auto a_step = [&](int s) { return (uint64_t)(8 * ((s + pad) >> 2) + 2 * ((s + pad) & 3)); };  // 128 B per 4 frames, 32 B per frame
  auto mma = [&](uint64_t a, uint64_t bdesc, uint32_t idesc, bool acc) {
    if constexpr (PAIR) ptx::umma_pair_f16(d_tmem, a, bdesc, idesc, acc);
    else ptx::umma_f16(d_tmem, a, bdesc, idesc, acc);
  };
  for (int s = trim; s < KS - trim; ++s) mma(da1 + a_step(s), db + (uint64_t)(kStep * s), idesc128, s != trim);
  for (int s = 0; s < trim; ++s) {
    mma(da1 + a_step(s), db64 + (uint64_t)(kStep * s), idesc64, true);
    mma(da1 + a_step(KS - 1 - s), db64 + (uint64_t)(kStep * (KS - 1 - s)), idesc64, true);
  }
  for (int s = trim; s < KS - trim; ++s) mma(da2 + a_step(s), db64 + (uint64_t)(kStep * s), idesc64, true);
}

// =============================================================================================
// analysis
// =============================================================================================
struct H4AnalysisParams {
  const float* x;        // [B, T]
  float* y;              // [B, 16, F]
  const uint16_t* bank;  // fp16 image [2 KS][128][8]
  long T, F;
  int off;               // 256 (offline only: streaming blocks are far smaller than a tile)
  int parity;
  int trim;              // edge K-steps without correction terms (h4_issue_mmas)
  long tiles_per_row, n_tiles;
#ifdef PQMF_H4_TRACE
  long long* trace;      // [iterations][8 warps][8] clock64 stamps of CTA 0 (experiments/trace_h4.cu)
#endif
};

#ifdef PQMF_H4_TRACE
#define H4_STAMP(k) do { if (blockIdx.x == 0 && (tid & 31) == 0 && it < 64) p.trace[((size_t)it * 8 + warp) * 8 + (k)] = clock64(); } while (0)
#else
#define H4_STAMP(k) do { } while (0)
#endif

template <int JLO, int KT, bool PAIR>
__global__ void __launch_bounds__(kH4Threads, 1) h4_analysis_kernel(H4AnalysisParams p) {
  using G = H4Geometry<KT, PAIR>;
  constexpr int KS = G::KS;
  constexpr int NQ = (G::ROWS * 16 + kH4Workers - 1) / kH4Workers;  // float4 loads per thread per tile (row = 16 quads)
  extern __shared__ __align__(1024) unsigned char h4_smem[];
  unsigned char* smem = h4_smem;
  unsigned char* bank = smem + G::OFF_BANK;
  unsigned char* planes = smem + G::OFF_P;
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);  // [2]
  uint64_t* mma_bar = pfull + 2;                                     // [2]
  uint64_t* bankfull = mma_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bankfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int kMmaWarp = kH4Workers / 32;
  // pfull lives in the leader CTA (rank 0): one arrival per worker warp of every CTA of the pair
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  if (tid == 0) {
    ptx::mbar_init(&pfull[0], (kH4Workers / 32) * (PAIR ? 2 : 1));
    ptx::mbar_init(&pfull[1], (kH4Workers / 32) * (PAIR ? 2 : 1));
    ptx::mbar_init(&mma_bar[0], 1);
    ptx::mbar_init(&mma_bar[1], 1);
    ptx::mbar_init(bankfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, 256);
    } else {
      ptx::tmem_alloc(tmem_slot, 256);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  ptx::tc_fence_after();
  const uint32_t pfull_leader = PAIR ? ptx::mapa_shared(ptx::smem_u32(pfull), 0) : 0u;
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(bankfull, G::BANK);
    ptx::bulk_g2s(bank, reinterpret_cast<const unsigned char*>(p.bank) + (size_t)rank * G::BANK, G::BANK, bankfull);  // PAIR: per-rank images
  }

  const unsigned tpr = (unsigned)p.tiles_per_row;
  const unsigned step_b = gridDim.x / tpr, step_c = gridDim.x % tpr;
  unsigned b = blockIdx.x / tpr, c = blockIdx.x % tpr;
  const unsigned first_tile = PAIR ? (blockIdx.x & ~1u) : blockIdx.x;   // both CTAs of a pair run the same number of tiles:
  const unsigned n_iter = (unsigned)((p.n_tiles - first_tile + gridDim.x - 1) / gridDim.x);
  const unsigned n_rows = (unsigned)(p.n_tiles / p.tiles_per_row);      // a tile with b >= n_rows is padding (zeros in, nothing out)
#ifdef PQMF_H4_TRACE
  long long cta_c0 = clock64(), cta_t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_t0));
#endif

  // this thread's share of the fp32 window of one tile (prefetched one tile ahead, straight from global memory)
  float4 xr[NQ];
  auto load_window = [&](unsigned bb, unsigned cc) {
    const long s0 = (long)cc * (kH4Frames * 16) + JLO - p.off;
    const float* xrow = p.x + (size_t)bb * p.T;
#pragma unroll
    for (int r = 0; r < NQ; ++r) {
      const int q = tid + kH4Workers * r;
      const long s = s0 + 4L * q;
      xr[r] = (bb < n_rows && q < G::ROWS * 16 && s >= 0 && s < p.T) ? ptx::ldg128_na(reinterpret_cast<const float4*>(xrow + s)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // D (TMEM) -> y: thread (row i, band half hb) owns frames 4 i .. 4 i + 3 of bands 8 hb .. 8 hb + 7: one float4 per band
  auto epilogue = [&](unsigned bb, unsigned cc, int dbuf) {
    const int i = tid & 127, hb = tid >> 7;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(dbuf * 128 + 8 * hb);
    const long n = (long)cc * kH4Frames + 4 * i;
    const float scale = 1.0f / (float)(1 << kH16ScaleLog2);
    float v[4][8];
#pragma unroll
    for (int dl = 0; dl < 4; ++dl) {
      uint32_t r0[8], r1[8];
      ptx::tmem_ld8(taddr + dl * 16, r0);
      ptx::tmem_ld8(taddr + 64 + dl * 16, r1);
      ptx::tmem_ld_wait();
      // sigma(k, n): odd bands flip on even frames; tiles and 4 i are even, so the frame parity is that of dl (+ p.parity)
      const uint32_t flip = (((dl + p.parity) & 1) == 0) ? 0x80000000u : 0u;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float t = (__uint_as_float(r0[kk]) + __uint_as_float(r1[kk])) * scale;
        v[dl][kk] = __uint_as_float(__float_as_uint(t) ^ ((kk & 1) ? flip : 0u));
      }
    }
    float* yp = p.y + ((size_t)bb * 16 + 8 * hb) * p.F + n;
    if (n + 3 < p.F && (p.F & 3) == 0) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) __stcs(reinterpret_cast<float4*>(yp + (size_t)kk * p.F), make_float4(v[0][kk], v[1][kk], v[2][kk], v[3][kk]));
    } else {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
#pragma unroll
        for (int dl = 0; dl < 4; ++dl)
          if (n + dl < p.F) yp[(size_t)kk * p.F + dl] = v[dl][kk];
    }
  };

  const uint32_t bank_addr = ptx::smem_u32(bank), plane_addr = ptx::smem_u32(planes);
  if (warp == kMmaWarp) {
    // ---- issuer warp: tcgen05.mma issue blocks once the tensor pipe's queue is full (a 54-MMA group keeps the issuing
    //      thread for ~3200 cycles), so this warp does nothing else and the pipe never waits for a converting thread
    for (unsigned it = 0; rank == 0 && it < n_iter; ++it) {
      const int pb = (int)(it & 1);
      ptx::mbar_wait(&pfull[pb], (it >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        h4_issue_mmas<KS, PAIR>(tmem + (uint32_t)(pb * 128), plane_addr + (2 * pb) * G::PLANE, plane_addr + (2 * pb + 1) * G::PLANE, bank_addr, 0, p.trim);
        if constexpr (PAIR) ptx::umma_pair_commit(&mma_bar[pb]);
        else ptx::umma_commit(&mma_bar[pb]);
      }
      __syncwarp();
    }
  } else {
    load_window(b, c);
    unsigned prev_b = 0, prev_c = 0;
    for (unsigned it = 0; it < n_iter; ++it) {
      const int pb = (int)(it & 1);
      H4_STAMP(0);
      // ---- fp32 window -> two fp16 planes (SWIZZLE_128B rows of 64 samples).  planes[pb] were last read by the MMAs of
      //      tile it-2, whose completion this thread observed before draining tile it-2.
      {
        unsigned char* p1 = planes + (2 * pb) * G::PLANE;
        unsigned char* p2 = p1 + G::PLANE;
#pragma unroll
        for (int r = 0; r < NQ; ++r) {
          const int q = tid + kH4Workers * r;
          if (q < G::ROWS * 16) {
            uint2 a, bq;
            split2_f16(xr[r].x, xr[r].y, a.x, bq.x);
            split2_f16(xr[r].z, xr[r].w, a.y, bq.y);
            const uint32_t o = sw128_offset((uint32_t)q * 8u);
            *reinterpret_cast<uint2*>(p1 + o) = a;
            *reinterpret_cast<uint2*>(p2 + o) = bq;
          }
        }
      }
      H4_STAMP(1);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) {
        if (it == 0) ptx::mbar_wait(bankfull, 0);   // this CTA's bank image has landed before its first arrival
        if constexpr (PAIR) ptx::mbar_arrive_cluster(pfull_leader + 8u * pb);
        else ptx::mbar_arrive(&pfull[pb]);
      }
      // prefetch the next tile's window (consumed at the top of the next iteration)
      unsigned nb = b + step_b, nc = c + step_c;
      if (nc >= tpr) {
        nc -= tpr;
        ++nb;
      }
      if (it + 1 < n_iter) load_window(nb, nc);
      H4_STAMP(2);
      H4_STAMP(3);
      if (it > 0) {
        ptx::mbar_wait(&mma_bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
        ptx::tc_fence_after();
        H4_STAMP(4);
        if (prev_b < n_rows) epilogue(prev_b, prev_c, (int)((it - 1) & 1));
      }
      H4_STAMP(5);
      prev_b = b;
      prev_c = c;
      b = nb;
      c = nc;
    }
    ptx::mbar_wait(&mma_bar[(n_iter - 1) & 1], ((n_iter - 1) >> 1) & 1);
    ptx::tc_fence_after();
    if (prev_b < n_rows) epilogue(prev_b, prev_c, (int)((n_iter - 1) & 1));
  }  // workers
  ptx::tc_fence_before();
  __syncthreads();
#ifdef PQMF_H4_TRACE
  if (tid == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.trace[64 * 64 + 2 * blockIdx.x] = clock64() - cta_c0;
    p.trace[64 * 64 + 2 * blockIdx.x + 1] = t1 - cta_t0;
  }
#endif
  if constexpr (PAIR) {
    ptx::cluster_sync_all();   // the peer may still be read by / signalled from the leader's last MMAs
    if (warp == 0) ptx::tmem_dealloc_pair(tmem, 256);
  } else {
    if (warp == 0) ptx::tmem_dealloc(tmem, 256);
  }
}

// =============================================================================================
// synthesis
// =============================================================================================
struct H4SynthesisParams {
  const float* s;        // [B, 16, F]
  float* out;            // [B, 16 F]
  const uint16_t* bank;
  long F;
  int o;                 // off2 / 16: 16 (PQMF.inverse) or 15 (CachedPQMF.inverse)
  int parity;
  int trim;
  long tiles_per_row, n_tiles;
#ifdef PQMF_H4_TRACE
  long long* trace;
#endif
};

constexpr int kH4SynWorkers = 288;                    // nine worker warps: 2 x ROWS (<= 274) load/convert items, one per thread
constexpr int kH4SynThreads = kH4SynWorkers + 32;     // + the issuer warp

template <int JLO, int KT, bool PAIR>
__global__ void __launch_bounds__(kH4SynThreads, 1) h4_synthesis_kernel(H4SynthesisParams p) {
  using G = H4Geometry<KT, PAIR>;
  constexpr int KS = G::KS;
  constexpr int EHI = (JLO + KT) / 16 - 1;        // largest frame lag with a non-zero tap
  static_assert(2 * G::ROWS <= kH4SynWorkers, "one (frame quad, band half) item per worker thread");
  extern __shared__ __align__(1024) unsigned char h4s_smem[];
  unsigned char* smem = h4s_smem;
  unsigned char* bank = smem + G::OFF_BANK;
  unsigned char* planes = smem + G::OFF_P;
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
  uint64_t* mma_bar = pfull + 2;
  uint64_t* bankfull = mma_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bankfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int kMmaWarp = kH4SynWorkers / 32;
  // pfull lives in the leader CTA (rank 0): one arrival per worker warp of every CTA of the pair
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  if (tid == 0) {
    ptx::mbar_init(&pfull[0], (kH4SynWorkers / 32) * (PAIR ? 2 : 1));
    ptx::mbar_init(&pfull[1], (kH4SynWorkers / 32) * (PAIR ? 2 : 1));
    ptx::mbar_init(&mma_bar[0], 1);
    ptx::mbar_init(&mma_bar[1], 1);
    ptx::mbar_init(bankfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, 256);
    } else {
      ptx::tmem_alloc(tmem_slot, 256);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  ptx::tc_fence_after();
  const uint32_t pfull_leader = PAIR ? ptx::mapa_shared(ptx::smem_u32(pfull), 0) : 0u;
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(bankfull, G::BANK);
    ptx::bulk_g2s(bank, reinterpret_cast<const unsigned char*>(p.bank) + (size_t)rank * G::BANK, G::BANK, bankfull);  // PAIR: per-rank images
  }

  const unsigned tpr = (unsigned)p.tiles_per_row;
  const unsigned step_b = gridDim.x / tpr, step_c = gridDim.x % tpr;
  unsigned b = blockIdx.x / tpr, c = blockIdx.x % tpr;
  const unsigned first_tile = PAIR ? (blockIdx.x & ~1u) : blockIdx.x;   // both CTAs of a pair run the same number of tiles:
  const unsigned n_iter = (unsigned)((p.n_tiles - first_tile + gridDim.x - 1) / gridDim.x);
  const unsigned n_rows = (unsigned)(p.n_tiles / p.tiles_per_row);      // a tile with b >= n_rows is padding (zeros in, nothing out)

  // plane frame m <-> sub-band frame n = 512 c + (o - EHI - pad) + m, pad = (o - EHI) mod 4, so that frame quads are
  // 16-byte aligned in global memory; K-step s of row i then reads plane frame 4 i + s + pad.
  const int pad = (p.o - EHI) & 3;
  const int nbase = p.o - EHI - pad;              // multiple of 4 (may be negative)
  const uint32_t bank_addr = ptx::smem_u32(bank), plane_addr = ptx::smem_u32(planes);
  if (warp == kMmaWarp) {
    // ---- issuer warp (see the analysis kernel)
    for (unsigned it = 0; rank == 0 && it < n_iter; ++it) {
      const int pb = (int)(it & 1);
      ptx::mbar_wait(&pfull[pb], (it >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        h4_issue_mmas<KS, PAIR>(tmem + (uint32_t)(pb * 128), plane_addr + (2 * pb) * G::PLANE, plane_addr + (2 * pb + 1) * G::PLANE, bank_addr, pad, p.trim);
        if constexpr (PAIR) ptx::umma_pair_commit(&mma_bar[pb]);
        else ptx::umma_commit(&mma_bar[pb]);
      }
      __syncwarp();
    }
  } else {
    // ---- workers.  Item (m4, ch): frames 4 m4 .. 4 m4 + 3 of bands 8 ch .. 8 ch + 7 = eight float4 loads (prefetched one
    //      tile ahead) -> four 16-byte chunks per fp16 plane.  ch-major thread order keeps a quarter-warp on eight
    //      consecutive rows, whose SWIZZLE_128B images of one chunk column hit eight different bank groups.
    const int ch = tid >= G::ROWS ? 1 : 0, m4 = tid - ch * G::ROWS;
    const bool has_item = tid < 2 * G::ROWS;
    float4 v[8];
    auto load_frames = [&](unsigned bb, unsigned cc) {
      const long n = (long)cc * kH4Frames + nbase + 4 * m4;
      const float* sp = p.s + ((size_t)bb * 16 + 8 * ch) * p.F + n;
      const bool ok = bb < n_rows && has_item && n >= 0 && n + 3 < p.F;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) v[kk] = ok ? ptx::ldg128_na(reinterpret_cast<const float4*>(sp + (size_t)kk * p.F)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // D (TMEM) -> out.  Thread (row i, half hb) drains output frames 4 i + 2 hb, + 1 = 32 consecutive samples = four 32-byte
    // chunks, but rows are 256 B apart: stored like that, every warp store would touch 32 lines.  A 4 x 4 chunk transpose
    // inside each lane quad (two shuffle stages) leaves lane r with chunk r & 3 of the quad's four rows, so one STG.256
    // covers eight whole 128-byte lines.
    auto epilogue = [&](unsigned bb, unsigned cc, int dbuf) {
      const int i = tid & 127, hb = tid >> 7, lane = tid & 31;
      const float scale = 1.0f / (float)(1 << kH16ScaleLog2);
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(dbuf * 128 + 2 * hb * 16);
      uint32_t r0[2][16], r1[2][16];
      ptx::tmem_ld16(taddr, r0[0]);
      ptx::tmem_ld16(taddr + 64, r1[0]);
      ptx::tmem_ld16(taddr + 16, r0[1]);
      ptx::tmem_ld16(taddr + 80, r1[1]);
      ptx::tmem_ld_wait();
      float val[4][8];   // chunk q = 2 dd + c8
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) val[q][e] = (__uint_as_float(r0[q >> 1][8 * (q & 1) + e]) + __uint_as_float(r1[q >> 1][8 * (q & 1) + e])) * scale;
      const bool b0 = lane & 1, b1 = lane & 2;
#pragma unroll
      for (int pr = 0; pr < 2; ++pr)   // lanes r, r ^ 1 swap the off-diagonal chunks of (2 pr, 2 pr + 1)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float recv = __shfl_xor_sync(0xffffffffu, b0 ? val[2 * pr][e] : val[2 * pr + 1][e], 1);
          val[2 * pr + 1][e] = b0 ? val[2 * pr + 1][e] : recv;
          val[2 * pr][e] = b0 ? recv : val[2 * pr][e];
        }
#pragma unroll
      for (int u = 0; u < 2; ++u)      // lanes r, r ^ 2 swap the off-diagonal chunks of (u, u + 2)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float recv = __shfl_xor_sync(0xffffffffu, b1 ? val[u][e] : val[u + 2][e], 2);
          val[u + 2][e] = b1 ? val[u + 2][e] : recv;
          val[u][e] = b1 ? recv : val[u][e];
        }
      // slot q now holds chunk (lane & 3) of row (i & ~3) | q
      const int cq = lane & 3;
      const long f0 = (long)cc * kH4Frames + 4 * (i & ~3) + 2 * hb + (cq >> 1);
      float* op = p.out + ((size_t)bb * p.F + f0) * 16 + 8 * (cq & 1);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (f0 + 4 * q < p.F) ptx::stg256_cs(op + (size_t)q * 64, val[q]);
    };

    load_frames(b, c);
    unsigned prev_b = 0, prev_c = 0;
    // sigma(k, n): odd bands (odd kk) flip on even global frames; quads start on multiples of 4, so the parity is j's
    const uint32_t flip_even = (p.parity & 1) ? 0u : 0x80000000u, flip_odd = flip_even ^ 0x80000000u;
    for (unsigned it = 0; it < n_iter; ++it) {
      const int pb = (int)(it & 1);
      H4_STAMP(0);
      if (has_item) {
        unsigned char* p1 = planes + (2 * pb) * G::PLANE;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t fl = (j & 1) ? flip_odd : flip_even;
          float w[8];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const float t = j == 0 ? v[kk].x : j == 1 ? v[kk].y : j == 2 ? v[kk].z : v[kk].w;
            w[kk] = (kk & 1) ? __uint_as_float(__float_as_uint(t) ^ fl) : t;
          }
          uint4 h1, h2;
          split2_f16(w[0], w[1], h1.x, h2.x);
          split2_f16(w[2], w[3], h1.y, h2.y);
          split2_f16(w[4], w[5], h1.z, h2.z);
          split2_f16(w[6], w[7], h1.w, h2.w);
          const uint32_t o = sw128_offset((uint32_t)m4 * 128u + 32u * j + 16u * ch);
          *reinterpret_cast<uint4*>(p1 + o) = h1;
          *reinterpret_cast<uint4*>(p1 + G::PLANE + o) = h2;
        }
      }
      H4_STAMP(1);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) {
        if (it == 0) ptx::mbar_wait(bankfull, 0);
        if constexpr (PAIR) ptx::mbar_arrive_cluster(pfull_leader + 8u * pb);
        else ptx::mbar_arrive(&pfull[pb]);
      }
      unsigned nb = b + step_b, nc = c + step_c;
      if (nc >= tpr) {
        nc -= tpr;
        ++nb;
      }
      if (it + 1 < n_iter) load_frames(nb, nc);
      H4_STAMP(2);
      H4_STAMP(3);
      if (it > 0) {
        // every worker waits (planes[pb ^ 1] are rewritten next iteration); the ninth warp has no TMEM rows to drain
        ptx::mbar_wait(&mma_bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
        ptx::tc_fence_after();
        H4_STAMP(4);
        if (warp < 8 && prev_b < n_rows) epilogue(prev_b, prev_c, (int)((it - 1) & 1));
      }
      H4_STAMP(5);
      prev_b = b;
      prev_c = c;
      b = nb;
      c = nc;
    }
    ptx::mbar_wait(&mma_bar[(n_iter - 1) & 1], ((n_iter - 1) >> 1) & 1);
    ptx::tc_fence_after();
    if (warp < 8 && prev_b < n_rows) epilogue(prev_b, prev_c, (int)((n_iter - 1) & 1));
  }  // workers
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    ptx::cluster_sync_all();   // the peer may still be read by / signalled from the leader's last MMAs
    if (warp == 0) ptx::tmem_dealloc_pair(tmem, 256);
  } else {
    if (warp == 0) ptx::tmem_dealloc(tmem, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
template <typename Kern>
inline int h4_configure(Kern kern, int bytes, bool (&configured)[64], int& sm_count_out) {
  static int sm_count[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sm_count[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sm_count[dev] = n > 0 ? n : 148;
  }
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    configured[dev] = true;
  }
  sm_count_out = sm_count[dev];
  return 0;
}

// one CTA (or CTA pair) per SM, static round-robin over the tiles; PAIR launches clusters of two
template <bool PAIR, typename Kern, typename Params>
inline int h4_launch(Kern kern, Params p, int B, int threads, int bytes, bool (&configured)[64], cudaStream_t st) {
  int sms = 0;
  if (int e = h4_configure(kern, bytes, configured, sms)) return e;
  p.tiles_per_row = (p.F + kH4Frames - 1) / kH4Frames;
  p.n_tiles = p.tiles_per_row * B;
  if (p.tiles_per_row >= (1L << 31) || p.n_tiles >= (1L << 40)) return -2;
  long grid = PAIR ? (sms & ~1) : sms;
  const long want = PAIR ? ((p.n_tiles + 1) & ~1L) : p.n_tiles;
  if (grid > want) grid = want;
  if constexpr (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, p);
  } else {
    kern<<<(unsigned)grid, threads, bytes, st>>>(p);
    return (int)cudaGetLastError();
  }
}

template <int JLO, int KT, bool PAIR = false>
int h4_launch_analysis(H4AnalysisParams p, int B, cudaStream_t st) {
  static bool configured[64] = {false};
  return h4_launch<PAIR>(h4_analysis_kernel<JLO, KT, PAIR>, p, B, kH4Threads, H4Geometry<KT, PAIR>::BYTES, configured, st);
}

template <int JLO, int KT, bool PAIR = false>
int h4_launch_synthesis(H4SynthesisParams p, int B, cudaStream_t st) {
  static bool configured[64] = {false};
  return h4_launch<PAIR>(h4_synthesis_kernel<JLO, KT, PAIR>, p, B, kH4SynThreads, H4Geometry<KT, PAIR>::BYTES, configured, st);
}

// ---------------------------------------------------------------------------------------------
// host: bank images, fp16 bits in UMMA K-major no-swizzle layout [chunk of 8 K][128 rows][8],
// rows = part * 64 + delta * 16 + (band | phase), part 0 = c1, part 1 = c2
// ---------------------------------------------------------------------------------------------
inline void hankel4_build_banks(const float* hk /*[16][512]*/, int jlo, int kt, uint16_t* img_analysis, uint16_t* img_synthesis) {
  auto bits = [](float v) {
    const __half h = __float2half_rn(v);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  };
  const int ks = kt / 16 + 3, kp = 16 * ks;
  const int dlo = jlo / 16, dhi = (jlo + kt) / 16 - 1;
  const float sa = (float)(1 << kH16ScaleLog2), ss = 16.f * sa;
  for (int kc = 0; kc < kp / 8; ++kc)
    for (int row = 0; row < 128; ++row)
      for (int e = 0; e < 8; ++e) {
        const int kap = 8 * kc + e;
        const int part = row / 64, delta = (row % 64) / 16, q = row % 16;
        const size_t at = ((size_t)kc * 128 + row) * 8 + e;
        {  // analysis: K index = tap offset within the (delta-shifted) window, q = band
          const int j = kap - 16 * delta;
          const float v = (j >= 0 && j < kt) ? sa * hk[q * 512 + jlo + j] : 0.f;
          const float c1 = __half2float(__float2half_rn(v));
          img_analysis[at] = part == 0 ? bits(c1) : bits(v - c1);
        }
        {  // synthesis: K index = (step s, band kb): lag e = dhi - s, tap 16 (e + delta) + q, q = output phase
          const int s2 = kap / 16, kb = kap % 16, d = dhi - s2 + delta;
          const float v = (d >= dlo && d <= dhi) ? ss * hk[kb * 512 + 16 * d + q] : 0.f;
          const float c1 = __half2float(__float2half_rn(v));
          img_synthesis[at] = part == 0 ? bits(c1) : bits(v - c1);
        }
      }
}

// per-rank images for the CTA-pair kernels: [rank][2 KS][96 rows][8]; rows 0-63 = rows 64 rank .. of the single-CTA image (rank 0:
// c1, rank 1: c2), rows 64-95 = c1 rows 32 rank .. 32 rank + 31 (this rank's half of the N = 64 operand)
inline void hankel4_pair_image(const uint16_t* single /*[2 KS][128][8]*/, int kt, uint16_t* pair /*[2][2 KS][96][8]*/) {
  const int chunks = 2 * (kt / 16 + 3);
  for (int r = 0; r < 2; ++r)
    for (int kc = 0; kc < chunks; ++kc)
      for (int row = 0; row < 96; ++row) {
        const int src = row < 64 ? 64 * r + row : 32 * r + (row - 64);
        memcpy(pair + (((size_t)r * chunks + kc) * 96 + row) * 8, single + ((size_t)kc * 128 + src) * 8, 16);
      }
}

// Largest number of edge K-steps (per side, <= 7) whose correction terms may be dropped: the dropped terms are bounded by
// 2 * 2^-11 * max|input| * (sum of the |bank| entries they multiply); returns the largest trim whose bound stays <= budget.
inline int hankel4_pick_trim(const float* hk /*[16][512]*/, int jlo, int kt, bool synthesis, double budget) {
  const int ks = kt / 16 + 3, dlo = jlo / 16, dhi = (jlo + kt) / 16 - 1;
  int best = 0;
  for (int trim = 1; trim <= 7 && ks - 2 * trim >= 1; ++trim) {
    double worst = 0.0;
    for (int delta = 0; delta < 4; ++delta)
      for (int q = 0; q < 16; ++q) {   // q = band (analysis) or output phase (synthesis)
        double sum = 0.0;
        for (int side = 0; side < 2; ++side)
          for (int t = 0; t < trim; ++t) {
            const int s = side ? ks - 1 - t : t;
            if (!synthesis) {
              for (int e = 0; e < 16; ++e) {
                const int j = 16 * s + e - 16 * delta;
                if (j >= 0 && j < kt) sum += fabs((double)hk[q * 512 + jlo + j]);
              }
            } else {
              const int d = dhi - s + delta;
              if (d >= dlo && d <= dhi)
                for (int k = 0; k < 16; ++k) sum += 16.0 * fabs((double)hk[k * 512 + 16 * d + q]);
            }
          }
        worst = sum > worst ? sum : worst;
      }
    if (2.0 * worst / 2048.0 <= budget) best = trim;
    else break;
  }
  return best;
}

}  // namespace pqmf
