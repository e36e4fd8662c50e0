// int16 PCM edge of the PQMF path: de-interleave + int16 -> fp32 fused into the analysis loads, fp32 -> int16 +
// interleave fused into the synthesis stores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqmf {

// ---- int16 PCM edge.  WAV frames are interleaved [clip][time][channel]; torchaudio.load turns them into
//      float32 / 32768 per channel (what every wrapper of the reference feeds the PQMF: PQMFWrapper.py:113,
//      1-PitchShifterWrapper.py:348) and PitchShifterPvoc/2-TestBlocks.py:26-30 down-mixes with mean(dim=0). ----
struct PcmIn {
  const int16_t* pcm;  // [clips, T, C] or nullptr (fp32 rows)
  int C;               // channels per WAV frame
  int downmix;         // 1: one row per clip = mean over the channels (fp32 sum in channel order, then / C, as torch.mean does)
};
// sample s of row b (row = clip * C + channel, or clip when down-mixing); T = frames per clip
__device__ __forceinline__ float pcm_sample(const PcmIn& in, long b, long s, long T) {
  constexpr float kInv = 1.0f / 32768.0f;
  if (in.downmix) {
    const int16_t* f = in.pcm + ((size_t)b * T + s) * in.C;
    float acc = 0.f;
    for (int c = 0; c < in.C; ++c) acc += (float)f[c] * kInv;
    return acc / (float)in.C;
  }
  const long clip = b / in.C;
  const int ch = (int)(b - clip * in.C);
  return (float)in.pcm[((size_t)clip * T + s) * in.C + ch] * kInv;
}
// fp32 -> PCM: round to nearest even, saturate (torch: clamp(round(v * 32768), -32768, 32767).to(int16))
__device__ __forceinline__ int16_t pcm_quantise(float v) {
  const float q = rintf(v * 32768.0f);
  return (int16_t)(q < -32768.f ? -32768.f : (q > 32767.f ? 32767.f : q));
}

// eight consecutive samples s .. s + 7 of row b (s a multiple of 8, T a multiple of 8): one 16-byte load for mono, one 32-byte load
// for stereo (either channel or the down-mix), element-wise otherwise
__device__ __forceinline__ void pcm_load8(const PcmIn& in, long b, long s, long T, float (&v)[8]) {
  constexpr float kInv = 1.0f / 32768.0f;
  if (in.C == 1) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(in.pcm + (size_t)b * T + s));
    const int w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = (float)(int16_t)(w[i] & 0xFFFF) * kInv;
      v[2 * i + 1] = (float)(int16_t)(w[i] >> 16) * kInv;
    }
  } else if (in.C == 2) {
    const long clip = in.downmix ? b : (b >> 1);
    const int ch = in.downmix ? 0 : (int)(b & 1);
    const int4* src = reinterpret_cast<const int4*>(in.pcm + ((size_t)clip * T + s) * 2);
    const int4 r0 = __ldg(src), r1 = __ldg(src + 1);
    const int w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};  // one WAV frame (left, right) per word
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float l = (float)(int16_t)(w[i] & 0xFFFF) * kInv, r = (float)(int16_t)(w[i] >> 16) * kInv;
      v[i] = in.downmix ? (l + r) / 2.0f : (ch ? r : l);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = pcm_sample(in, b, s + i, T);
  }
}

// eight consecutive output samples t .. t + 7 (t a multiple of 8) of row b -> interleaved int16 PCM [clips, total, C]
__device__ __forceinline__ void pcm_store8(int16_t* pcm_out, int C, long b, long t, long total, const float (&v)[8]) {
  if (C == 1) {
    int4 w;
    int* wp = reinterpret_cast<int*>(&w);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      wp[i] = (int)(uint16_t)pcm_quantise(v[2 * i]) | ((int)(uint16_t)pcm_quantise(v[2 * i + 1]) << 16);
    *reinterpret_cast<int4*>(pcm_out + (size_t)b * total + t) = w;
  } else {
    const long clip = b / C;
    const int ch = (int)(b - clip * C);
    int16_t* q = pcm_out + ((size_t)clip * total + t) * C + ch;
#pragma unroll
    for (int i = 0; i < 8; ++i) q[(size_t)i * C] = pcm_quantise(v[i]);
  }
}

}  // namespace pqmf
