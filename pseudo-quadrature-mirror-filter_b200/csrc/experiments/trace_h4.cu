// Timeline of the Hankel-4 analysis kernel: clock64 stamps of every warp of CTA 0 at the phase boundaries of each tile
// (hankel4.cuh, H4_STAMP) -> where does a tile's ~5700 cycles go?
#define PQMF_H4_TRACE 1
#include <cstdio>
#include <vector>
#include "../hankel4.cuh"
using namespace pqmf;
int main(int argc, char** argv) {
  const bool pair = argc > 4 && atoi(argv[4]) != 0;
  const int B = 64; const long T = 1 << 20, F = T / 16;
  float *x, *y; uint16_t* bank; long long* tr;
  cudaMalloc(&x, (size_t)B * T * 4); cudaMalloc(&y, (size_t)B * T * 4); cudaMalloc(&bank, 2 * 27 * 4096); cudaMalloc(&tr, (64 * 64 + 512) * 8);
  cudaMemset(x, 0, (size_t)B * T * 4); cudaMemset(bank, 0, 2 * 27 * 4096); cudaMemset(tr, 0, 64 * 64 * 8);
  if (argc > 1) {   // random signal and bank instead of zeros (does the data change the timing?)
    std::vector<float> hx(1 << 22), hk(16 * 512);
    srand(3);
    for (auto& v : hx) v = (float)rand() / RAND_MAX - 0.5f;
    for (auto& v : hk) v = ((float)rand() / RAND_MAX - 0.5f) * 0.05f;
    for (size_t o = 0; o < (size_t)B * T; o += hx.size()) cudaMemcpy(x + o, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    std::vector<uint16_t> ia(27 * 2048), is(27 * 2048);
    hankel4_build_banks(hk.data(), 16, 512, 64, 384, ia.data(), is.data());
    cudaMemcpy(bank, ia.data(), ia.size() * 2, cudaMemcpyHostToDevice);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  H4AnalysisParams p{};
  p.x = x; p.y = y; p.bank = bank; p.T = T; p.F = F; p.off = 256; p.parity = 0; p.trace = tr; p.trim_lo = p.trim_hi = argc > 3 ? atoi(argv[3]) : 0; p.g = h4_shape(16, 64, 384, pair, false);
  for (int rep = 0; rep < 3; ++rep) {
    int rc = (pair ? h4_launch_analysis<16, true>(p, B, 0) : h4_launch_analysis<16, false>(p, B, 0));
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e) { printf("launch rc=%d cuda=%s\n", rc, cudaGetErrorString(e)); return 1; }
  }
  const bool synth = argc > 2 && argv[2][0] == 's';   // trace_h4 <random> <a|s> <trim> <pair>
  H4SynthesisParams q{};
  q.s = x; q.out = y; q.bank = bank; q.F = F; q.o = 16; q.parity = 0; q.trace = tr; q.trim_lo = q.trim_hi = argc > 3 ? atoi(argv[3]) : 0; q.g = h4_shape(16, 64, 384, pair, true);
  if (synth) {
    cudaMemset(tr, 0, 64 * 64 * 8);
    (pair ? h4_launch_synthesis<16, true>(q, B, 0) : h4_launch_synthesis<16, false>(q, B, 0));
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("synthesis cuda=%s\n", cudaGetErrorString(e)); return 1; }
  }
  cudaEventRecord(e0);
  for (int rep = 0; rep < 20; ++rep) { if (synth) (pair ? h4_launch_synthesis<16, true>(q, B, 0) : h4_launch_synthesis<16, false>(q, B, 0)); else (pair ? h4_launch_analysis<16, true>(p, B, 0) : h4_launch_analysis<16, false>(p, B, 0)); }
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  printf("%s %s data: %.1f us per launch (20 back to back)\n", synth ? "synthesis" : "analysis", argc > 1 ? "random" : "zero", ms * 50.f);
  std::vector<long long> h(64 * 64 + 512);
  cudaMemcpy(h.data(), tr, h.size() * 8, cudaMemcpyDeviceToHost);
  {
    long long cmin = 1LL << 60, cmax = 0, nmin = 1LL << 60, nmax = 0;
    for (int i = 0; i < 148; ++i) {
      const long long c = h[64 * 64 + 2 * i], n = h[64 * 64 + 2 * i + 1];
      cmin = c < cmin ? c : cmin; cmax = c > cmax ? c : cmax; nmin = n < nmin ? n : nmin; nmax = n > nmax ? n : nmax;
    }
    printf("per-CTA lifetime: %lld..%lld cycles, %lld..%lld ns -> SM clock %.0f MHz (CTA 0: %lld cycles / %lld ns)\n", cmin, cmax, nmin, nmax,
           1e3 * (double)cmax / (double)nmax, h[64 * 64], h[64 * 64 + 1]);
  }
  const long long t00 = h[(0 * 8 + 1) * 8 + 0];
  printf("warp 1 (MMA warp) and warp 5; cycles relative to the first stamp\n");
  printf("it |  start   conv   loads+publish  wait+drain(it-lag) | period\n");
  long long prev = t00;
  for (int it = 0; it < 56; ++it) {
    for (int w : {1, 5}) {
      long long* r = &h[((size_t)it * 8 + w) * 8];
      if (!r[0]) continue;
      printf("%2d w%d %8lld %6lld %6lld %10lld", it, w, r[0] - t00, r[1] - r[0], r[2] - r[1], r[5] - r[2]);
      if (w == 1) { printf(" | %lld", r[0] - prev); prev = r[0]; }
      printf("\n");
    }
    if (it == 12) it = 40;
  }
  return 0;
}
