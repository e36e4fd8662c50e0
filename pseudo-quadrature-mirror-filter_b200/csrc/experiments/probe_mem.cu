// Data-movement ceilings for the analysis / synthesis access patterns (no arithmetic):
//   A: TMA bulk load of a contiguous 9728-byte window per tile -> smem -> 16 sub-band rows x 512 B coalesced stores
//   B: same loads, stores contiguous 8 KB per tile (is the 16-row scatter the problem?)
//   C: plain float4 copy (grid-stride), the MEASURED_PEAKS-style reference
#include <cuda_runtime.h>
#include <cstdio>
#include "../ptx.cuh"
using namespace pqmf::ptx;

constexpr int XS = 2432;
template <int MODE, int NBUF>
__global__ void __launch_bounds__(128, 4) pattern_kernel(const float* __restrict__ x, float* __restrict__ y, long T, long F, long tiles_per_row, long n_tiles) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* xs = reinterpret_cast<float*>(sm);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + NBUF * XS * 4);
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < NBUF; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  auto stage = [&](long tile, int buf) {
    const long b = tile / tiles_per_row, c = tile % tiles_per_row;
    long s0 = c * 2048 - 192; if (s0 < 0) s0 = 0; if (s0 + XS > T) s0 = T - XS;
    mbar_arrive_expect_tx(&full[buf], XS * 4);
    bulk_g2s(xs + buf * XS, x + b * T + s0, XS * 4, &full[buf]);
  };
  if (tid == 0) for (int i = 0; i < NBUF; ++i) { long t = blockIdx.x + (long)i * gridDim.x; if (t < n_tiles) stage(t, i); }
  unsigned it = 0;
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int buf = it % NBUF;
    mbar_wait(&full[buf], (it / NBUF) & 1);
    const long b = tile / tiles_per_row, c = tile % tiles_per_row;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = xs[buf * XS + 192 + k * 128 + tid];
    __syncthreads();
    if (tid == 0) { long nt = tile + (long)NBUF * gridDim.x; if (nt < n_tiles) stage(nt, buf); }
    if (MODE == 0) {
      float* yp = y + b * 16 * F + c * 128 + tid;
#pragma unroll
      for (int k = 0; k < 16; ++k) __stcs(yp + k * F, v[k]);
    } else {
      float* yp = y + b * T + c * 2048 + tid;
#pragma unroll
      for (int k = 0; k < 16; ++k) __stcs(yp + k * 128, v[k]);
    }
  }
}
__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}
template <typename L> float timeit(L launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  float best = 1e9;
  for (int i = 0; i < 10; ++i) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  const long B = 64, T = 1 << 20, F = T / 16, tpr = F / 128, n_tiles = B * tpr;
  float *x, *y; cudaMalloc(&x, B * T * 4); cudaMalloc(&y, B * T * 4); cudaMemset(x, 0, B * T * 4);
  const double bytes = 2.0 * B * T * 4;
  auto run = [&](auto kern, int nbuf, int ctas, const char* name) {
    const size_t smem = nbuf * XS * 4 + 64;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float ms = timeit([&] { kern<<<148 * ctas, 128, smem>>>(x, y, T, F, tpr, n_tiles); });
    printf("%-52s %7.3f ms  %7.0f GB/s  [%s]\n", name, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
  };
  run(pattern_kernel<0, 2>, 2, 4, "A  TMA window in, 16 rows x 512 B out, 2 buf, 4 CTA/SM");
  run(pattern_kernel<0, 3>, 3, 4, "A  TMA window in, 16 rows x 512 B out, 3 buf, 4 CTA/SM");
  run(pattern_kernel<0, 3>, 3, 8, "A  TMA window in, 16 rows x 512 B out, 3 buf, 8 CTA/SM");
  run(pattern_kernel<1, 3>, 3, 4, "B  TMA window in, contiguous 8 KB out,  3 buf, 4 CTA/SM");
  run(pattern_kernel<1, 3>, 3, 8, "B  TMA window in, contiguous 8 KB out,  3 buf, 8 CTA/SM");
  float ms = timeit([&] { copy_kernel<<<148 * 16, 256>>>((const float4*)x, (float4*)y, B * T / 4); });
  printf("%-52s %7.3f ms  %7.0f GB/s\n", "C  float4 grid-stride copy", ms, bytes / ms * 1e-6);
  return 0;
}
