// Validates the hand-built tcgen05 descriptors used by fast16.cuh before the production kernel depends on them:
// D[128 x N] = A[128 x K] * B[N x K]^T in kind::tf32, A/B K-major no-swizzle in shared memory with a padded
// leading byte offset, D in TMEM, read back with tcgen05.ld.  Case 1 (analysis shape): K = 32, N = 16.
// Case 2 (synthesis shape): K = 16, N = 32.  Inputs are tf32-exact small values -> the result must be exact.
// Case 3: 3xTF32 split of random fp32 data against float64.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../ptx.cuh"
using namespace pqmf::ptx;

constexpr int LBO_A = 2064, SBO = 128;

// A planes: [term][K/4 chunks][128 rows][4 floats] with chunk stride LBO_A; B planes: chunk stride N*16
template <int K, int N, int TERMS>
__global__ void __launch_bounds__(128) umma_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  // A: [TERMS][128][K] row-major global; B: [TERMS][N][K]; D: [128][N]
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int KC = K / 4;
  constexpr int LBO_B = N * 16;
  constexpr int APL = KC * LBO_A, BPL = KC * LBO_B;
  unsigned char* sa = smem;
  unsigned char* sb = smem + TERMS * APL;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int t = 0; t < TERMS; ++t) {
    for (int e = tid; e < 128 * K; e += 128) {
      const int row = e / K, k = e % K;
      *reinterpret_cast<float*>(sa + t * APL + (k / 4) * LBO_A + row * 16 + (k % 4) * 4) = A[(t * 128 + row) * K + k];
    }
    for (int e = tid; e < N * K; e += 128) {
      const int row = e / K, k = e % K;
      *reinterpret_cast<float*>(sb + t * BPL + (k / 4) * LBO_B + row * 16 + (k % 4) * 4) = B[(t * N + row) * K + k];
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, N);
    bool acc = false;
    for (int t = 0; t < TERMS; ++t)
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = umma_desc(smem_u32(sa + t * APL + ks * 2 * LBO_A), LBO_A, SBO);
        const uint64_t db = umma_desc(smem_u32(sb + t * BPL + ks * 2 * LBO_B), LBO_B, SBO);
        umma_tf32(tm, da, db, idesc, acc);
        acc = true;
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
  if (N == 16) {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[tid * N + j] = __uint_as_float(r[j]);
  } else {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * N + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 32);
}

static float tf32_trunc(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; memcpy(&v, &u, 4); return v; }
static float tf32_rn(float v) { uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xffffe000u; memcpy(&v, &u, 4); return v; }

template <int K, int N, int TERMS>
int run_case(const char* name, const std::vector<float>& A, const std::vector<float>& B, const std::vector<double>& ref, double tol) {
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 128 * N * 4);
  const size_t smem = TERMS * ((K / 4) * LBO_A + (K / 4) * N * 16) + 256;
  auto kern = umma_kernel<K, N, TERMS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(128 * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < 128 * N; ++i) {
    const double err = std::fabs((double)D[i] - ref[i]);
    if (!(err <= tol)) { if (bad < 5) printf("  mismatch [%d,%d]: got %g want %g\n", i / N, i % N, D[i], ref[i]); ++bad; }
    if (err > maxerr || err != err) maxerr = err;
  }
  printf("%-34s cuda=%s  max|err|=%.3e  mismatches=%d  -> %s\n", name, cudaGetErrorString(e), maxerr, bad, (bad == 0 && e == cudaSuccess) ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return (bad == 0 && e == cudaSuccess) ? 0 : 1;
}

int main() {
  int fails = 0;
  srand(1);
  {  // case 1: K=32, N=16, exact small values
    constexpr int K = 32, N = 16;
    std::vector<float> A(128 * K), B(N * K);
    for (auto& v : A) v = (float)((rand() % 17) - 8) * 0.125f;
    for (auto& v : B) v = (float)((rand() % 13) - 6) * 0.25f;
    std::vector<double> ref(128 * N, 0.0);
    for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) ref[i * N + j] += (double)A[i * K + k] * B[j * K + k];
    fails += run_case<K, N, 1>("analysis shape 128x16x32 exact", A, B, ref, 0.0);
  }
  {  // case 2: K=16, N=32
    constexpr int K = 16, N = 32;
    std::vector<float> A(128 * K), B(N * K);
    for (auto& v : A) v = (float)((rand() % 17) - 8) * 0.125f;
    for (auto& v : B) v = (float)((rand() % 13) - 6) * 0.25f;
    std::vector<double> ref(128 * N, 0.0);
    for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) ref[i * N + j] += (double)A[i * K + k] * B[j * K + k];
    fails += run_case<K, N, 1>("synthesis shape 128x32x16 exact", A, B, ref, 0.0);
  }
  {  // case 3: 3xTF32 on random fp32 data: terms (Ahi,Bhi), (Alo,Bhi), (Ahi,Blo)
    constexpr int K = 32, N = 16;
    std::vector<float> a(128 * K), b(N * K);
    for (auto& v : a) v = 0.1f * ((float)rand() / RAND_MAX - 0.5f);
    for (auto& v : b) v = 4.0f * ((float)rand() / RAND_MAX - 0.5f);
    std::vector<float> A(3 * 128 * K), B(3 * N * K);
    for (int i = 0; i < 128 * K; ++i) { float hi = tf32_rn(a[i]); float lo = a[i] - hi; A[i] = hi; A[128 * K + i] = lo; A[2 * 128 * K + i] = hi; }
    for (int i = 0; i < N * K; ++i) { float hi = tf32_rn(b[i]); float lo = tf32_rn(b[i] - hi); B[i] = hi; B[N * K + i] = hi; B[2 * N * K + i] = lo; }
    std::vector<double> ref(128 * N, 0.0);
    for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) ref[i * N + j] += (double)a[i * K + k] * b[j * K + k];
    fails += run_case<K, N, 3>("3xTF32 split vs float64 (tol 2e-7)", A, B, ref, 2e-7);
    // for scale: single-pass tf32 error on the same data
    std::vector<float> A1(a), B1(b);
    fails += 0 * run_case<K, N, 1>("1xTF32 (informational)", A1, B1, ref, 1e30);
  }
  printf(fails ? "UMMA PROBE FAILED\n" : "UMMA PROBE OK\n");
  return fails;
}
