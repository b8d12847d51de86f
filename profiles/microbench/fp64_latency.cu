// Dependent-issue latencies that bound the 128 x 128 Cholesky leaf: DFMA chain, 64-bit shuffle, MUFU-seeded
// rsqrt (csrc/fastmath.cuh), shared-memory round trip, __syncthreads with 256 threads.  One warp / one CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../additivecausalexpansion_b200/csrc fp64_latency.cu
#include <cstdio>
#include "fastmath.cuh"
using namespace ace;

__global__ void lat(double* out, long long* cyc, double seed) {
  __shared__ double sm[256];
  const int t = threadIdx.x;
  double x = seed + t * 1e-9, y = 1.0000001;
  long long c0, c1;
  const int N = 1024;
  // 1. dependent DFMA chain
  c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-12);
  c1 = clock64();
  if (t == 0) cyc[0] = (c1 - c0);
  // 2. dependent 64-bit shuffle chain
  c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (t + 1) & 31);
  c1 = clock64();
  if (t == 0) cyc[1] = (c1 - c0);
  // 3. dependent fast_rsqrt chain
  x = fabs(x) + 1.0;
  c0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = fast_rsqrt(x) + 1.0;
  c1 = clock64();
  if (t == 0) cyc[2] = (c1 - c0);
  // 4. smem store -> __syncthreads -> load round trip (256 threads)
  c0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    sm[t] = x;
    __syncthreads();
    x = sm[(t + 1) & 255] + 1e-9;
    __syncthreads();
  }
  c1 = clock64();
  if (t == 0) cyc[3] = (c1 - c0);
  // 5. smem store -> __syncwarp -> load (warp-local)
  c0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    sm[t] = x;
    __syncwarp();
    x = sm[(t & ~31) + ((t + 1) & 31)] + 1e-9;
    __syncwarp();
  }
  c1 = clock64();
  if (t == 0) cyc[4] = (c1 - c0);
  // 6. DFMA throughput: 8 independent chains per thread, 256 threads (8 warps, 2 per SMSP)
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  c0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    a0 = fma(a0, y, 1e-12); a1 = fma(a1, y, 1e-12); a2 = fma(a2, y, 1e-12); a3 = fma(a3, y, 1e-12);
    a4 = fma(a4, y, 1e-12); a5 = fma(a5, y, 1e-12); a6 = fma(a6, y, 1e-12); a7 = fma(a7, y, 1e-12);
  }
  c1 = clock64();
  if (t == 0) cyc[5] = (c1 - c0);
  out[t] = x + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8 * 8);
  for (int rep = 0; rep < 2; ++rep) lat<<<1, 256>>>(out, cyc, 1.5);
  long long h[8];
  cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  printf("{\"dfma_dependent_cycles\": %.1f, \"shfl64_dependent_cycles\": %.1f, \"fast_rsqrt_plus_add_cycles\": %.1f, "
         "\"sts_bar_lds_bar_256thr_cycles\": %.1f, \"sts_syncwarp_lds_cycles\": %.1f, \"dfma_8chains_8warps_cycles_per_iter\": %.1f}\n",
         h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0, h[5] / 1024.0);
  return 0;
}
