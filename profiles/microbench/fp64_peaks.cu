// FP64 peak calibrator for B200 (sm_100a): DMMA.8x8x4 and DFMA issue-rate microbenchmarks.
// MEASURED_PEAKS.json carries no FP64 figure, so every FP64 roofline fraction in this repo is
// quoted against the numbers this program prints (copied to profiles/fp64_peaks_r01.json).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(1024) dmma_tput(double* out, int iters) {
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(1024) dfma_tput(double* out, int iters) {
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  if (s == 123.456) out[0] = s;
}

// mixed: does DFMA issue alongside DMMA (separate pipes?) or share the FP64 datapath
template <int NF>
__global__ void __launch_bounds__(1024) mixed_tput(double* out, int iters) {
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[8][2];
  double f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      if (i < NF) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, 8));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.clockRate);
  const int iters = 20000;
  printf(" \"dmma\": [\n");
  int tpbs[] = {128, 256, 512, 1024};
  bool first = true;
  for (int tpb : tpbs) {
    for (int ilp : {1, 2, 4, 8, 16}) {
      float ms = 0;
      auto run = [&](auto kern) { ms = time_ms([&] { kern<<<sms, tpb>>>(out, iters); }); };
      if (ilp == 1) run(dmma_tput<1>);
      if (ilp == 2) run(dmma_tput<2>);
      if (ilp == 4) run(dmma_tput<4>);
      if (ilp == 8) run(dmma_tput<8>);
      if (ilp == 16) run(dmma_tput<16>);
      double flops = 2.0 * 256 * (double)ilp * iters * (tpb / 32) * sms;
      printf("%s  {\"warps_per_sm\": %d, \"ilp\": %d, \"ms\": %.4f, \"tflops\": %.3f}", first ? "" : ",\n",
             tpb / 32, ilp, ms, flops / ms * 1e-9);
      first = false;
    }
  }
  printf("\n ],\n \"dfma\": [\n");
  first = true;
  for (int tpb : tpbs) {
    for (int ilp : {1, 4, 8}) {
      float ms = 0;
      auto run = [&](auto kern) { ms = time_ms([&] { kern<<<sms, tpb>>>(out, iters); }); };
      if (ilp == 1) run(dfma_tput<1>);
      if (ilp == 4) run(dfma_tput<4>);
      if (ilp == 8) run(dfma_tput<8>);
      double flops = 2.0 * (double)ilp * iters * tpb * sms;
      printf("%s  {\"warps_per_sm\": %d, \"ilp\": %d, \"ms\": %.4f, \"tflops\": %.3f}", first ? "" : ",\n",
             tpb / 32, ilp, ms, flops / ms * 1e-9);
      first = false;
    }
  }
  printf("\n ],\n \"mixed_dmma8_plus_dfma\": [\n");
  first = true;
  for (int nf : {0, 2, 4, 8}) {
    float ms = 0;
    const int tpb = 256;
    auto run = [&](auto kern) { ms = time_ms([&] { kern<<<sms, tpb>>>(out, iters); }); };
    if (nf == 0) run(mixed_tput<0>);
    if (nf == 2) run(mixed_tput<2>);
    if (nf == 4) run(mixed_tput<4>);
    if (nf == 8) run(mixed_tput<8>);
    double fl_mma = 2.0 * 256 * 8.0 * iters * (tpb / 32) * sms;
    double fl_fma = 2.0 * nf * (double)iters * tpb * sms;
    printf("%s  {\"dfma_per_8dmma\": %d, \"ms\": %.4f, \"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f}",
           first ? "" : ",\n", nf, ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    first = false;
  }
  // sustained: run the best DMMA config for ~3 s to see the power-capped clock
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int reps = 60;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) dmma_tput<8><<<sms, 256>>>(out, iters * 10);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 256 * 8.0 * iters * 10.0 * 8 * sms * reps;
    printf("\n ],\n \"dmma_sustained\": {\"seconds\": %.3f, \"tflops\": %.3f}\n}\n", ms * 1e-3, flops / ms * 1e-9);
  }
  return 0;
}
