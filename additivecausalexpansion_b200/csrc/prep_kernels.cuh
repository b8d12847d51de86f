// prep_kernels.cuh -- GPU-side preprocessing: normalize_train / normalize_test of
// src/utilities_cpp.cpp:13-118 with the data resident on the device.
//
// Parity target: BIT-exact against the host versions (csrc/host_utils.inl, themselves bit-exact against the compiled
// reference, tests/test_oracle_vs_reference.py).  What makes that possible:
//   * the column statistics the reference takes (unique count, min, max, median, max |.|) are order statistics: one
//     bitonic sort per column (one CTA per column, in a global scratch row) gives all of them, independent of any
//     summation order;
//   * every transform is an element-wise IEEE subtraction / division by a scalar (col_affine_kernel): the device
//     executes the reference's sequence of column operations literally, including its index quirks (the host keeps
//     the control flow, see ace_normalize_train_gpu);
//   * mean and standard deviation of y are order-DEPENDENT sums: y_moments_kernel replays Armadillo's two-accumulator
//     loops (arrayops::accumulate, op_var::direct_var) sequentially on one thread from a shared-memory copy, with
//     __dadd_rn / __dmul_rn so that nothing is contracted into an FMA.
// The spline bases stay on the host: arma::pow(x - k, 3) is glibc's pow, which no device routine reproduces bit for bit.
#pragma once
#include "common.cuh"

namespace ace {

constexpr int PREP_STATS = 8;  // per column: min, max, lower middle, upper middle, unique count, -, -, -

// One CTA per column: sort a copy (padded with +inf to a power of two) and take the order statistics.
__global__ void __launch_bounds__(1024) col_sort_stats_kernel(const double* __restrict__ cols, long col_stride, int n,
                                                              int npow2, double* __restrict__ scratch,
                                                              double* __restrict__ stats) {
  const double* c = cols + (size_t)blockIdx.x * col_stride;
  double* s = scratch + (size_t)blockIdx.x * npow2;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) s[i] = (i < n) ? c[i] : inf;
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = s[i], b = s[l];
          const bool up = (i & k) == 0;
          if (up ? (b < a) : (a < b)) {
            s[i] = b;
            s[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  __shared__ int nuniq;
  if (threadIdx.x == 0) nuniq = 1;
  __syncthreads();
  int loc = 0;
  for (int i = 1 + threadIdx.x; i < n; i += blockDim.x) loc += (s[i] != s[i - 1]) ? 1 : 0;
  if (loc) atomicAdd(&nuniq, loc);
  __syncthreads();
  if (threadIdx.x == 0) {
    double* o = stats + (size_t)blockIdx.x * PREP_STATS;
    o[0] = s[0];
    o[1] = s[n - 1];
    o[2] = s[(n - 1) / 2];  // n even: lower middle; n odd: the median
    o[3] = s[n / 2];        // n even: upper middle; n odd: the median
    o[4] = (double)nuniq;
  }
}

// mode 0: c = (c - a) / b;  1: c -= a;  2: c /= b;  3: c = 0
__global__ void __launch_bounds__(256) col_affine_kernel(double* __restrict__ c, int n, double a, double b, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = c[i];
  double r;
  if (mode == 0) r = __ddiv_rn(__dsub_rn(v, a), b);
  else if (mode == 1) r = __dsub_rn(v, a);
  else if (mode == 2) r = __ddiv_rn(v, b);
  else r = 0.0;
  c[i] = r;
}

// y <- (y - mean) / sd with Armadillo's summation order; out[0] = mean, out[1] = sd (n - 1 form).
// The vector is staged in shared memory when it fits (`in_smem`), thread 0 runs the sequential sums.
__global__ void __launch_bounds__(1024) y_moments_kernel(double* __restrict__ y, int n, int in_smem,
                                                         double* __restrict__ out) {
  extern __shared__ double ys[];
  double* v = in_smem ? ys : y;
  if (in_smem) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) ys[i] = y[i];
    __syncthreads();
  }
  __shared__ double s_mean, s_sd;
  if (threadIdx.x == 0) {
    auto accumulate = [&]() {  // arrayops::accumulate
      double acc1 = 0.0, acc2 = 0.0;
      int i, j;
      for (i = 0, j = 1; j < n; i += 2, j += 2) {
        acc1 = __dadd_rn(acc1, v[i]);
        acc2 = __dadd_rn(acc2, v[j]);
      }
      if (i < n) acc1 = __dadd_rn(acc1, v[i]);
      return __dadd_rn(acc1, acc2);
    };
    const double mean0 = __ddiv_rn(accumulate(), (double)n);
    for (int i = 0; i < n; ++i) v[i] = __dsub_rn(v[i], mean0);
    const double mean = __ddiv_rn(accumulate(), (double)n);  // op_var::direct_var takes the mean of the centred vector again
    double a2 = 0.0, a3 = 0.0;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
      const double ti = __dsub_rn(mean, v[i]), tj = __dsub_rn(mean, v[j]);
      a2 = __dadd_rn(a2, __dadd_rn(__dmul_rn(ti, ti), __dmul_rn(tj, tj)));
      a3 = __dadd_rn(a3, __dadd_rn(ti, tj));
    }
    if (i < n) {
      const double ti = __dsub_rn(mean, v[i]);
      a2 = __dadd_rn(a2, __dmul_rn(ti, ti));
      a3 = __dadd_rn(a3, ti);
    }
    const double var = __ddiv_rn(__dsub_rn(a2, __ddiv_rn(__dmul_rn(a3, a3), (double)n)), (double)(n - 1));
    s_mean = mean0;
    s_sd = __dsqrt_rn(var);
    out[0] = s_mean;
    out[1] = s_sd;
  }
  __syncthreads();
  const double sd = s_sd;
  for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = __ddiv_rn(v[i], sd);
}

}  // namespace ace
