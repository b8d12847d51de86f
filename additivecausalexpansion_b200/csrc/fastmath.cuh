// fastmath.cuh -- branch-free FP64 exp / sqrt / rsqrt / reciprocal for the pair kernels and the Cholesky leaf.
//
// The build and gradient kernels evaluate one exp (SE) or exp + sqrt (+ reciprocal, Matern gradient)
// per (pair, additive term); with CUDA's IEEE-exact library versions those calls were ~70 instructions
// per term and dominated both kernels (ncu, profiles/r01).  These versions have no slow paths and are
// accurate to ~1-2 ulp, far inside the 1e-12 absolute / 1e-9 relative parity budget.  Polynomial
// coefficients live in __constant__ memory so that DFMA takes them as c[][] operands: as 64-bit immediates
// they cost two UMOV each and, re-materialised per call under register pressure, were ~15 % of the
// gradient kernel's instruction stream.
#pragma once
#include "common.cuh"

namespace ace {

// 1/k!, k = 12 .. 2 (Taylor degree 12 on |r| <= ln2/2: remainder < 2e-16 relative), then ln2 split, log2(e)
__constant__ double kExpC[16] = {
    2.08767569878681e-09,    // 1/12!
    2.505210838544172e-08,   // 1/11!
    2.755731922398589e-07,   // 1/10!
    2.7557319223985893e-06,  // 1/9!
    2.48015873015873e-05,    // 1/8!
    1.984126984126984e-04,   // 1/7!
    1.388888888888889e-03,   // 1/6!
    8.333333333333333e-03,   // 1/5!
    4.1666666666666664e-02,  // 1/4!
    1.6666666666666666e-01,  // 1/3!
    0.5,                     // 1/2!
    -6.93147180369123816490e-01,  // -ln2 hi
    -1.90821492927058770002e-10,  // -ln2 lo
    1.4426950408889634,           // log2(e)
    6755399441055744.0,           // 1.5 * 2^52
    0.0};

// exp(x).  Arguments below -700 are clamped (result < 1e-304, i.e. 0 for our purposes); above 709.4 (2^n would need n = 1024) the
// result is +inf like the reference's std::exp beyond 709.78 (a diverged run must surface as "gradients are not finite",
// R/optimizer_classes.R:26-29, not as a silently zeroed kernel term); NaN propagates (the comparisons are false
// for NaN and the final step is a multiplication).
__device__ __forceinline__ double fast_exp(double x0) {
  const double x = (x0 < -700.0) ? -700.0 : x0;
  double t = fma(x, kExpC[13], kExpC[14]);  // round(x / ln2) via the 1.5 * 2^52 trick
  const int n = __double2loint(t);
  t -= kExpC[14];
  double r = fma(t, kExpC[11], x);  // x - n ln2 (hi, lo)
  r = fma(t, kExpC[12], r);
  double p = kExpC[0];
#pragma unroll
  for (int k = 1; k <= 10; ++k) p = fma(p, r, kExpC[k]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double v = p * __hiloint2double((n + 1023) << 20, 0);  // * 2^n, n in [-1010, 1023]
  return (x0 > 709.4) ? __longlong_as_double(0x7ff0000000000000LL) : v;
}

// 1/x for normal x (|x| in [1e-300, 1e300]): MUFU seed (~2^-23) + cubic step + quadratic step -> ~1 ulp
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  y = fma(y, t, y);
  const double e2 = fma(-x, y, 1.0);
  return fma(y, e2, y);
}

// x^-1/2 for normal x > 0 (NaN for x <= 0): MUFU seed + cubic step + one quadratic step -> ~1 ulp
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);
  e = fma(-x * y, y, 1.0);
  return fma(0.5 * y, e, y);
}

// sqrt(x) for x >= 0 (exact 0 for x == 0): MUFU rsqrt seed + cubic step + one Newton step on the root
__device__ __forceinline__ double fast_sqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);  // y ~ x^-1/2 to ~2^-60
  double s = x * y;
  s = fma(fma(-s, s, x), 0.5 * y, s);
  return (x > 1e-290) ? s : 0.0;
}

}  // namespace ace
