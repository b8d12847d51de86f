// fastmath.cuh -- branch-free FP64 exp / sqrt / reciprocal for the pair kernels.
//
// The build and gradient kernels evaluate one exp (SE) or exp + sqrt (+ reciprocal, Matern gradient)
// per (pair, additive term); with CUDA's IEEE-exact library versions those calls were ~70 instructions
// per term and dominated both kernels (ncu, profiles/r01).  These versions have no slow paths and are
// accurate to ~1-2 ulp, far inside the 1e-12 absolute / 1e-9 relative parity budget.
#pragma once
#include "common.cuh"

namespace ace {

// exp(x) for x <= ~700.  Arguments below -700 are clamped (result < 1e-304, i.e. 0 for our purposes).
__device__ __forceinline__ double fast_exp(double x0) {
  const double x = (x0 < -700.0) ? -700.0 : x0;  // NaN stays NaN (and is returned as such below)
  double t = fma(x, 1.4426950408889634, 6755399441055744.0);  // round(x / ln2) via the 1.5 * 2^52 trick
  const int n = __double2loint(t);
  t -= 6755399441055744.0;
  double r = fma(t, -6.93147180369123816490e-01, x);  // x - n ln2 (hi, lo)
  r = fma(t, -1.90821492927058770002e-10, r);
  // Taylor degree 13 on |r| <= ln2/2: remainder < 5e-18
  double p = 1.6059043836821613e-10;
  p = fma(p, r, 2.08767569878681e-09);
  p = fma(p, r, 2.505210838544172e-08);
  p = fma(p, r, 2.755731922398589e-07);
  p = fma(p, r, 2.7557319223985893e-06);
  p = fma(p, r, 2.48015873015873e-05);
  p = fma(p, r, 1.984126984126984e-04);
  p = fma(p, r, 1.388888888888889e-03);
  p = fma(p, r, 8.333333333333333e-03);
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));  // * 2^n
  return (x0 == x0) ? res : x0;
}

// 1/x for normal x (|x| in [1e-300, 1e300]): MUFU seed (~2^-23) + one cubic step -> ~1 ulp
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  y = fma(y, t, y);
  // the seed only carries ~20 bits: a second (quadratic) step makes the result independent of its quality
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
