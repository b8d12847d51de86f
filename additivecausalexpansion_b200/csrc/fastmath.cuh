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

// exp(x) = 2^m * 2^(j/64) * e^r with n = round(64 x / ln2) = 64 m + j and |r| <= ln2/128: a 64-entry table (read
// through the read-only path: the index differs per lane) and a degree-5 Taylor polynomial (r^6/720 < 4e-17) instead
// of a degree-12 one on |r| <= ln2/2 -- 11 FP64 instructions instead of 18 per call, and the pair kernels are bound by
// the FP64 pipe.  Error <= 1.4 ulp (checked against mpmath over [-700, 700]).
static __device__ const double kExp2Tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};
static __constant__ double kExpC[12] = {
    8.333333333333333e-03,   // 1/5!
    4.1666666666666664e-02,  // 1/4!
    1.6666666666666666e-01,  // 1/3!
    0.5,                     // 1/2!
    -0.01083042469326756,    // -ln2/64, high part (32 significant bits: n * hi is exact)
    -2.9815858269852933e-12, // -ln2/64, low part
    92.33248261689366,       // 64 / ln2
    6755399441055744.0,      // 1.5 * 2^52
    0.0, 0.0, 0.0, 0.0};

// exp(x).  Arguments below -700 are clamped (result < 1e-304, i.e. 0 for our purposes); above 709.4 the result is
// +inf like the reference's std::exp beyond 709.78 (a diverged run must surface as "gradients are not finite",
// R/optimizer_classes.R:26-29, not as a silently zeroed kernel term); NaN propagates (the comparisons are false
// for NaN and the final steps are multiplications).
__device__ __forceinline__ double fast_exp(double x0) {
  const double x = (x0 < -700.0) ? -700.0 : x0;
  double t = fma(x, kExpC[6], kExpC[7]);  // round(64 x / ln2) via the 1.5 * 2^52 trick
  const int n = __double2loint(t);
  t -= kExpC[7];
  double r = fma(t, kExpC[4], x);  // x - n ln2/64 (hi, lo)
  r = fma(t, kExpC[5], r);
  double p = fma(kExpC[0], r, kExpC[1]);
  p = fma(p, r, kExpC[2]);
  p = fma(p, r, kExpC[3]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= __ldg(&kExp2Tab[n & 63]);
  const double v = p * __hiloint2double(((n >> 6) + 1023) << 20, 0);  // * 2^m, m in [-1010, 1023]
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

// ---------------------------------------------------------------------------------------------
// Lock-step versions: N independent arguments advance one operation at a time, so that the instruction stream
// carries N-fold instruction-level parallelism by construction.  The pair kernels run 4 warps per scheduler and a
// dependent DFMA chain issues one instruction per 9 cycles (profiles/microbench/fp64_latency_r01.json); ncu r02 showed
// the scalar versions above scheduled back to back (0.48 IPC, `wait` the top stall).
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void fast_exp_n(double (&x)[N]) {  // x[k] <- exp(x[k]), same semantics as fast_exp
  double xc[N], t[N], r[N], p[N], tb[N];
  int n[N];
#pragma unroll
  for (int k = 0; k < N; ++k) xc[k] = (x[k] < -700.0) ? -700.0 : x[k];
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] = fma(xc[k], kExpC[6], kExpC[7]);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    n[k] = __double2loint(t[k]);
    tb[k] = __ldg(&kExp2Tab[n[k] & 63]);
  }
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] -= kExpC[7];
#pragma unroll
  for (int k = 0; k < N; ++k) r[k] = fma(t[k], kExpC[4], xc[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) r[k] = fma(t[k], kExpC[5], r[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = fma(kExpC[0], r[k], kExpC[1]);
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = fma(p[k], r[k], kExpC[2]);
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = fma(p[k], r[k], kExpC[3]);
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = fma(p[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = fma(p[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] *= tb[k];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double v = p[k] * __hiloint2double(((n[k] >> 6) + 1023) << 20, 0);
    x[k] = (x[k] > 709.4) ? __longlong_as_double(0x7ff0000000000000LL) : v;
  }
}

// x[k] <- 1 / x[k] for normal x: MUFU seed (~2^-23) + one cubic step (relative error ~2^-69 + rounding: <= 1 ulp)
template <int N>
__device__ __forceinline__ void fast_rcp_n(double (&x)[N]) {
  double y[N], e[N];
#pragma unroll
  for (int k = 0; k < N; ++k) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(x[k]));
#pragma unroll
  for (int k = 0; k < N; ++k) e[k] = fma(-x[k], y[k], 1.0);
#pragma unroll
  for (int k = 0; k < N; ++k) e[k] = fma(e[k], e[k], e[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) x[k] = fma(y[k], e[k], y[k]);
}

// x[k] <- sqrt(x[k]) for x >= 0 (exact 0 for x == 0).  REFINE adds the Newton step on the root that makes the result
// correctly rounded in all but rare cases (kernel build: entries compared at 1e-12 absolute); without it the root
// carries the rounding of two multiplications (<= 2 ulp), enough for the gradient pass.
template <int N, bool REFINE>
__device__ __forceinline__ void fast_sqrt_n(double (&x)[N]) {
  double y[N], e[N], s[N];
#pragma unroll
  for (int k = 0; k < N; ++k) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(x[k]));
#pragma unroll
  for (int k = 0; k < N; ++k) e[k] = -x[k] * y[k];
#pragma unroll
  for (int k = 0; k < N; ++k) e[k] = fma(e[k], y[k], 1.0);
#pragma unroll
  for (int k = 0; k < N; ++k) s[k] = fma(0.375, e[k], 0.5);
#pragma unroll
  for (int k = 0; k < N; ++k) e[k] *= y[k];
#pragma unroll
  for (int k = 0; k < N; ++k) y[k] = fma(e[k], s[k], y[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) s[k] = x[k] * y[k];
  if (REFINE) {
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = fma(-s[k], s[k], x[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) y[k] *= 0.5;
#pragma unroll
    for (int k = 0; k < N; ++k) s[k] = fma(e[k], y[k], s[k]);
  }
#pragma unroll
  for (int k = 0; k < N; ++k) x[k] = (x[k] > 1e-290) ? s[k] : 0.0;
}

}  // namespace ace
