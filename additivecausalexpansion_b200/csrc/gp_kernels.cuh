// gp_kernels.cuh -- the O(n^2) kernels of the empirical-Bayes GP step (sm_100a):
//   prep_tables_kernel   theta -> exp(-L) tables, e^sigma                       (O(P))
//   logabs_kernel        log|Z| once per upload
//   kernmat_kernel       fused additive kernel build (SE / Matern-3/2, symmetric or rectangular)
//   gemv2_kernel         u = K^-1 y, s = K^-1 1 in one pass over K^-1
//   alpha_kernel         mu closed form, alpha = u - mu s
//   grad_kernel          fused trace-gradient pass: reads K^-1 once, recomputes k_b and D^2_d per pair
//   finalize_kernel      assembles gradients/stats, clip, Nadam/Adam/Nesterov, mu refresh
//
// Reference semantics (paths relative to /root/reference), including its quirks:
//   build      src/kernel_SE_cpp.cpp:9-134, src/kernel_Matern_cpp.cpp:52-93,190-240
//              length-scale of (d, b) read at theta[1 + b + B*(d+1)]            (Q1)
//   gradient   src/kernel_SE_cpp.cpp:161-243, src/kernel_Matern_cpp.cpp:340-377,420-467,
//              src/include/ace_kernel_utils.hpp:23-36; length-scale at theta[2 + B + b + B*d];
//              Matern constant -0.25*9 and 1+sqrt(3*D_grad) denominator        (Q2)
//              evidence uses y'alpha                                           (Q3)
//   mu         src/utilities_cpp.cpp:6-10 (0.5 * sum(K^-1 y) / sum(K^-1))       (Q4)
//   clip/opt   src/utilities_cpp.cpp:121-129, src/optimizer_cpp.cpp:8-63       (Q5)
// The reference materialises the n x n x B cube; nothing here does (only the API-compat cube
// output of kernmat_kernel, on request).
#pragma once
#include "common.cuh"

namespace ace {

constexpr int PMAX = 64;   // max confounder columns supported by the fused kernels
constexpr int BMAXT = 32;  // max additive terms (B = Bz + 1)

// derived-table layout (doubles) in the device buffer `tab`
constexpr int TAB_ESIG = 0;                       // exp(theta[0])
constexpr int TAB_LAM = 8;                        // lambda_b            [BMAXT]
constexpr int TAB_WB = TAB_LAM + BMAXT;           // exp(-L_build[d][b]) [PMAX][BMAXT]
constexpr int TAB_WG = TAB_WB + PMAX * BMAXT;     // exp(-L_grad[d][b])  [PMAX][BMAXT]
constexpr int TAB_SIZE = TAB_WG + PMAX * BMAXT;

// scalar-slot layout (doubles) in the device buffer `sc`
enum {
  SC_SUM_U = 0, SC_SUM_S, SC_MU_USED, SC_RMSE, SC_EVID, SC_GNORM, SC_FINITE, SC_ITER, SC_LOGDET,
  SC_YTALPHA, SC_SUM_ALPHA, SC_TRW, SC_RESID2, SC_COUNT
};

__global__ void prep_tables_kernel(const double* __restrict__ theta, int p, int B, double* __restrict__ tab) {
  const int t = threadIdx.x;
  if (t == 0) tab[TAB_ESIG] = exp(theta[0]);
  for (int b = t; b < BMAXT; b += blockDim.x) tab[TAB_LAM + b] = (b < B) ? theta[2 + b] : 0.0;
  for (int idx = t; idx < PMAX * BMAXT; idx += blockDim.x) {
    const int d = idx / BMAXT, b = idx % BMAXT;
    double wb = 0.0, wg = 0.0;
    if (d < p && b < B) {
      wb = exp(-theta[1 + b + B * (d + 1)]);  // build indexing (Q1)
      wg = exp(-theta[2 + B + b + B * d]);    // gradient indexing
    }
    tab[TAB_WB + idx] = wb;
    tab[TAB_WG + idx] = wg;
  }
}

__global__ void logabs_kernel(const double* __restrict__ z, double* __restrict__ lz, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) lz[i] = log(fabs(z[i]));
}

__device__ __forceinline__ double sgn(double x) { return (double)((0.0 < x) - (x < 0.0)); }

// one additive term of the kernel for one pair.  `first`/`second` follow the reference's evaluation
// order: lambda - D + log|z_first| + log|z_second| (SE, src/kernel_SE_cpp.cpp:53,119) and
// (1+sqrt3 r) exp(lambda - sqrt3 r) z_first z_second (Matern, src/kernel_Matern_cpp.cpp:86,227).
template <int KIND>
__device__ __forceinline__ double term_value(int b, double lam, double D, double z1, double z2, double lz1,
                                             double lz2) {
  if (KIND == 0) {
    if (b == 0) return exp(lam - D);
    if (z1 == 0.0 || z2 == 0.0) return 0.0;
    return (sgn(z1) * sgn(z2)) * exp(lam - D + lz1 + lz2);
  } else {
    const double s3 = 1.7320508075688772;  // sqrt(3.0)
    const double r = sqrt(D);
    const double base = (1.0 + s3 * r) * exp(lam - s3 * r);
    if (b == 0) return base;
    if (z1 == 0.0 || z2 == 0.0) return 0.0;
    return base * z1 * z2;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused kernel build.  64 x 64 pair tile per CTA, X/Z/log|Z| tiles staged by TMA bulk copies.
// ---------------------------------------------------------------------------------------------
struct KernArgs {
  const double *X1, *Z1, *LZ1;  // row points   (n1, ld1)
  const double *X2, *Z2, *LZ2;  // column points (n2, ld2)
  long ld1, ld2;
  int n1, n2, n1_pad, n2_pad, p, B;
  const double* tab;
  double* K;        // n1_pad x n2_pad, ldk
  long ldk;
  double* cube;     // optional: B slices of (ldk x n2_pad)
  long cube_slice;
  int sym;          // 1: X1 == X2, lower tiles computed and mirrored, exactly symmetric output
  int add_noise;    // sym: K_ii += e^sigma for i < n
  int pad_identity; // sym: rows/cols >= n form an identity block
  int skip0;        // 1: leave the nuisance term b = 0 out of the sum (marginal kernels, src/pred_cpp.cpp:55-63)
};

namespace kb {
constexpr int T = 64;        // tile edge
constexpr int LDT = T + 1;   // staging tile stride
inline size_t smem_bytes(int p, int Bz, int bmax, bool sym) {
  size_t d = (size_t)(2 * p + 4 * Bz) * T + (size_t)p * bmax + bmax + (sym ? (size_t)T * LDT : 0);
  return d * 8 + 16;
}
}  // namespace kb

template <int BMAX, int KIND>
__global__ void __launch_bounds__(256) kernmat_kernel(const KernArgs a) {
  using namespace kb;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int p = a.p, B = a.B, Bz = a.B - 1;
  double* Xi = reinterpret_cast<double*>(smraw);
  double* Xj = Xi + p * T;
  double* Zi = Xj + p * T;
  double* Zj = Zi + Bz * T;
  double* LZi = Zj + Bz * T;
  double* LZj = LZi + Bz * T;
  double* wb = LZj + Bz * T;        // [p][BMAX]
  double* lam = wb + p * BMAX;      // [BMAX]
  double* Tt = lam + BMAX;          // [T][LDT] (sym only)
  uint64_t* bar = reinterpret_cast<uint64_t*>(Tt + (a.sym ? T * LDT : 0));

  int ti, tj;
  if (a.sym) {
    const long L = blockIdx.x;
    long t = (long)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while (t * (t + 1) / 2 > L) --t;
    while ((t + 1) * (t + 2) / 2 <= L) ++t;
    ti = (int)t;
    tj = (int)(L - t * (t + 1) / 2);
  } else {
    const int tm = a.n1_pad / T;
    ti = blockIdx.x % tm;
    tj = blockIdx.x / tm;
  }
  const int i0 = ti * T, j0 = tj * T;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)((2 * p + 4 * Bz) * T * 8));
    __syncwarp();
    for (int c = lane; c < p; c += 32) {
      tma_bulk_g2s(Xi + c * T, a.X1 + i0 + (size_t)c * a.ld1, T * 8, bar);
      tma_bulk_g2s(Xj + c * T, a.X2 + j0 + (size_t)c * a.ld2, T * 8, bar);
    }
    for (int c = lane; c < Bz; c += 32) {
      tma_bulk_g2s(Zi + c * T, a.Z1 + i0 + (size_t)c * a.ld1, T * 8, bar);
      tma_bulk_g2s(Zj + c * T, a.Z2 + j0 + (size_t)c * a.ld2, T * 8, bar);
      tma_bulk_g2s(LZi + c * T, a.LZ1 + i0 + (size_t)c * a.ld1, T * 8, bar);
      tma_bulk_g2s(LZj + c * T, a.LZ2 + j0 + (size_t)c * a.ld2, T * 8, bar);
    }
  }
  for (int idx = threadIdx.x; idx < p * BMAX; idx += 256) {
    const int d = idx / BMAX, b = idx % BMAX;
    wb[idx] = a.tab[TAB_WB + d * BMAXT + b];
  }
  if (threadIdx.x < BMAX) lam[threadIdx.x] = a.tab[TAB_LAM + threadIdx.x];
  const double esig = a.tab[TAB_ESIG];
  __syncthreads();
  mbar_wait(bar, 0);

  const int li = threadIdx.x & 63, cg = threadIdx.x >> 6;
  const int gi = i0 + li;
  constexpr int Q = (BMAX <= 16) ? 4 : 2;  // pairs per thread per pass (bounds the accumulator registers)
#pragma unroll 1
  for (int step = 0; step < 16 / Q; ++step) {
    const int jj0 = cg * 16 + step * Q;
    double acc[Q][BMAX];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int b = 0; b < BMAX; ++b) acc[q][b] = 0.0;
#pragma unroll 2
    for (int d = 0; d < p; ++d) {
      const double xi = Xi[d * T + li];
      double d2[Q];
#pragma unroll
      for (int q = 0; q < Q; q += 2) {
        const double2 xa = *reinterpret_cast<const double2*>(Xj + d * T + jj0 + q);
        d2[q] = (xi - xa.x) * (xi - xa.x);
        d2[q + 1] = (xi - xa.y) * (xi - xa.y);
      }
#pragma unroll
      for (int b = 0; b < BMAX; b += 2) {
        const double2 w = *reinterpret_cast<const double2*>(wb + d * BMAX + b);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          acc[q][b] = fma(d2[q], w.x, acc[q][b]);
          acc[q][b + 1] = fma(d2[q], w.y, acc[q][b + 1]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int jj = jj0 + q;
      const int gj = j0 + jj;
      double ksum = 0.0;
#pragma unroll
      for (int b = 0; b < BMAX; ++b) {
        if (b < B) {
          double zi = 1.0, zj = 1.0, lzi = 0.0, lzj = 0.0;
          if (b > 0) {
            zi = Zi[(b - 1) * T + li];
            zj = Zj[(b - 1) * T + jj];
            lzi = LZi[(b - 1) * T + li];
            lzj = LZj[(b - 1) * T + jj];
          }
          // symmetric build: the reference evaluates the r <= c half, i.e. the smaller index (our
          // column point in a lower tile) comes first; rectangular build: row point first.
          const double kv = a.sym ? term_value<KIND>(b, lam[b], acc[q][b], zj, zi, lzj, lzi)
                                  : term_value<KIND>(b, lam[b], acc[q][b], zi, zj, lzi, lzj);
          if (!(a.skip0 && b == 0)) ksum += kv;
          if (a.cube != nullptr && gi < a.n1 && gj < a.n2) {
            if (!a.sym) {
              a.cube[(size_t)b * a.cube_slice + gi + (size_t)gj * a.ldk] = kv;
            } else if (gi >= gj) {
              a.cube[(size_t)b * a.cube_slice + gi + (size_t)gj * a.ldk] = kv;
              a.cube[(size_t)b * a.cube_slice + gj + (size_t)gi * a.ldk] = kv;
            }
          }
        }
      }
      if (a.sym) {
        Tt[li * LDT + jj] = ksum;
      } else {
        a.K[gi + (size_t)gj * a.ldk] = (gi < a.n1 && gj < a.n2) ? ksum : 0.0;
      }
    }
  }
  if (!a.sym) return;
  __syncthreads();
  const bool diag_tile = (ti == tj);
  // pass 1: K[i0+ii, j0+jj]  (ii fastest -> coalesced)
  for (int idx = threadIdx.x; idx < T * T; idx += 256) {
    const int ii = idx & 63, jj = idx >> 6;
    const int gr = i0 + ii, gc = j0 + jj;
    double v = (diag_tile && ii < jj) ? Tt[jj * LDT + ii] : Tt[ii * LDT + jj];
    if (gr >= a.n1 || gc >= a.n1) v = (a.pad_identity && gr == gc) ? 1.0 : 0.0;
    else if (a.add_noise && gr == gc) v += esig;
    a.K[gr + (size_t)gc * a.ldk] = v;
  }
  if (diag_tile) return;
  // pass 2: mirror K[j0+jj, i0+ii]  (jj fastest)
  for (int idx = threadIdx.x; idx < T * T; idx += 256) {
    const int jj = idx & 63, ii = idx >> 6;
    const int gr = j0 + jj, gc = i0 + ii;
    double v = Tt[ii * LDT + jj];
    if (gr >= a.n1 || gc >= a.n1) v = 0.0;
    a.K[gr + (size_t)gc * a.ldk] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// u = Kinv y, s = Kinv 1 (columns j < n only): partial sums per column chunk, deterministic.
// ---------------------------------------------------------------------------------------------
namespace gv {
constexpr int ROWS = 256, CHUNK = 512;
}

__global__ void __launch_bounds__(gv::ROWS) gemv2_kernel(const double* __restrict__ Kinv, long ld, int n,
                                                         const double* __restrict__ y, double* __restrict__ pu,
                                                         double* __restrict__ ps, int n_pad) {
  using namespace gv;
  __shared__ double ys[CHUNK];
  const int i = blockIdx.x * ROWS + threadIdx.x;
  const int c0 = blockIdx.y * CHUNK;
  for (int t = threadIdx.x; t < CHUNK; t += ROWS) {
    const int j = c0 + t;
    ys[t] = (j < n) ? y[j] : 0.0;
  }
  __syncthreads();
  const int cend = min(CHUNK, n - c0);  // columns of this chunk that are real data
  double u = 0.0, s = 0.0;
  if (i < n_pad) {
    const double* col = Kinv + i + (size_t)c0 * ld;
    int t = 0;
    for (; t + 8 <= cend; t += 8) {
      double k[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) k[e] = col[(size_t)(t + e) * ld];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        u = fma(k[e], ys[t + e], u);
        s += k[e];
      }
    }
    for (; t < cend; ++t) {
      const double k = col[(size_t)t * ld];
      u = fma(k, ys[t], u);
      s += k;
    }
    pu[(size_t)blockIdx.y * n_pad + i] = u;
    ps[(size_t)blockIdx.y * n_pad + i] = s;
  }
}

__device__ __forceinline__ double block_sum_1024(double v, double* red) {
  // deterministic block reduction, blockDim.x <= 1024; red has 32 doubles; result broadcast to all
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double r = (threadIdx.x < (blockDim.x + 31) / 32) ? red[threadIdx.x] : 0.0;
  if (w == 0) {
    r = warp_sum(r);
    if (l == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

// one CTA: gathers the gemv partials, mu closed form (first iteration only), alpha = u - mu*s
__global__ void __launch_bounds__(1024) alpha_kernel(const double* __restrict__ pu, const double* __restrict__ ps,
                                                     int nchunks, int n, int n_pad, double* __restrict__ theta,
                                                     double* __restrict__ uvec, double* __restrict__ svec,
                                                     double* __restrict__ alpha, double* __restrict__ Ka,
                                                     double* __restrict__ sc, int set_mu_first_iter) {
  __shared__ double red[32];
  double su = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    double u = 0.0, s = 0.0;
    if (i < n) {
      for (int c = 0; c < nchunks; ++c) {
        u += pu[(size_t)c * n_pad + i];
        s += ps[(size_t)c * n_pad + i];
      }
    }
    uvec[i] = u;
    svec[i] = s;
    su += u;
    ss += s;
  }
  su = block_sum_1024(su, red);
  ss = block_sum_1024(ss, red);
  // R/kernel_SE_R6.R:45 -- on the first iteration mu is replaced by its closed form BEFORE the gradient
  double mu = theta[1];
  if (set_mu_first_iter && sc[SC_ITER] == 1.0) mu = 0.5 * su / ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    theta[1] = mu;
    sc[SC_SUM_U] = su;
    sc[SC_SUM_S] = ss;
    sc[SC_MU_USED] = mu;
  }
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    alpha[i] = (i < n) ? (uvec[i] - mu * svec[i]) : 0.0;
    Ka[i] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused trace-gradient pass.  Persistent CTAs stride over the lower 64x64 pair tiles; a thread
// owns one row of the tile and BT of the additive terms (its warp's b-group), and keeps the
// PD*BT length-scale sums + BT scale sums in registers across all tiles.  Per pair it reads
// K^-1(i,j) once, rebuilds D^2_d and k_b in registers, and never touches an n x n x B cube.
// Partial sums leave through warp shuffles -> smem -> one row of `partials` per CTA (no atomics).
// ---------------------------------------------------------------------------------------------
struct GradArgs {
  const double *X, *Z, *LZ;  // n_pad x p, n_pad x Bz (ld = ldx)
  long ldx;
  const double* Kinv;
  long ld;
  const double* alpha;
  double* Ka;        // K * alpha accumulated with atomics (statistic only)
  const double* tab;
  double* partials;  // [gridDim.y * gridDim.x][P]
  int n, p, B, P;
  int ntiles_side;   // ceil(n / 64)
};

namespace gk {
constexpr int T = 64;
constexpr int GROUPS_PER_CTA = 4;
inline size_t smem_bytes(int PD, int Bz, int BT, int kind) {
  const int BTP = 4 * ((BT + 3) / 4);
  size_t d = (size_t)2 * PD * T + (size_t)4 * Bz * T + 2 * T + (size_t)GROUPS_PER_CTA * PD * BTP * (kind ? 2 : 1) +
             BMAXT + (size_t)8 * (PD * BT + BT);
  return d * 8 + 16;
}
}  // namespace gk

template <int PD, int BT, int KIND>
__global__ void __launch_bounds__(256, 1) grad_kernel(const GradArgs a) {
  using namespace gk;
  static_assert(BT <= 16, "BT too large");
  constexpr int BTP = 4 * ((BT + 3) / 4);  // packed weights per d, padded for 16-byte loads
  extern __shared__ __align__(128) unsigned char smraw[];
  const int p = a.p, B = a.B, Bz = a.B - 1;
  const int ngroups_cta = blockDim.x / 64;  // b-groups handled by this CTA (<= 4)
  double* Xi = reinterpret_cast<double*>(smraw);
  double* Xj = Xi + PD * T;
  double* Zi = Xj + PD * T;
  double* Zj = Zi + Bz * T;
  double* LZi = Zj + Bz * T;
  double* LZj = LZi + Bz * T;
  double* ai = LZj + Bz * T;  // alpha rows
  double* aj = ai + T;        // alpha cols
  double* wbp = aj + T;       // [GROUPS_PER_CTA][PD][BTP] build weights packed per group
  double* wgp = wbp + GROUPS_PER_CTA * PD * BTP;  // same for gradient weights (Matern only)
  double* lam = wgp + (KIND ? GROUPS_PER_CTA * PD * BTP : 0);
  double* red = lam + BMAXT;  // [8 warps][PD*BT + BT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 8 * (PD * BT + BT));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pw = warp & 1;             // which half of the 64 rows
  const int gl = warp >> 1;            // local b-group
  const int bg = blockIdx.y * GROUPS_PER_CTA + gl;  // global b-group
  const int b0 = bg * BT;
  const int li = pw * 32 + lane;

  // tables
  for (int idx = threadIdx.x; idx < ngroups_cta * PD * BTP; idx += blockDim.x) {
    const int g = idx / (PD * BTP), r = idx % (PD * BTP), d = r / BTP, t = r % BTP;
    const int b = (blockIdx.y * GROUPS_PER_CTA + g) * BT + t;
    const bool ok = (t < BT) && (b < B) && (d < p);
    wbp[idx] = ok ? a.tab[TAB_WB + d * BMAXT + b] : 0.0;
    if (KIND) wgp[idx] = ok ? a.tab[TAB_WG + d * BMAXT + b] : 0.0;
  }
  for (int b = threadIdx.x; b < BMAXT; b += blockDim.x) lam[b] = a.tab[TAB_LAM + b];
  // zero the padded d rows of the X tiles once (TMA only ever writes rows d < p)
  for (int idx = threadIdx.x; idx < (PD - p) * T; idx += blockDim.x) {
    Xi[p * T + idx] = 0.0;
    Xj[p * T + idx] = 0.0;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  double S[PD][BT];
  double Sb[BT];
#pragma unroll
  for (int d = 0; d < PD; ++d)
#pragma unroll
    for (int t = 0; t < BT; ++t) S[d][t] = 0.0;
#pragma unroll
  for (int t = 0; t < BT; ++t) Sb[t] = 0.0;

  const double* wb_mine = wbp + gl * PD * BTP;
  const double* wg_mine = wgp + gl * PD * BTP;
  const long ntiles = (long)a.ntiles_side * (a.ntiles_side + 1) / 2;
  uint32_t phase = 0;
  for (long L = blockIdx.x; L < ntiles; L += gridDim.x) {
    long tt = (long)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while (tt * (tt + 1) / 2 > L) --tt;
    while ((tt + 1) * (tt + 2) / 2 <= L) ++tt;
    const int ti = (int)tt, tj = (int)(L - tt * (tt + 1) / 2);
    const int i0 = ti * T, j0 = tj * T;
    const bool diag_tile = (ti == tj);
    const double wt = diag_tile ? 1.0 : 2.0;

    __syncthreads();  // previous tile fully consumed before the TMA overwrites the staging tiles
    if (warp == 0) {
      if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)((2 * p + 4 * Bz + 2) * T * 8));
      __syncwarp();
      for (int c = lane; c < p; c += 32) {
        tma_bulk_g2s(Xi + c * T, a.X + i0 + (size_t)c * a.ldx, T * 8, bar);
        tma_bulk_g2s(Xj + c * T, a.X + j0 + (size_t)c * a.ldx, T * 8, bar);
      }
      for (int c = lane; c < Bz; c += 32) {
        tma_bulk_g2s(Zi + c * T, a.Z + i0 + (size_t)c * a.ldx, T * 8, bar);
        tma_bulk_g2s(Zj + c * T, a.Z + j0 + (size_t)c * a.ldx, T * 8, bar);
        tma_bulk_g2s(LZi + c * T, a.LZ + i0 + (size_t)c * a.ldx, T * 8, bar);
        tma_bulk_g2s(LZj + c * T, a.LZ + j0 + (size_t)c * a.ldx, T * 8, bar);
      }
      if (lane == 0) {
        tma_bulk_g2s(ai, a.alpha + i0, T * 8, bar);
        tma_bulk_g2s(aj, a.alpha + j0, T * 8, bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;

    const int gi = i0 + li;
    const double alpha_i = ai[li];
    const double* kcol = a.Kinv + gi + (size_t)j0 * a.ld;
    double rowacc = 0.0;
    double knext = kcol[0];
#pragma unroll 1
    for (int jj = 0; jj < T; ++jj) {
      const int gj = j0 + jj;
      const double kinv = knext;
      if (jj + 1 < T) knext = kcol[(size_t)(jj + 1) * a.ld];
      const double alpha_j = aj[jj];
      const bool valid = (gi < a.n) && (gj < a.n);
      const double W = valid ? wt * (kinv - alpha_i * alpha_j) : 0.0;
      double d2[PD];
      double D[BT], Dg[BT];
#pragma unroll
      for (int t = 0; t < BT; ++t) D[t] = Dg[t] = 0.0;
#pragma unroll
      for (int d = 0; d < PD; ++d) {
        const double df = Xi[d * T + li] - Xj[d * T + jj];
        d2[d] = df * df;
        double w[BTP];
#pragma unroll
        for (int t = 0; t < BTP; t += 2) {
          const double2 v = *reinterpret_cast<const double2*>(wb_mine + d * BTP + t);
          w[t] = v.x;
          w[t + 1] = v.y;
        }
#pragma unroll
        for (int t = 0; t < BT; ++t) D[t] = fma(d2[d], w[t], D[t]);
        if (KIND) {
#pragma unroll
          for (int t = 0; t < BTP; t += 2) {
            const double2 v = *reinterpret_cast<const double2*>(wg_mine + d * BTP + t);
            w[t] = v.x;
            w[t + 1] = v.y;
          }
#pragma unroll
          for (int t = 0; t < BT; ++t) Dg[t] = fma(d2[d], w[t], Dg[t]);
        }
      }
      double kpart = 0.0;
#pragma unroll
      for (int t = 0; t < BT; ++t) {
        const int b = b0 + t;
        if (b < B) {
          double zi = 1.0, zj = 1.0, lzi = 0.0, lzj = 0.0;
          if (b > 0) {
            zi = Zi[(b - 1) * T + li];
            zj = Zj[(b - 1) * T + jj];
            lzi = LZi[(b - 1) * T + li];
            lzj = LZj[(b - 1) * T + jj];
          }
          const double kv = term_value<KIND>(b, lam[b], D[t], zj, zi, lzj, lzi);
          kpart += kv;
          Sb[t] = fma(W, kv, Sb[t]);
          // SE: dK/dL = K_b * D2_d * exp(-L);  Matern (as written): K_b / (1 + sqrt(3 D_grad)) * D2_d * exp(-L)
          const double tv = KIND ? W * (kv / (1.0 + sqrt(3.0 * Dg[t]))) : W * kv;
#pragma unroll
          for (int d = 0; d < PD; ++d) S[d][t] = fma(tv, d2[d], S[d][t]);
        }
      }
      // K*alpha for the RMSE statistic (src/kernel_SE_cpp.cpp:238): row part in a register, column
      // part (mirror tile) reduced over the 32 rows of this warp
      rowacc = fma(kpart, alpha_j, rowacc);
      if (!diag_tile) {
        const double cpart = warp_sum(kpart * alpha_i);
        if (lane == 0) atomicAdd(a.Ka + gj, cpart);
      }
    }
    atomicAdd(a.Ka + gi, rowacc);
  }

  // ---- CTA reduction: lanes -> warps -> one partial row per CTA -----------------------------
  constexpr int NV = PD * BT + BT;
  __syncthreads();
#pragma unroll
  for (int d = 0; d < PD; ++d)
#pragma unroll
    for (int t = 0; t < BT; ++t) {
      const double v = warp_sum(S[d][t]);
      if (lane == 0) red[warp * NV + d * BT + t] = v;
    }
#pragma unroll
  for (int t = 0; t < BT; ++t) {
    const double v = warp_sum(Sb[t]);
    if (lane == 0) red[warp * NV + PD * BT + t] = v;
  }
  __syncthreads();
  double* out = a.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * a.P;
  for (int idx = threadIdx.x; idx < a.P; idx += blockDim.x) out[idx] = 0.0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < ngroups_cta * NV; idx += blockDim.x) {
    const int g = idx / NV, r = idx % NV;
    const double v = red[(2 * g) * NV + r] + red[(2 * g + 1) * NV + r];
    const int gb0 = (blockIdx.y * GROUPS_PER_CTA + g) * BT;
    if (r < PD * BT) {
      const int d = r / BT, b = gb0 + r % BT;
      if (d < p && b < B) out[2 + B + b + B * d] = v;
    } else {
      const int b = gb0 + (r - PD * BT);
      if (b < B) out[2 + b] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// One CTA: reduce partials, assemble gradient + statistics, clip, optimiser step, mu refresh.
// ---------------------------------------------------------------------------------------------
struct FinalizeArgs {
  const double* partials;
  int nparts;
  const double *y, *alpha, *Ka, *dvec, *Kinv;
  long ld;
  const double* tab;
  double *theta, *m, *v, *grad, *sc;
  int n, p, B, P, kind;
  int optimizer;  // 0 Nadam, 1 Adam, 2 Nesterov
  double lr, beta1, beta2, eps, momentum, std_y, clip_at;
  int norm_clip;
  int do_update;  // 0: gradients + stats only (per-function API)
};

__global__ void __launch_bounds__(1024) finalize_kernel(const FinalizeArgs a) {
  __shared__ double red[32];
  const int n = a.n, B = a.B, P = a.P;
  const double mu = a.sc[SC_MU_USED];
  double s_alpha = 0, s_ya = 0, s_res = 0, s_logd = 0, s_trw = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double al = a.alpha[i];
    s_alpha += al;
    s_ya = fma(a.y[i], al, s_ya);
    const double r = (a.y[i] - mu) - a.Ka[i];
    s_res = fma(r, r, s_res);
    s_logd += log(a.dvec[i]);
    s_trw += a.Kinv[i + (size_t)i * a.ld] - al * al;
  }
  s_alpha = block_sum_1024(s_alpha, red);
  s_ya = block_sum_1024(s_ya, red);
  s_res = block_sum_1024(s_res, red);
  s_logd = block_sum_1024(s_logd, red);
  s_trw = block_sum_1024(s_trw, red);

  // gradient assembly
  double gn2 = 0.0;
  int bad = 0;
  for (int k = threadIdx.x; k < P; k += blockDim.x) {
    double g;
    if (k == 0) {
      g = -0.5 * s_trw * a.tab[TAB_ESIG];  // sigma_gradient
    } else if (k == 1) {
      g = (a.kind == 0) ? s_alpha : 0.0;   // SE: sum(K^-1 ybar); Matern: forced 0
    } else {
      double s = 0.0;
      for (int c = 0; c < a.nparts; ++c) s += a.partials[(size_t)c * P + k];
      if (k < 2 + B) {
        g = -0.5 * s;
      } else {
        const int r = k - 2 - B, b = r % B, d = r / B;
        const double e = a.tab[TAB_WG + d * BMAXT + b];
        g = (a.kind == 0) ? (-0.5 * s) * e : -0.25 * 9 * s * e;
      }
    }
    a.grad[k] = g;
    gn2 = fma(g, g, gn2);
    if (!isfinite(g)) bad = 1;
  }
  gn2 = block_sum_1024(gn2, red);
  const double nbad = block_sum_1024((double)bad, red);
  double L2 = sqrt(gn2);
  // norm_clip_cpp: rescale to unit norm (Q5)
  const bool clip = a.norm_clip && (L2 > a.clip_at) && isfinite(L2) && (L2 != 0.0);
  const double iter = a.sc[SC_ITER];
  const double c1 = 1.0 - pow(a.beta1, iter), c2 = 1.0 - pow(a.beta2, iter);
  double gn2c = 0.0;
  for (int k = threadIdx.x; k < P; k += blockDim.x) {
    double g = a.grad[k];
    if (clip) g = g / L2;
    if (a.do_update) {
      a.grad[k] = g;
      double th = a.theta[k];
      if (a.optimizer == 2) {
        const double nu = a.momentum * a.m[k] + a.lr * g;
        a.m[k] = nu;
        th = th + nu;
      } else {
        const double mk = a.beta1 * a.m[k] + (1 - a.beta1) * g;
        const double vk = a.beta2 * a.v[k] + (1 - a.beta2) * (g * g);
        a.m[k] = mk;
        a.v[k] = vk;
        if (a.optimizer == 0)
          th = th + a.lr * ((a.beta1 * mk + (1 - a.beta1) * g) / c1) / (sqrt(vk / c2) + a.eps);
        else
          th = th + a.lr * (mk / c1) / (sqrt(vk / c2) + a.eps);
      }
      a.theta[k] = th;
    }
    gn2c = fma(g, g, gn2c);
  }
  gn2c = block_sum_1024(gn2c, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    a.sc[SC_RMSE] = a.std_y * sqrt(s_res) / sqrt((double)n);
    a.sc[SC_LOGDET] = 2.0 * s_logd;
    a.sc[SC_YTALPHA] = s_ya;
    a.sc[SC_EVID] = -0.5 * ((double)n * log(2.0 * 3.14159265358979323846) + 2.0 * s_logd + s_ya);
    a.sc[SC_SUM_ALPHA] = s_alpha;
    a.sc[SC_TRW] = s_trw;
    a.sc[SC_RESID2] = s_res;
    a.sc[SC_GNORM] = a.do_update ? sqrt(gn2c) : L2;
    a.sc[SC_FINITE] = (nbad == 0.0) ? 1.0 : 0.0;
    if (a.do_update) {
      // mean_solution with the pre-update inverse (R/kernel_SE_R6.R:54, Q4/Q6)
      a.theta[1] = 0.5 * a.sc[SC_SUM_U] / a.sc[SC_SUM_S];
      a.sc[SC_ITER] = iter + 1.0;
    }
  }
}

// statistics only (stats_cpp, src/stats_cpp.cpp:9-32): needs alpha = Kinv (y - mu) and K alpha
__global__ void __launch_bounds__(1024) stats_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                                                     const double* __restrict__ Ka, const double* __restrict__ dvec,
                                                     int n, double mu, double std_y, double* __restrict__ sc) {
  __shared__ double red[32];
  double s_ya = 0, s_res = 0, s_logd = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_ya = fma(y[i], alpha[i], s_ya);
    const double r = (y[i] - mu) - Ka[i];
    s_res = fma(r, r, s_res);
    s_logd += log(dvec[i]);
  }
  s_ya = block_sum_1024(s_ya, red);
  s_res = block_sum_1024(s_res, red);
  s_logd = block_sum_1024(s_logd, red);
  if (threadIdx.x == 0) {
    sc[SC_RMSE] = std_y * sqrt(s_res) / sqrt((double)n);
    sc[SC_EVID] = -0.5 * ((double)n * log(2.0 * 3.14159265358979323846) + 2.0 * s_logd + s_ya);
    sc[SC_LOGDET] = 2.0 * s_logd;
  }
}

}  // namespace ace
