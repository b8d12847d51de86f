// kernmat2_kernel.cuh -- second-generation fused additive kernel build (sm_100a), exact-shape instantiations.
//
// Same tiling and output paths as kernmat_kernel (pair_kernels.cuh: 64 x 64 pair tile per CTA, lower tiles mirrored
// through shared memory in the symmetric case, optional cube output, rectangular blocks of the sharded build), with the
// instruction stream of grad3_kernel's treatment (ncu r02 of kernmat_kernel<1, 12, 2> at C3: FP64 pipe 52 % busy,
// 0.57 IPC, `wait` the top stall):
//   * BX additive terms compiled in, all of them in one pass over the dimensions (QC x BX accumulators);
//   * length-scale weights and lambda_b in __constant__ memory (copied from the device table before the launch, the
//     launches serialised per device by the caller), distance loop fully unrolled: every weight is a uniform-register
//     operand of its DFMA instead of a shared-memory load;
//   * sqrt / exp in lock step over the QC x G (column, term) pairs of a group (fastmath.cuh: fast_*_n).
// Reference semantics: src/kernel_SE_cpp.cpp:9-134, src/kernel_Matern_cpp.cpp:52-93,190-240 (quirk Q1: the build reads
// the length-scale of (d, b) at theta[1 + b + B (d + 1)] = column b of the extended table).
#pragma once
#include "pair_common.cuh"

namespace ace {

// lambda_b and we[d][c] of the launch in flight (layout: pair_common.cuh G3_*); one copy per translation unit
static __constant__ double cKB[G3_SIZE];

namespace kb2 {
constexpr int T = 64;        // tile edge
constexpr int LDT = T + 1;   // staging tile stride
#if defined(KB2_QC) && defined(KB2_GRP)  // tuning experiments (Makefile EXTRA)
constexpr int qc(int) { return KB2_QC; }
constexpr int grp(int) { return KB2_GRP; }
#else
constexpr int qc(int BX) { return BX <= 6 ? 4 : (BX <= 12 ? 2 : 1); }      // columns per step
constexpr int grp(int BX) { return BX <= 6 ? 1 : (BX <= 12 ? 3 : 4); }     // terms per lock-step group (measured: scripts/variant_time.py)
#endif
inline size_t smem_bytes(int p, int Bz, bool sym) {
  size_t d = (size_t)(2 * p + 4 * Bz) * T + (sym ? (size_t)T * LDT : 0);
  return d * 8 + 16;
}
}  // namespace kb2

// CUBE: also write the B per-term slices (API-compat entry points only; the fit handle never forms the cube)
// SYM: a.sym compiled in (operand order, mirrored output)
template <int KIND, int BX, bool CUBE, bool SYM>
__global__ void __launch_bounds__(256, 2) kernmat2_kernel(const KernArgs a) {
  using namespace kb2;
  static_assert(BX <= G3_LAM && BX <= G3_WS, "shape exceeds the constant table");
  constexpr int QC = qc(BX), G = grp(BX), Bz = BX - 1, NL = QC * G;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int p = a.p;
  double* Xi = reinterpret_cast<double*>(smraw);
  double* Xj = Xi + p * T;
  double* Zi = Xj + p * T;
  double* Zj = Zi + Bz * T;
  double* LZi = Zj + Bz * T;
  double* LZj = LZi + Bz * T;
  double* Tt = LZj + Bz * T;            // [T][LDT] (sym only)
  uint64_t* bar = reinterpret_cast<uint64_t*>(Tt + (SYM ? T * LDT : 0));

  int ti, tj;
  if (SYM) {
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
  const double esig = a.tab[TAB_ESIG];
  mbar_wait(bar, 0);

  const int li = threadIdx.x & 63, cg = threadIdx.x >> 6;
  const int gi = i0 + li;
  constexpr bool sym = SYM;
#pragma unroll 1
  for (int step = 0; step < 16 / QC; ++step) {
    const int jj0 = cg * 16 + step * QC;
    double acc[QC][BX];
#pragma unroll
    for (int q = 0; q < QC; ++q)
#pragma unroll
      for (int t = 0; t < BX; ++t) acc[q][t] = 0.0;
    {
      const double* xi = Xi + li;
      const double* xj = Xj + jj0;
#pragma unroll
      for (int d = 0; d < G3_PD; ++d) {
        if (d >= p) break;  // uniform
        const double xv = xi[d * T];
        double d2[QC];
        if (QC >= 2) {
#pragma unroll
          for (int q = 0; q + 1 < QC; q += 2) {
            const double2 xa = *reinterpret_cast<const double2*>(xj + d * T + q);
            d2[q] = (xv - xa.x) * (xv - xa.x);
            d2[q + 1] = (xv - xa.y) * (xv - xa.y);
          }
        } else {
          const double xa = xj[d * T];
          d2[0] = (xv - xa) * (xv - xa);
        }
#pragma unroll
        for (int t = 0; t < BX; ++t)
#pragma unroll
          for (int q = 0; q < QC; ++q) acc[q][t] = fma(d2[q], cKB[G3_LAM + d * G3_WS + t], acc[q][t]);
      }
    }
    double ksum[QC];
#pragma unroll
    for (int q = 0; q < QC; ++q) ksum[q] = 0.0;
    // operand order without selects: base pointers and the stride over the step's columns
    const double* zfirst = sym ? Zj + jj0 : Zi + li;
    const double* zsecond = sym ? Zi + li : Zj + jj0;
    const double* lzfirst = sym ? LZj + jj0 : LZi + li;
    const double* lzsecond = sym ? LZi + li : LZj + jj0;
    constexpr int qfirst = sym ? 1 : 0, qsecond = sym ? 0 : 1;
#pragma unroll
    for (int t0 = 0; t0 < BX; t0 += G) {
      // lock-step lanes l = g * QC + q over the group's (term, column) pairs; the tail group repeats the last term
      double v[NL], ex[NL], z1[NL], z2[NL];
      bool live[NL];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int b = (t0 + g < BX) ? t0 + g : BX - 1;
#pragma unroll
        for (int q = 0; q < QC; ++q) v[g * QC + q] = acc[q][b];
      }
      if (KIND) fast_sqrt_n<NL, true>(v);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int b = (t0 + g < BX) ? t0 + g : BX - 1;
        const double lb = cKB[b];
#pragma unroll
        for (int q = 0; q < QC; ++q) {
          const int l = g * QC + q;
          // `first` / `second` as the reference evaluates them (term_value, pair_common.cuh): symmetric build = the
          // smaller index (our column point in a lower tile) first, rectangular build = row point first
          double za = 1.0, zb = 1.0;
          if (b > 0) {
            za = zfirst[(b - 1) * T + q * qfirst];
            zb = zsecond[(b - 1) * T + q * qsecond];
          }
          z1[l] = za;
          z2[l] = zb;
          if (KIND == 0) {
            double arg = lb - v[l];
            if (b > 0) arg = arg + lzfirst[(b - 1) * T + q * qfirst] + lzsecond[(b - 1) * T + q * qsecond];
            ex[l] = arg;
            live[l] = (b == 0) || !(za == 0.0 || zb == 0.0);
          } else {
            v[l] = SQRT3 * v[l];
            ex[l] = lb - v[l];
            live[l] = true;
          }
        }
      }
      fast_exp_n<NL>(ex);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (t0 + g < BX) {
          const int b = t0 + g;
#pragma unroll
          for (int q = 0; q < QC; ++q) {
            const int l = g * QC + q;
            double kv;
            if (KIND == 0) {
              kv = (b == 0) ? ex[l] : (live[l] ? (sgn(z1[l]) * sgn(z2[l])) * ex[l] : 0.0);
            } else {
              const double base = (1.0 + v[l]) * ex[l];
              kv = (b == 0) ? base : base * z1[l] * z2[l];  // exactly (+-)0 when a basis value is 0
            }
            if (!(a.skip0 && b == 0)) ksum[q] += kv;
            if (CUBE && a.cube != nullptr) {
              const int gj = j0 + jj0 + q;
              if (gi < a.n1 && gj < a.n2) {
                if (!sym) {
                  a.cube[(size_t)b * a.cube_slice + gi + (size_t)gj * a.ldk] = kv;
                } else if (gi >= gj) {
                  a.cube[(size_t)b * a.cube_slice + gi + (size_t)gj * a.ldk] = kv;
                  a.cube[(size_t)b * a.cube_slice + gj + (size_t)gi * a.ldk] = kv;
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < QC; ++q) {
      const int jj = jj0 + q, gj = j0 + jj;
      if (sym) {
        Tt[li * LDT + jj] = ksum[q];
      } else {
        double vv = (gi < a.n1 && gj < a.n2) ? ksum[q] : 0.0;
        if (a.row_off + gi == a.col_off + gj) {
          if (gi < a.n1 && gj < a.n2) vv += a.add_noise ? esig : 0.0;
          else if (a.pad_identity) vv = 1.0;
        }
        a.K[gi + (size_t)gj * a.ldk] = vv;
      }
    }
  }
  if (!sym) return;
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

}  // namespace ace
