// grad2_kernel.cuh -- second-generation fused trace-gradient pass (sm_100a).
//
// grad_kernel (pair_kernels.cuh) gives every thread a slice of the additive terms, so the D^2_d of a pair are
// rebuilt by each of the B/BT thread groups and the kernel is instruction-issue bound (ncu r01: 655 warp
// instructions per (pair, group), FP64 pipe 47 %).  Here ONE thread owns a pair for ALL terms:
//   phase 1  D^2_d (p values) and the B+1 distance sums E_c = sum_d we[d][c] D^2_d      (DFMA)
//   phase 2  k_b, W k_b, and t_b = W k_b [/ (1 + sqrt(3 E_{b+1}))]                       (fastmath)
//   phase 3  S[d][b] += t_b D^2_d  -- a (B x 32 pairs) x (32 pairs x p) product per warp -- on the FP64
//            TENSOR pipe: the 32 pairs of a warp are staged through a warp-private smem tile and consumed
//            as DMMA.8x8x4 fragments, so the p*B length-scale sums live in 2*ceil(B/8)*ceil(p/8) registers
//            per thread instead of p*B.
// Same inputs, outputs (one row of `partials` per CTA, K*alpha by atomics) and semantics as grad_kernel.
#pragma once
#include "pair_kernels.cuh"

namespace ace {

namespace g2 {
constexpr int T = 64;      // tile edge
constexpr int LDS_ = 36;   // stage row stride (32 pairs + 4): conflict-free DMMA fragment loads
inline int wstride(int B) { return 2 * ((B + 2) / 2); }  // >= B + 1, even
// nwarps: 8 (256 threads, up to 255 registers) or 16 (512 threads, 128 registers: twice the warps per scheduler to
// cover the fixed-latency FP64 dependencies -- ncu r02: 2 warps per scheduler issue 0.47 IPC, FP64 pipe 52 %).
// kind 1 (Matern) does not use log|z|: its tiles are not staged.
inline size_t smem_bytes(int PD8, int BD8, int Bz, int nwarps = 8, int kind = 0) {
  const int B = Bz + 1;
  size_t d = (size_t)2 * PD8 * T + (size_t)(kind ? 2 : 4) * Bz * T + 2 * T + (size_t)PD8 * wstride(B) + BD8 +
             (size_t)nwarps * (BD8 + PD8) * LDS_;
  return d * 8 + 16;
}
}  // namespace g2

template <int PD8, int BD8, int KIND, int NWARPS = 8>
__global__ void __launch_bounds__(NWARPS * 32, 1) grad2_kernel(const GradArgs a) {
  using namespace g2;
  constexpr int NTHR = NWARPS * 32;
  constexpr int JCOLS = 128 / NWARPS;  // columns of the 64 x 64 tile per warp: 16 (8 warps) or 8 (16 warps)
  constexpr int MT = BD8 / 8, NT = PD8 / 8;
  constexpr int NE = BD8 + 1;  // distance sums kept per thread (c = 0..B, B <= BD8)
  extern __shared__ __align__(128) unsigned char smraw[];
  const int p = a.p, B = a.B, Bz = a.B - 1;
  const int WS = 2 * ((B + 2) / 2);
  const int p4 = (p + 3) / 4 * 4;  // <= PD8
  double* Xi = reinterpret_cast<double*>(smraw);
  double* Xj = Xi + PD8 * T;
  double* Zi = Xj + PD8 * T;
  double* Zj = Zi + Bz * T;
  double* LZi = Zj + Bz * T;                      // Matern: not staged (term_value<1> does not read log|z|)
  double* LZj = LZi + (KIND ? 0 : Bz * T);
  double* ai = LZj + (KIND ? 0 : Bz * T);
  double* aj = ai + T;
  double* wt_s = aj + T;            // [PD8][WS] extended weight table rows (columns 0..B)
  double* lam = wt_s + PD8 * WS;    // [BD8]
  double* stage = lam + BD8;        // [NWARPS][(BD8 + PD8)][LDS_]
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + NWARPS * (BD8 + PD8) * LDS_);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int li = (warp & 1) * 32 + lane;   // row of the tile owned by this thread
  const int jbase = (warp >> 1) * JCOLS;   // this warp's columns of the tile
  double* Ts = stage + warp * (BD8 + PD8) * LDS_;  // [BD8][LDS_]  t_b of the warp's 32 pairs
  double* Ds = Ts + BD8 * LDS_;                    // [PD8][LDS_]  D^2_d of the warp's 32 pairs

  for (int idx = threadIdx.x; idx < PD8 * WS; idx += NTHR) {
    const int d = idx / WS, c = idx % WS;
    wt_s[idx] = (d < p && c <= B) ? a.tab[TAB_WE + d * WSTRIDE + c] : 0.0;
  }
  for (int b = threadIdx.x; b < BD8; b += NTHR) lam[b] = (b < B) ? a.tab[TAB_LAM + b] : 0.0;
  for (int idx = threadIdx.x; idx < (PD8 - p) * T; idx += NTHR) {  // padded d rows stay zero (TMA never writes them)
    Xi[p * T + idx] = 0.0;
    Xj[p * T + idx] = 0.0;
  }
  for (int idx = threadIdx.x; idx < NWARPS * (BD8 + PD8) * LDS_; idx += NTHR) stage[idx] = 0.0;  // rows b >= B stay zero
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  double acc[MT][NT][2];   // S[b = 8 mt + g][d = 8 nt + 2 tq + e], summed over this warp's pairs
  double Sb[BD8];          // sum W k_b over this thread's pairs
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll
  for (int b = 0; b < BD8; ++b) Sb[b] = 0.0;

  const long nwork = grad_work_items(a);
  uint32_t phase = 0;
  for (long wk = blockIdx.x; wk < nwork; wk += gridDim.x) {
    int ti, tj;
    if (!grad_work_tile(a, wk, ti, tj)) continue;  // uniform over the CTA
    const int i0 = ti * T, j0 = tj * T;
    const bool diag_tile = (ti == tj);
    const double wt = diag_tile ? 1.0 : 2.0;

    __syncthreads();  // previous tile fully consumed before the TMA overwrites the staging tiles
    if (warp == 0) {
      if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)((2 * p + (KIND ? 2 : 4) * Bz + 2) * T * 8));
      __syncwarp();
      for (int c = lane; c < p; c += 32) {
        tma_bulk_g2s(Xi + c * T, a.X + i0 + (size_t)c * a.ldx, T * 8, bar);
        tma_bulk_g2s(Xj + c * T, a.X + j0 + (size_t)c * a.ldx, T * 8, bar);
      }
      for (int c = lane; c < Bz; c += 32) {
        tma_bulk_g2s(Zi + c * T, a.Z + i0 + (size_t)c * a.ldx, T * 8, bar);
        tma_bulk_g2s(Zj + c * T, a.Z + j0 + (size_t)c * a.ldx, T * 8, bar);
        if (KIND == 0) {
          tma_bulk_g2s(LZi + c * T, a.LZ + i0 + (size_t)c * a.ldx, T * 8, bar);
          tma_bulk_g2s(LZj + c * T, a.LZ + j0 + (size_t)c * a.ldx, T * 8, bar);
        }
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
    const double* kcol = a.Kinv + gi + (size_t)(j0 + jbase) * a.ld;
    double rowacc = 0.0;
    double knext = kcol[0];
#pragma unroll 1
    for (int jc = 0; jc < JCOLS; ++jc) {
      const int jj = jbase + jc;
      const int gj = j0 + jj;
      const double kinv = knext;
      if (jc + 1 < JCOLS) knext = kcol[(size_t)(jc + 1) * a.ld];
      const double alpha_j = aj[jj];
      const bool valid = (gi < a.n) && (gj < a.n);
      const double W = valid ? wt * (kinv - alpha_i * alpha_j) : 0.0;

      // ---- phase 1: D^2_d -> stage, distance sums E_c ---------------------------------------------
      double E[NE];
#pragma unroll
      for (int c = 0; c < NE; ++c) E[c] = 0.0;
      // d in chunks of 4 up to p rounded up to 4 (rolled: D^2_d goes to the stage, nothing else is indexed by d);
      // stage rows d >= p4 are never written and stay zero
#pragma unroll 1
      for (int d0 = 0; d0 < p4; d0 += 4) {
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          const int d = d0 + dd;
          const double df = Xi[d * T + li] - Xj[d * T + jj];
          const double d2 = df * df;
          Ds[d * LDS_ + lane] = d2;
          const double* wr = wt_s + d * WS;
#pragma unroll
          for (int c = 0; c < NE; c += 2) {
            if (c <= B) {  // uniform
              const double2 wv = *reinterpret_cast<const double2*>(wr + c);
              E[c] = fma(d2, wv.x, E[c]);
              if (c + 1 < NE) E[c + 1] = fma(d2, wv.y, E[c + 1]);
            }
          }
        }
      }
      if (KIND) {
#pragma unroll
        for (int c = 0; c < NE; ++c)
          if (c <= B) E[c] = fast_sqrt(E[c]);
      }
      // ---- phase 2: terms ---------------------------------------------------------------------------
      double kpart = 0.0;
#pragma unroll
      for (int b = 0; b < BD8; ++b) {
        if (b < B) {  // uniform
          double zi = 1.0, zj = 1.0, lzi = 0.0, lzj = 0.0;
          if (b > 0) {
            zi = Zi[(b - 1) * T + li];
            zj = Zj[(b - 1) * T + jj];
            if (KIND == 0) {
              lzi = LZi[(b - 1) * T + li];
              lzj = LZj[(b - 1) * T + jj];
            }
          }
          const double kv = term_value<KIND>(b, lam[b], E[b], zj, zi, lzj, lzi);
          kpart += kv;
          double tv = W * kv;
          Sb[b] += tv;
          if (KIND) tv *= fast_rcp(1.0 + SQRT3 * E[b + 1]);
          Ts[b * LDS_ + lane] = tv;
        }
      }
      rowacc = fma(kpart, alpha_j, rowacc);
      if (!diag_tile) {
        const double cpart = warp_sum(kpart * alpha_i);
        if (lane == 0) atomicAdd(a.Ka + gj, cpart);
      }
      __syncwarp();
      // ---- phase 3: S[b][d] += sum over the warp's 32 pairs of t_b * D^2_d  (FP64 tensor pipe) -------
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        double af[MT], bf[NT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) af[mt] = Ts[(8 * mt + g) * LDS_ + 4 * s + tq];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) bf[nt] = Ds[(8 * nt + g) * LDS_ + 4 * s + tq];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
      }
      __syncwarp();
    }
    atomicAdd(a.Ka + gi, rowacc);
  }

  // ---- CTA reduction: per-warp fragments / lane sums -> smem -> one partial row per CTA -------------
  constexpr int NV = BD8 * PD8 + BD8;
  __syncthreads();           // everybody is done with the stage buffers; reuse them as red[8][NV]
  double* red = stage;       // NWARPS * NV <= NWARPS * (BD8 + PD8) * 36 for every instantiated shape (checked on host)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) red[warp * NV + (8 * mt + g) * PD8 + 8 * nt + 2 * tq + e] = acc[mt][nt][e];
#pragma unroll
  for (int b = 0; b < BD8; ++b) {
    const double v = warp_sum(Sb[b]);
    if (lane == 0) red[warp * NV + BD8 * PD8 + b] = v;
  }
  __syncthreads();
  double* out = a.partials + (size_t)blockIdx.x * a.P;
  for (int idx = threadIdx.x; idx < a.P; idx += NTHR) out[idx] = 0.0;
  __syncthreads();
  for (int r = threadIdx.x; r < NV; r += NTHR) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) v += red[w * NV + r];
    if (r < BD8 * PD8) {
      const int b = r / PD8, d = r % PD8;
      if (d < p && b < B) out[2 + B + b + B * d] = v;
    } else {
      const int b = r - BD8 * PD8;
      if (b < B) out[2 + b] = v;
    }
  }
}

}  // namespace ace
