// dgemm_nt.cuh -- FP64 tensor-core (DMMA) GEMM  C = beta*C + alpha * A * B^T  for sm_100a.
//
// This one kernel is the dense engine of the whole path: the SYRK/GEMM trailing updates and the
// TRSM-as-GEMM of the blocked Cholesky, both triangular products of the triangular inverse, the
// U*U^T product that forms (K + e^sigma I)^-1, and the K_xX * K^-1 product of the posterior.
// It replaces LAPACK dsyevd + BLAS dsyrk/dgemm behind arma::eig_sym / operator* in the reference
// (src/kernel_SE_cpp.cpp:144,153; src/pred_cpp.cpp:19,22,69,72).
//
// Layout: everything column-major (R/Armadillo).  A is M x K, B is N x K ("NT": both operands are
// contiguous along their long edge), so a BK-deep operand chunk is BK contiguous column segments,
// each fetched by ONE 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) into a padded smem row.
// Padding (ld = 128+4 / 64+4 doubles) makes the DMMA fragment loads bank-conflict free.
//
// CTA = 128 x 64 output tile, 4 consumer warps (2 x 2, each 64 x 32 = 8 x 4 DMMA.8x8x4 tiles, 64
// FP64 accumulators per thread) + 1 producer warp; 3-stage mbarrier full/empty ring; 2 CTAs per SM
// (77 KB smem each -- a third slot of >= 138 KB stays free for the latency-critical leaf kernels of the
// Cholesky look-ahead stream --, <= 200 regs) so one CTA's C read-modify-write epilogue hides under the other's
// main loop.  FP64 has no tcgen05/TMEM kind on Blackwell: DMMA via mma.sync is the FP64 tensor path.
//
// All of M, N, K are multiples of 128 (callers pad to the 128-tile grid with identity), so there
// is no edge predication anywhere in the kernel.
#pragma once
#include "common.cuh"

namespace ace {

struct GemmNT {
  const double* A;      // M x K, element (i,k) at A[i + k*lda]
  long lda;
  const double* B;      // N x K, element (j,k) at B[j + k*ldb]
  long ldb;
  double* C;            // M x N
  long ldc;
  double* Ct;           // optional transposed copy: Ct[j + i*ldct] = C(i,j); may be nullptr
  long ldct;
  const double* Adiag;  // optional: dense 128x128 tiles (ld 128, tile t at +t*16384) that replace the
  const double* Bdiag;  //           diagonal 128-blocks (row block == k block) of A / B
  int M, N, K;
  int a_tri;            // 0: A dense.  1: A upper triangular on the 128-block grid (k starts at the row
                        //    block).  2: A lower triangular (k ends with the row block).
  int b_tri;            // 2: B lower triangular on the 128-block grid (k ends with the column's block)
  int lower_only;       // 1: C is square, only tiles that intersect i >= j are computed (SYRK)
  double alpha, beta;
  // multi-GPU sharding of ONE product: this launch computes the tiles L = tile_first + q * tile_stride of the
  // linear tile order (tile_stride 0 or 1: all tiles), so G ranks with tile_first = rank cover the output
  // exactly once with a balanced share of the triangular k-ranges
  int tile_first, tile_stride;
  // A is a row slice starting at row a_row_off (multiple of 128) of the triangular operand: the k trimming and
  // the diagonal-tile redirect use the row block index ti + a_row_off/128 (Adiag stays the unshifted tile array)
  int a_row_off;
  int s_row_off;        // batched: problem q uses a_row_off + q * s_row_off
  // strided batch (blockIdx.y): problem q uses every pointer advanced by q * its stride (elements)
  int batch;            // 0 or 1: single problem
  long sA, sB, sC, sCt, sAdiag, sBdiag;
  int m_dec;            // batched problems of shrinking height: problem q has M - q * m_dec rows (tiles beyond exit)
  // split-K: batch entry q = problem (q / ksplit), K chunk (q % ksplit) = absolute k in [chunk * klen, (chunk+1) * klen);
  // A, B, Adiag, Bdiag, a_row_off advance with the problem, C / Ct with the entry (partial results, summed by the caller)
  int ksplit, klen;
};

namespace gemm {
constexpr int BM = 128, BN = 64, BK = 16, STAGES = 3;
constexpr int LDAS = BM + 4, LDBS = BN + 4;
constexpr int A_STAGE = BK * LDAS, B_STAGE = BK * LDBS;  // doubles
constexpr int CONSUMER_WARPS = 4;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr uint32_t STAGE_TX_BYTES = BK * (BM + BN) * 8;
constexpr size_t SMEM_BYTES = (size_t)STAGES * (A_STAGE + B_STAGE) * 8 + 2 * STAGES * 8;
constexpr int GROUP = 8;  // tile-rows per rasterisation group (L2 reuse of B panels)
}  // namespace gemm

__global__ void __launch_bounds__(gemm::THREADS, 2) dgemm_nt_kernel(const GemmNT p) {
  using namespace gemm;
  // strided batch: every base pointer of problem q = blockIdx.y (plain locals; the parameter struct is
  // never modified)
  const long qe = blockIdx.y;                                    // batch entry
  const long q = (p.ksplit > 1) ? qe / p.ksplit : qe;            // problem
  const int kchunk = (p.ksplit > 1) ? (int)(qe % p.ksplit) : 0;  // its K chunk
  const double* const gA = p.A + q * p.sA;
  const double* const gB = p.B + q * p.sB;
  double* const gC = p.C + qe * p.sC;
  double* const gCt = (p.Ct != nullptr) ? p.Ct + qe * p.sCt : nullptr;
  const double* const gAdiag = (p.Adiag != nullptr) ? p.Adiag + q * p.sAdiag : nullptr;
  const double* const gBdiag = (p.Bdiag != nullptr) ? p.Bdiag + q * p.sBdiag : nullptr;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);
  double* Bs = As + STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(Bs + STAGES * B_STAGE);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile decode -------------------------------------------------------------------------
  int ti, tj;
  {
    const long L = (p.tile_stride > 1) ? (long)blockIdx.x * p.tile_stride + p.tile_first : (long)blockIdx.x;
    if (p.lower_only) {
      // row ti holds tiles tj = 0 .. 2*ti+1 ; rows before it hold ti*(ti+1) tiles
      long t = (long)((sqrt(4.0 * (double)L + 1.0) - 1.0) * 0.5);
      while (t * (t + 1) > L) --t;
      while ((t + 1) * (t + 2) <= L) ++t;
      ti = (int)t;
      tj = (int)(L - t * (t + 1));
    } else {
      const int tiles_m = p.M / BM, tiles_n = p.N / BN;
      const long per_group = (long)GROUP * tiles_n;
      const int g = (int)(L / per_group);
      const int first_m = g * GROUP;
      const int gsize = min(GROUP, tiles_m - first_m);
      const int r = (int)(L % per_group);
      ti = first_m + r % gsize;
      tj = r / gsize;
    }
  }
  const int i0 = ti * BM, j0 = tj * BN;
  if (i0 >= p.M - (int)q * p.m_dec) return;  // whole CTA, before any barrier exists
  int k_begin = 0, k_end = p.K;
  const int a_off = p.a_row_off + (int)q * p.s_row_off;
  const int tia = ti + a_off / TB;  // row block of this tile inside the (possibly sliced) A operand
  if (p.a_tri == 1) k_begin = i0 + a_off;
  if (p.a_tri == 2) k_end = i0 + a_off + BM;
  if (p.b_tri == 2) k_end = min(k_end, (j0 / TB + 1) * TB);
  if (p.ksplit > 1) {
    k_begin = max(k_begin, kchunk * p.klen);
    k_end = min(k_end, (kchunk + 1) * p.klen);
  }
  const int nchunks = (k_end - k_begin) / BK;  // <= 0: empty range, the tile stores beta * C (zeros for beta = 0)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMER_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ================================ TMA producer warp ======================================
    const int kk = lane & 15;
    const bool isA = lane < 16;
    const int jb = j0 / TB, joff = j0 % TB;
    for (int it = 0; it < nchunks; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full[s], STAGE_TX_BYTES);
      __syncwarp();
      const int k = k_begin + it * BK + kk;
      const int kb = k / TB;
      if (isA) {
        const double* src = (gAdiag != nullptr && kb == tia)
                                ? gAdiag + (size_t)tia * (TB * TB) + (size_t)(k % TB) * TB
                                : gA + i0 + (size_t)k * p.lda;
        tma_bulk_g2s(As + s * A_STAGE + kk * LDAS, src, BM * 8, &full[s]);
      } else {
        const double* src = (gBdiag != nullptr && kb == jb)
                                ? gBdiag + (size_t)jb * (TB * TB) + (size_t)(k % TB) * TB + joff
                                : gB + j0 + (size_t)k * p.ldb;
        tma_bulk_g2s(Bs + s * B_STAGE + kk * LDBS, src, BN * 8, &full[s]);
      }
    }
    return;
  }

  // ================================== DMMA consumer warps =====================================
  const int wm = warp >> 1, wn = warp & 1;  // 2 x 2 warps, warp tile 64 x 32
  const int g = lane >> 2, tq = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int mt = 0; mt < 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  for (int it = 0; it < nchunks; ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const double* as = As + s * A_STAGE + wm * 64 + g;
    const double* bs = Bs + s * B_STAGE + wn * 32 + g;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int k = ks * 4 + tq;
      double a[8], b[4];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) a[mt] = as[k * LDAS + mt * 8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b[nt] = bs[k * LDBS + nt * 8];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ---- epilogue: C = beta*C + alpha*acc, optional transposed copy ---------------------------
  const double alpha = p.alpha, beta = p.beta;
  const int row0 = i0 + wm * 64 + g;
  const int col0 = j0 + wn * 32 + 2 * tq;
  if (beta != 0.0) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double* cc = gC + (size_t)(col0 + nt * 8 + e) * p.ldc + row0;
        double old[8];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) old[mt] = cc[mt * 8];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) acc[mt][nt][e] = fma(beta, old[mt], alpha * acc[mt][nt][e]);
      }
  } else {
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[mt][nt][0] *= alpha;
        acc[mt][nt][1] *= alpha;
      }
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double* cc = gC + (size_t)(col0 + nt * 8 + e) * p.ldc + row0;
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) cc[mt * 8] = acc[mt][nt][e];
    }
  if (p.Ct != nullptr) {
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        double2 v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<double2*>(gCt + (size_t)(row0 + mt * 8) * p.ldct + col0 + nt * 8) = v;
      }
  }
}

// Host launcher.  Returns 0 or a negative CUDA status.
inline int launch_gemm_nt(const GemmNT& p, cudaStream_t stream) {
  using namespace gemm;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
  if (p.M % BM || p.N % TB || p.K % TB) {
    set_error("launch_gemm_nt: dimensions must be multiples of 128");
    return -2;
  }
  long tiles;
  if (p.lower_only) {
    const long T = p.M / BM;
    tiles = T * (T + 1);
  } else {
    tiles = (long)(p.M / BM) * (p.N / BN);
  }
  if (p.tile_stride > 1) {
    if (p.tile_first < 0 || p.tile_first >= p.tile_stride || p.batch > 1) {
      set_error("launch_gemm_nt: bad tile subset");
      return -2;
    }
    tiles = (tiles - p.tile_first + p.tile_stride - 1) / p.tile_stride;
    if (tiles <= 0) return 0;
  }
  if (p.a_row_off % TB || p.s_row_off % TB) {
    set_error("launch_gemm_nt: a_row_off must be a multiple of 128");
    return -2;
  }
  if (p.ksplit > 1 && (p.klen % TB || p.klen <= 0)) {
    set_error("launch_gemm_nt: split-K chunk must be a positive multiple of 128");
    return -2;
  }
  dim3 grid((unsigned)tiles, (unsigned)((p.batch > 1 ? p.batch : 1) * (p.ksplit > 1 ? p.ksplit : 1)));
  dgemm_nt_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(p);
  ACE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ace
