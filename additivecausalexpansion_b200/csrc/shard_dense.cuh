// shard_dense.cuh -- the dense part of ONE fit spread over the G GPUs of a node.
//
// All three n^3/3 phases are sharded; every rank ends with the complete factor L, U = L^-T, X = L^-1:
//
//  * triangular inverse, default while the panel chain bounds the Cholesky (potrf_sharded + gather_inverse_sharded):
//    block row J of X for a column panel c only needs X[c0:j0, c], i.e. the rank's OWN columns, so every rank grows
//    X[:, its panels] behind the panels on a background stream (batched split-K merges) with no exchange; one padded
//    ncclAllGather at the end distributes the column panels.
//  * triangular inverse, otherwise (trtri_merge_sharded): the bottom-up merge tree of chol.cuh is kept; its LOW levels (small nodes, ~6 % of the
//    flops) run redundantly on every rank, the HIGH levels are split.  For a node with children [a,c), [c,b) the
//    rows of U12 = (X21)^T = -(U11 L21^T) X22^T are independent, so the h*128 rows are cut into 2G slices and
//    rank r computes slices r and 2G-1-r (U11 is triangular: the pairing balances the k-ranges exactly).  The
//    slices are written packed into a staging buffer, exchanged with ONE ncclAllGather per level over NVLink,
//    and unpacked into upper(A) (and, transposed, into lower(A) when a higher level still needs X22).
//  * K^-1 = U U^T: the output tiles of the single lower-triangular launch are dealt round-robin (tile L to rank
//    L mod G).  No exchange follows: the gradient pass uses exactly the same ownership (GradArgs.tile_mode 1),
//    alpha = U (U^T ybar) and diag(K^-1) come from triangular matrix-vector products with the replicated U.
//
//  * Cholesky (potrf_sharded): right-looking over 512-wide column panels dealt cyclically (panel J to rank
//    J mod G).  The owner factors the diagonal block, inverts it and solves the rows below with one triangular
//    GEMM into a PACKED buffer ((n - j0) x 512, contiguous), which goes to everybody with one grouped
//    ncclBroadcast (panel + DX/DU tiles + diagonal); every rank applies the panel to the panels it owns, the
//    owner of panel J+1 first and on the high-priority stream (look-ahead), and copies it into its own A off the
//    critical path, so that all ranks end up with the complete L.  No rank ever needs K columns it does not own:
//    the kernel build is sharded the same way and its exchange disappears.
//
// `emulate`: a single process plays all G ranks one after the other on one GPU (no NCCL): the same world-strided
// batched launches, tile subsets and packed buffers a real rank uses, with the exchanges replaced by the shared
// memory of the one device (so the collectives themselves, and the cross-stream ordering around them, are only
// exercised by the real 2-rank test, tests/test_gpu_shard.py).  Lets the single-GPU test tier cover the indexing.
#pragma once
#include <cstdio>
#include "chol.cuh"
#include "nccl_dyn.h"

namespace ace {

struct ShardCtx {
  int rank = 0, world = 1;
  bool emulate = false;
  ncclComm_t comm = nullptr;
  int h_min = 16;  // levels with child size h >= h_min (in 128-blocks) and h % (2*world) == 0 are split
  ncclComm_t comm2 = nullptr;       // second communicator: bulk panel broadcasts (never queue behind the heads)
  cudaEvent_t* events = nullptr;    // SHARD_EVENT_KINDS * panels events of potrf_sharded
  double* head[2] = {nullptr, nullptr};  // packed panel heads, (2 * panel width) x (panel width) doubles each
  cudaStream_t bulk_stream = nullptr;
  // third piece of a panel: the FIRST block of the bulk, L[J+2 rows, J], sent early on its own stream and communicator
  // (see potrf_sharded); packed (panel width)^2 doubles each
  ncclComm_t comm3 = nullptr;
  double* mid[2] = {nullptr, nullptr};
  cudaStream_t mid_stream = nullptr;
  double* chain_ws = nullptr;  // split-K partials of the chain's small GEMMs (chain_gemm): 4 x (panel width)^2 doubles
  bool incr = false;  // incremental inverse behind the panels: every rank grows X[:, its column panels] in the background
  bool mine(int panel) const { return emulate || panel % world == rank; }
};

inline int shard_panels(int nb, int pb) { return (nb + pb - 1) / pb; }

// out[e] = sum_s part[(prob * S + s) * len + e]: sums the S split-K partials of every problem (len elements each)
__global__ void __launch_bounds__(256) splitk_sum_kernel(const double* __restrict__ part, int S, size_t len, size_t total,
                                                         double* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const size_t prob = idx / len, e = idx % len;
  const double* src = part + prob * (size_t)S * len + e;
  double acc = src[0];
  for (int s2 = 1; s2 < S; ++s2) acc += src[(size_t)s2 * len];
  out[idx] = acc;
}

// C (ld ldc) = beta * C + sum of the S split-K partials (each R x Cc, packed)
__global__ void __launch_bounds__(256) splitk_sum_ld_kernel(const double* __restrict__ part, int S, int R, int Cc,
                                                            double* __restrict__ C, long ldc, double beta) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, len = (size_t)R * Cc;
  if (idx >= len) return;
  double acc = part[idx];
  for (int s2 = 1; s2 < S; ++s2) acc += part[(size_t)s2 * len + idx];
  double* dst = C + (idx % R) + (idx / R) * ldc;
  *dst = (beta != 0.0) ? fma(beta, *dst, acc) : acc;
}

// A small GEMM of the serial panel chain (head TRSM, diagonal-block update: 512^3) next to the trailing updates of
// other panels: its 32 tiles of 128 x 64 x 512 share their SMs with a resident update CTA and take ~70 us each, 0.18 ms
// per launch on the chain (profiles/r02/shard_trace_w8.log).  Split over K into chunks of 128 the 128 shorter CTAs
// finish in a quarter of that; the partials are summed into C by one small kernel.
inline int chain_gemm(GemmNT g, double* ws, cudaStream_t st) {
  const int S = g.K / TB;
  if (ws == nullptr || S <= 1 || g.K % TB || g.batch > 1 || g.ksplit > 1) return launch_gemm_nt(g, st);
  double* C = g.C;
  const long ldc = g.ldc;
  const double beta = g.beta;
  g.C = ws; g.ldc = g.M; g.beta = 0.0;
  g.ksplit = S; g.klen = TB; g.sA = 0; g.sB = 0; g.sC = (long)g.M * g.N;
  ACE_TRY(launch_gemm_nt(g, st));
  const size_t len = (size_t)g.M * g.N;
  splitk_sum_ld_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(ws, S, g.M, g.N, C, ldc, beta);
  ACE_CUDA(cudaGetLastError());
  return 0;
}

constexpr int SHARD_EVENT_KINDS = 8;

// ACE_SHARD_TRACE=1: per-panel timeline of potrf_sharded (CUDA events with timing), printed by shard_trace_dump
struct ShardTrace {
  cudaEvent_t t0 = nullptr;
  std::vector<cudaEvent_t> ev;  // 14 per panel: s0 s1 s2 s3 | b1 b2 b3 | m0 m1 | mid: begin trsm bcast applied | diag kernel done
  int np = 0;
  std::vector<cudaEvent_t> lev;   // triangular inverse: 5 per level (begin, gemm1, gemm2, gather, unpack)
  std::vector<int> lev_h;
  void level_mark(int h, int k, cudaStream_t st) {
    if (!t0) return;
    if (k == 0) lev_h.push_back(h);
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    lev.push_back(e);
  }
  void mark(int J, int k, cudaStream_t st) {
    if (!t0) return;
    cudaEventRecord(ev[(size_t)J * 14 + k], st);
  }
};
inline ShardTrace& shard_trace() {
  static ShardTrace t;
  return t;
}
inline void shard_trace_begin(int NP, cudaStream_t st) {
  ShardTrace& t = shard_trace();
  static const bool on = std::getenv("ACE_SHARD_TRACE") != nullptr;
  if (!on) return;
  if (!t.t0) cudaEventCreate(&t.t0);
  while ((int)t.ev.size() < 14 * NP) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    t.ev.push_back(e);
  }
  t.np = NP;
  for (auto e : t.lev) cudaEventDestroy(e);
  t.lev.clear();
  t.lev_h.clear();
  cudaEventRecord(t.t0, st);
}
inline void shard_trace_dump(int rank) {
  ShardTrace& t = shard_trace();
  if (!t.t0) return;
  cudaDeviceSynchronize();
  std::fprintf(stderr, "[shard trace rank %d] J: factor_begin head_ready head_bcast diag_applied | bulk_trsm bulk_bcast la_applied | main_begin main_end | mid_begin mid_trsm mid_bcast mid_applied | diag_kernel_done (ms)\n", rank);
  for (int J = 0; J < t.np; ++J) {
    float v[14];
    for (int k = 0; k < 14; ++k)
      if (cudaEventElapsedTime(&v[k], t.t0, t.ev[(size_t)J * 14 + k]) != cudaSuccess) v[k] = -1.f;
    std::fprintf(stderr, "[shard trace rank %d] %2d: %7.3f %7.3f %7.3f %7.3f | %7.3f %7.3f %7.3f | %7.3f %7.3f | %7.3f %7.3f %7.3f %7.3f | %7.3f\n",
                 rank, J, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13]);
  }
  for (size_t l = 0; l < t.lev_h.size(); ++l) {
    float v[5];
    for (int k = 0; k < 5; ++k)
      if (cudaEventElapsedTime(&v[k], t.t0, t.lev[l * 5 + k]) != cudaSuccess) v[k] = -1.f;
    std::fprintf(stderr, "[shard trace rank %d] trtri level h=%d (%s): begin %.3f gemm1 +%.3f gemm2 +%.3f gather +%.3f unpack +%.3f\n",
                 rank, t.lev_h[l] < 0 ? -t.lev_h[l] : t.lev_h[l], t.lev_h[l] < 0 ? "redundant" : "split", v[0], v[1] - v[0],
                 v[2] - v[1], v[3] - v[2], v[4] - v[3]);
  }
  cudaGetLastError();
}

// Phase 1 for a sharded fit.  Needs w.Wp[0..1] (packed bulk of a panel), cx.head[0..1] (packed head), w.Wsmall.
//
// A panel travels in two pieces so that the serial chain of the factorisation only carries what it must:
//   head(J) = diagonal block (factor + its inverse) and the rows of the NEXT panel, L[J+1 rows, J]: all the owner
//             of panel J+1 needs to finish and factor ITS diagonal block.  Small (<= 1024 x 512), own communicator,
//             high-priority stream -> chain per panel = diagonal factorisation + one small GEMM + one small broadcast.
//   bulk(J) = the rows below, needed only when the next head is solved, i.e. one diagonal factorisation later; own
//             stream and communicator, so the two transfers never queue behind each other.
//   mid(J)  = the first block of the bulk, L[J+2 rows, J], sent a second time and EARLY (third stream and
//             communicator): it is all that head(J+1) needs of bulk(J) (the update of A[J+2 rows, J+1]).  Without it
//             the whole bulk transfer (46 MB at panel 8 of C3, 0.3 ms) and its look-ahead GEMM sat on the serial chain:
//             0.75 ms per panel on 8 GPUs against 0.25 ms for diagonal block + head (profiles/r02/shard_trace_w8.log).
//             mid(J) itself needs the look-ahead of bulk(J-1), so the bulk path now has a whole panel period of slack.
inline int potrf_sharded(const DenseWork& w, const ShardCtx& cx) {
  const int nb = w.nb, pb = w.panel_blocks, NP = shard_panels(nb, pb);
  if (!w.Wp[0] || !cx.events || !cx.head[0] || !cx.bulk_stream) {
    set_error("potrf_sharded: packed panel buffers missing");
    return -2;
  }
  cudaEvent_t* ev_head = cx.events;            // side : head(J) available in head[J & 1]
  cudaEvent_t* ev_bulk = cx.events + NP;       // bulk : bulk(J) available in Wp[J & 1]
  cudaEvent_t* ev_la = cx.events + 2 * NP;     // bulk : panel J fully applied to panel J+1 (look-ahead, rows below its diagonal block)
  cudaEvent_t* ev_first = cx.events + 3 * NP;  // main : panel J applied to my soonest-needed panel
  cudaEvent_t* ev_step = cx.events + 4 * NP;   // main : panel J applied to all my panels
  cudaEvent_t* ev_copy = cx.events + 5 * NP;   // aux  : panel J copied into A
  cudaEvent_t* ev_diag = cx.events + 6 * NP;   // side : head(J) applied to the diagonal block of panel J+1
  cudaEvent_t* ev_midla = cx.events + 7 * NP;  // mid  : mid(J) applied to A[J+2 rows, J+1] (look-ahead, first block)
  const bool use_mid = cx.mid[0] != nullptr && cx.mid_stream != nullptr && (cx.emulate || cx.comm3 != nullptr);
  cudaStream_t mids = cx.mid_stream;
  NcclApi& nc = nccl_api();
  cudaStream_t bulk = cx.bulk_stream;
  ACE_CUDA(cudaMemsetAsync(w.info, 0, sizeof(int), w.main));
  ACE_CUDA(cudaEventRecord(w.ev_upd[1], w.main));  // fork: the other streams join after everything queued on main
  ACE_CUDA(cudaStreamWaitEvent(w.side, w.ev_upd[1], 0));
  ACE_CUDA(cudaStreamWaitEvent(bulk, w.ev_upd[1], 0));
  if (use_mid) ACE_CUDA(cudaStreamWaitEvent(mids, w.ev_upd[1], 0));
  ACE_CUDA(cudaStreamWaitEvent(w.aux, w.ev_upd[1], 0));
  shard_trace_begin(NP, w.main);
  ShardTrace& tr = shard_trace();
  for (int J = 0; J < NP; ++J) {
    const int j0 = J * pb, j1 = std::min(j0 + pb, nb), j2 = std::min(j1 + pb, nb);
    const long wJ = (long)(j1 - j0) * TB, wN = (long)(j2 - j1) * TB;  // panel width, next panel's width
    const long hJ = wJ + wN, mB = (long)(nb - j2) * TB;               // head rows, bulk rows
    const int j3 = std::min(j2 + pb, nb);
    const long wM = use_mid ? (long)(j3 - j2) * TB : 0;               // mid rows (first block of the bulk)
    double* Hd = cx.head[J & 1];
    double* Wp = w.Wp[J & 1];
    double* Md = cx.mid[J & 1];
    const bool mine = cx.mine(J), la = (J + 1 < NP) && cx.mine(J + 1);
    // ================= head(J): side stream, communicator `comm`
    if (J >= 2) {  // head buffer free: its readers were side (in order), the bulk look-ahead and the copy-out of J-2
      ACE_CUDA(cudaStreamWaitEvent(w.side, ev_la[J - 2], 0));
      if (use_mid) ACE_CUDA(cudaStreamWaitEvent(w.side, ev_midla[J - 2], 0));
      ACE_CUDA(cudaStreamWaitEvent(w.side, ev_copy[J - 2], 0));
    }
    tr.mark(J, 0, w.side);
    if (mine) {
      ACE_TRY(diag_block_factor_invert(w, j0, j1, w.side, cx.world));  // L_JJ, X_JJ / U_JJ in place
      tr.mark(J, 13, w.side);
      ACE_CUDA(cudaMemcpy2DAsync(Hd, sizeof(double) * hJ, blkptr(w, j0, j0), sizeof(double) * w.ld,
                                 sizeof(double) * wJ, (size_t)wJ, cudaMemcpyDeviceToDevice, w.side));
      if (wN > 0) {
        // the next panel's rows of my panel are up to date: that block received panel J-1 through mid(J-1)
        if (J >= 1) ACE_CUDA(cudaStreamWaitEvent(w.side, use_mid ? ev_midla[J - 1] : ev_la[J - 1], 0));
        GemmNT t{};
        t.A = blkptr(w, j1, j0); t.lda = w.ld;
        t.B = blkptr(w, j0, j0); t.ldb = w.ld; t.b_tri = 2; t.Bdiag = w.DX + (size_t)j0 * TB * TB;
        t.C = Hd + wJ; t.ldc = hJ;
        t.M = (int)wN; t.N = (int)wJ; t.K = (int)wJ; t.alpha = 1.0; t.beta = 0.0;
        ACE_TRY(chain_gemm(t, cx.chain_ws, w.side));
      }
    }
    tr.mark(J, 1, w.side);
    if (!cx.emulate) {
      const int root = J % cx.world;
      double* dx = w.DX + (size_t)j0 * TB * TB;
      double* du = w.DU + (size_t)j0 * TB * TB;
      double* dv = w.dvec + (size_t)j0 * TB;
      ACE_NCCL(nc.GroupStart());
      ACE_NCCL(nc.Broadcast(Hd, Hd, (size_t)hJ * wJ, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.Broadcast(dx, dx, (size_t)wJ * TB, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.Broadcast(du, du, (size_t)wJ * TB, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.Broadcast(dv, dv, (size_t)wJ, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.GroupEnd());
    }
    ACE_CUDA(cudaEventRecord(ev_head[J], w.side));
    tr.mark(J, 2, w.side);
    if (la) {  // diagonal block of the next panel: A[J+1, J+1] -= Hn Hn^T with Hn = L[J+1 rows, J]
      if (J >= 1) ACE_CUDA(cudaStreamWaitEvent(w.side, ev_first[J - 1], 0));  // panels <= J-1 already applied to it
      GemmNT g{};
      g.A = Hd + wJ; g.lda = hJ; g.B = Hd + wJ; g.ldb = hJ; g.C = blkptr(w, j1, j1); g.ldc = w.ld;
      g.M = (int)wN; g.N = (int)wN; g.K = (int)wJ; g.alpha = -1.0; g.beta = 1.0;
      ACE_TRY(chain_gemm(g, cx.chain_ws, w.side));
    }
    ACE_CUDA(cudaEventRecord(ev_diag[J], w.side));
    tr.mark(J, 3, w.side);
    // ================= bulk(J): rows [j2, nb), bulk stream, communicator `comm2`
    if (J >= 2) {  // Wp[J & 1] free: readers were main (step J-2), the bulk stream itself and the copy-out
      ACE_CUDA(cudaStreamWaitEvent(bulk, ev_step[J - 2], 0));
      ACE_CUDA(cudaStreamWaitEvent(bulk, ev_copy[J - 2], 0));
    }
    ACE_CUDA(cudaStreamWaitEvent(bulk, ev_head[J], 0));  // X_JJ (owner) / head data (look-ahead apply below)
    if (mB > 0) {
      if (mine) {
        GemmNT t{};
        t.A = blkptr(w, j2, j0); t.lda = w.ld;
        t.B = blkptr(w, j0, j0); t.ldb = w.ld; t.b_tri = 2; t.Bdiag = w.DX + (size_t)j0 * TB * TB;
        t.C = Wp; t.ldc = mB;
        t.M = (int)mB; t.N = (int)wJ; t.K = (int)wJ; t.alpha = 1.0; t.beta = 0.0;
        ACE_TRY(launch_gemm_nt(t, bulk));
      }
      tr.mark(J, 4, bulk);
      if (!cx.emulate) ACE_NCCL(nc.Broadcast(Wp, Wp, (size_t)mB * wJ, ncclFloat64, J % cx.world, cx.comm2, bulk));
    } else {
      tr.mark(J, 4, bulk);
    }
    ACE_CUDA(cudaEventRecord(ev_bulk[J], bulk));
    tr.mark(J, 5, bulk);
    if (la && mB > wM) {  // rows below the next panel's diagonal block (and below mid): A[j3:, J+1] -= bulk * Hn^T
      if (J >= 1) ACE_CUDA(cudaStreamWaitEvent(bulk, ev_first[J - 1], 0));
      GemmNT g{};
      g.A = Wp + wM; g.lda = mB; g.B = Hd + wJ; g.ldb = hJ; g.C = blkptr(w, j2, j1) + wM; g.ldc = w.ld;
      g.M = (int)(mB - wM); g.N = (int)wN; g.K = (int)wJ; g.alpha = -1.0; g.beta = 1.0;
      ACE_TRY(launch_gemm_nt(g, bulk));
    }
    ACE_CUDA(cudaEventRecord(ev_la[J], bulk));
    tr.mark(J, 6, bulk);
    // ================= mid(J): rows [j2, j3), mid stream, communicator `comm3`
    if (use_mid) {
      ACE_CUDA(cudaStreamWaitEvent(mids, ev_head[J], 0));  // X_JJ (owner) / head data (look-ahead apply below)
      tr.mark(J, 9, mids);
      if (wM > 0) {
        if (mine) {
          // these rows of my panel received panel J-1 with the rest of its bulk (and everything older before that)
          if (J >= 1) ACE_CUDA(cudaStreamWaitEvent(mids, ev_la[J - 1], 0));
          GemmNT t{};
          t.A = blkptr(w, j2, j0); t.lda = w.ld;
          t.B = blkptr(w, j0, j0); t.ldb = w.ld; t.b_tri = 2; t.Bdiag = w.DX + (size_t)j0 * TB * TB;
          t.C = Md; t.ldc = wM;
          t.M = (int)wM; t.N = (int)wJ; t.K = (int)wJ; t.alpha = 1.0; t.beta = 0.0;
          ACE_TRY(launch_gemm_nt(t, mids));
        }
        tr.mark(J, 10, mids);
        if (!cx.emulate) ACE_NCCL(nc.Broadcast(Md, Md, (size_t)wM * wJ, ncclFloat64, J % cx.world, cx.comm3, mids));
        tr.mark(J, 11, mids);
        if (la) {  // A[j2:j3, J+1] -= mid * Hn^T
          if (J >= 1) ACE_CUDA(cudaStreamWaitEvent(mids, ev_first[J - 1], 0));
          GemmNT g{};
          g.A = Md; g.lda = wM; g.B = Hd + wJ; g.ldb = hJ; g.C = blkptr(w, j2, j1); g.ldc = w.ld;
          g.M = (int)wM; g.N = (int)wN; g.K = (int)wJ; g.alpha = -1.0; g.beta = 1.0;
          ACE_TRY(launch_gemm_nt(g, mids));
        }
      }
      ACE_CUDA(cudaEventRecord(ev_midla[J], mids));
      tr.mark(J, 12, mids);
    }
    // ================= aux stream: the panel into this rank's A (the owner already has its diagonal block)
    ACE_CUDA(cudaStreamWaitEvent(w.aux, ev_head[J], 0));
    ACE_CUDA(cudaStreamWaitEvent(w.aux, ev_bulk[J], 0));
    if (use_mid) ACE_CUDA(cudaStreamWaitEvent(w.aux, ev_midla[J], 0));  // mid's TRSM has read A[j2:j3, J]
    {
      const long skip = mine ? wJ : 0;
      if (hJ > skip)
        ACE_CUDA(cudaMemcpy2DAsync(blkptr(w, j0, j0) + skip, sizeof(double) * w.ld, Hd + skip, sizeof(double) * hJ,
                                   sizeof(double) * (hJ - skip), (size_t)wJ, cudaMemcpyDeviceToDevice, w.aux));
      if (mB > 0)
        ACE_CUDA(cudaMemcpy2DAsync(blkptr(w, j2, j0), sizeof(double) * w.ld, Wp, sizeof(double) * mB,
                                   sizeof(double) * mB, (size_t)wJ, cudaMemcpyDeviceToDevice, w.aux));
    }
    ACE_CUDA(cudaEventRecord(ev_copy[J], w.aux));
    // ================= background stream: rows J of the inverse for MY column panels c < J,
    //   X[J, c] = -X_JJ * ( L[J, c0:j0] * X[c0:j0, c] ),
    // as two batched GEMMs over the owned panels (Wt_c = U[c rows, c0:j0] * L[J, c0:j0]^T; X[J,c] = -X_JJ Wt_c^T,
    // with the transposed copy U[c rows, J]).  A rank only ever needs its own columns of X: no exchange until the end.
    if (cx.incr && J >= 1 && w.bg) {
      // Emulation plays every rank in turn with the SAME world-strided batches a real rank launches.  All first
      // products run before any second one: a real rank overwrites L[J, its panels] with X in its own copy of A,
      // here the one shared A must keep block row J of L until every played rank has read it.
      const int Gs = cx.world;
      const int r_lo = cx.emulate ? 0 : cx.rank, r_hi = cx.emulate ? cx.world : cx.rank + 1;
      const long pw = (long)pb * TB, step = (long)Gs * pw;
      // few tiles (32 per problem) with a K range that grows to n: split K into chunks of <= 2048 so that the
      // background work is spread over the SMs instead of running as a handful of millisecond-long tiles
      const int KL = 2048;
      const int S = std::max(1, (j0 * TB + KL - 1) / KL);
      const size_t len = (size_t)pw * wJ;
      auto count_of = [&](int cf) { return (cf < J) ? (J - cf + Gs - 1) / Gs : 0; };
      size_t total_cnt = 0;
      for (int cf = r_lo; cf < r_hi; ++cf) total_cnt += (size_t)count_of(cf);
      ACE_CUDA(cudaStreamWaitEvent(w.bg, ev_copy[J], 0));  // block row J of L and X_JJ are in A (copies are in order)
      size_t off = 0;
      for (int cf = r_lo; cf < r_hi; ++cf) {
        const int cnt = count_of(cf);
        if (cnt <= 0) continue;
        double* Wt = w.Bf + off * len;
        double* part = (S > 1) ? w.Bf + total_cnt * len : Wt;  // partials behind all the summed Wt
        GemmNT p{};
        p.A = w.A + (size_t)cf * pw; p.lda = w.ld; p.a_tri = 1; p.a_row_off = (int)(cf * pw); p.s_row_off = (int)step;
        p.Adiag = w.DU;
        p.B = blkptr(w, j0, 0); p.ldb = w.ld;
        p.C = part; p.ldc = pw;
        p.M = (int)pw; p.N = (int)wJ; p.K = j0 * TB; p.alpha = 1.0; p.beta = 0.0;
        p.batch = cnt; p.sA = step; p.sB = 0; p.sC = (long)len;
        if (S > 1) {
          p.ksplit = S; p.klen = KL;
        }
        ACE_TRY(launch_gemm_nt(p, w.bg));
        if (S > 1) {
          const size_t total = (size_t)cnt * len;
          splitk_sum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, w.bg>>>(part, S, len, total, Wt);
          ACE_CUDA(cudaGetLastError());
        }
        off += (size_t)cnt;
      }
      off = 0;
      for (int cf = r_lo; cf < r_hi; ++cf) {
        const int cnt = count_of(cf);
        if (cnt <= 0) continue;
        GemmNT r{};
        r.A = blkptr(w, j0, j0); r.lda = w.ld; r.a_tri = 2; r.Adiag = w.DX + (size_t)j0 * TB * TB;
        r.B = w.Bf + off * len; r.ldb = pw;
        r.C = blkptr(w, j0, cf * pb); r.ldc = w.ld;
        r.Ct = blkptr(w, cf * pb, j0); r.ldct = w.ld;
        r.M = (int)wJ; r.N = (int)pw; r.K = (int)wJ; r.alpha = -1.0; r.beta = 0.0;
        r.batch = cnt; r.sA = 0; r.sB = pw * wJ; r.sC = step * w.ld; r.sCt = step;
        ACE_TRY(launch_gemm_nt(r, w.bg));
        off += (size_t)cnt;
      }
      ACE_CUDA(cudaEventRecord(w.ev_aux, w.bg));
    }
    // ================= main stream: panel J applied to the rest of my panels (all below row j2: bulk operands)
    ACE_CUDA(cudaStreamWaitEvent(w.main, ev_bulk[J], 0));
    tr.mark(J, 7, w.main);
    bool first = true;
    {
      // a rank's panels c >= J+2: the soonest needed one alone (its completion releases the look-ahead), all its other
      // full-width ones in ONE batched launch of shrinking height (one tail instead of one per panel), a ragged
      // last panel alone.  Emulation plays the ranks in turn, the owner of panel J+2 first, each with exactly the
      // world-strided launches of a real rank.
      const int G = cx.world;
      auto apply = [&](int c, int count) -> int {
        const int c0 = c * pb, c1 = std::min(c0 + pb, nb);
        const double* pan = Wp + (size_t)(c0 - j2) * TB;
        GemmNT g{};
        g.A = pan; g.lda = mB; g.B = pan; g.ldb = mB; g.C = blkptr(w, c0, c0); g.ldc = w.ld;
        g.M = (nb - c0) * TB; g.N = (c1 - c0) * TB; g.K = (int)wJ; g.alpha = -1.0; g.beta = 1.0;
        if (count > 1) {
          const long step = (long)G * pb * TB;
          g.batch = count; g.sA = step; g.sB = step; g.sC = step * (w.ld + 1); g.m_dec = (int)step;
        }
        return launch_gemm_nt(g, w.main);
      };
      const int nplay = cx.emulate ? G : 1;
      for (int k = 0; k < nplay; ++k) {
        const int r = cx.emulate ? (J + 2 + k) % G : cx.rank;
        int cf = J + 2;
        while (cf < NP && cf % G != r) ++cf;
        if (cf >= NP) continue;
        ACE_TRY(apply(cf, 1));
        if (first) {
          ACE_CUDA(cudaEventRecord(ev_first[J], w.main));
          first = false;
        }
        int cnt = 0, last_ragged = -1;
        for (int c = cf + G; c < NP; c += G) {
          if ((c + 1) * pb > nb) last_ragged = c; else ++cnt;
        }
        if (cnt > 0) ACE_TRY(apply(cf + G, cnt));
        if (last_ragged >= 0) ACE_TRY(apply(last_ragged, 1));
      }
    }
    if (first) ACE_CUDA(cudaEventRecord(ev_first[J], w.main));
    ACE_CUDA(cudaEventRecord(ev_step[J], w.main));
    tr.mark(J, 8, w.main);
  }
  ACE_CUDA(cudaStreamWaitEvent(w.main, ev_diag[NP - 1], 0));
  ACE_CUDA(cudaStreamWaitEvent(w.main, ev_la[NP - 1], 0));
  if (use_mid) ACE_CUDA(cudaStreamWaitEvent(w.main, ev_midla[NP - 1], 0));
  ACE_CUDA(cudaStreamWaitEvent(w.main, ev_copy[NP - 1], 0));
  if (!cx.emulate)  // a failed pivot anywhere is everybody's failure
    ACE_NCCL(nc.AllReduce(w.info, w.info, 1, ncclInt32, ncclMax, cx.comm, w.main));
  return 0;
}

// packed piece (R x C, ld R) -> dst (ld ldd); optionally also its transpose -> dstT (C x R, ld lddt)
__global__ void __launch_bounds__(256) unpack_piece_kernel(const double* __restrict__ src, int R, int C,
                                                           double* __restrict__ dst, long ldd,
                                                           double* __restrict__ dstT, long lddt) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const double v = src[(size_t)c * R + r0 + tx];
    dst[(size_t)c * ldd + r0 + tx] = v;
    tile[ty + 8 * k][tx] = v;
  }
  if (dstT == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k;
    dstT[(size_t)r * lddt + c0 + tx] = tile[tx][ty + 8 * k];
  }
}

inline bool level_is_split(const ShardCtx& cx, int h) {
  return cx.world > 1 && h >= cx.h_min && h % (2 * cx.world) == 0;
}

// one split level (child size h) of the merges inside [lo, lo + len); `ws` needs <= 2 * (len*128/2)^2 doubles
inline int trtri_level_sharded(const DenseWork& w, int lo, int len, int h, cudaStream_t st, double* ws,
                               const ShardCtx& cx, bool need_lower) {
  const int G = cx.world;
  const int hs = h / (2 * G);        // 128-blocks per slice
  const int Rs = hs * TB, s1 = h * TB;
  struct Node { int a, c, s2; size_t off; };  // off: offset of the node's two slices inside a rank chunk
  std::vector<Node> nodes;
  size_t chunk = 0;
  for (int q = 0; q * 2 * h < len; ++q) {
    const int a = lo + q * 2 * h, c = a + h;
    if (c >= lo + len) break;
    const int s2 = std::min(h, lo + len - c) * TB;
    nodes.push_back({a, c, s2, chunk});
    chunk += (size_t)2 * Rs * s2;
  }
  if (nodes.empty()) return 0;
  double* stage = ws;                    // [G][chunk]: the all-gather buffer
  double* wt = ws + (size_t)G * chunk;   // Wt slices of the rank(s) this process computes
  const int r_lo = cx.emulate ? 0 : cx.rank, r_hi = cx.emulate ? G : cx.rank + 1;
  auto slice_of = [&](int r, int e) { return e == 0 ? r : 2 * G - 1 - r; };
  // The slice products of one phase are independent and individually too small to fill the GPU at the lower split
  // levels, so they are dealt over three streams (st + the two helper streams of the dense engine) and joined.
  cudaStream_t ss[3] = {st, w.side, w.aux};
  const int ns = (w.side && w.aux && st == w.main) ? 3 : 1;
  auto fork = [&]() -> int {
    if (ns == 1) return 0;
    ACE_CUDA(cudaEventRecord(w.ev_upd[0], st));
    ACE_CUDA(cudaStreamWaitEvent(ss[1], w.ev_upd[0], 0));
    ACE_CUDA(cudaStreamWaitEvent(ss[2], w.ev_upd[0], 0));
    return 0;
  };
  auto join = [&]() -> int {
    if (ns == 1) return 0;
    ACE_CUDA(cudaEventRecord(w.ev_panel[0], ss[1]));
    ACE_CUDA(cudaEventRecord(w.ev_panel[1], ss[2]));
    ACE_CUDA(cudaStreamWaitEvent(st, w.ev_panel[0], 0));
    ACE_CUDA(cudaStreamWaitEvent(st, w.ev_panel[1], 0));
    return 0;
  };
  // Wt[slice] = U11[slice rows, :] * L21^T    (all of them BEFORE any X21 overwrites L21)
  shard_trace().level_mark(h, 0, st);
  ACE_TRY(fork());
  int li = 0;
  for (int r = r_lo; r < r_hi; ++r)
    for (const Node& nd : nodes)
      for (int e = 0; e < 2; ++e) {
        const int ro = slice_of(r, e) * Rs;
        GemmNT p{};
        p.A = blkptr(w, nd.a, nd.a) + ro; p.lda = w.ld; p.a_tri = 1; p.a_row_off = ro;
        p.Adiag = w.DU + (size_t)nd.a * TB * TB;
        p.B = blkptr(w, nd.c, nd.a); p.ldb = w.ld;
        p.C = wt + (size_t)(r - r_lo) * chunk + nd.off + (size_t)e * Rs * nd.s2; p.ldc = Rs;
        p.M = Rs; p.N = nd.s2; p.K = s1; p.alpha = 1.0; p.beta = 0.0;
        ACE_TRY(launch_gemm_nt(p, ss[li++ % ns]));
      }
  ACE_TRY(join());
  shard_trace().level_mark(h, 1, st);
  // X21[:, slice] = -X22 * Wt[slice]^T, transposed copy (= U12[slice rows, :]) packed into the staging chunk
  ACE_TRY(fork());
  li = 0;
  for (int r = r_lo; r < r_hi; ++r)
    for (const Node& nd : nodes)
      for (int e = 0; e < 2; ++e) {
        const int ro = slice_of(r, e) * Rs;
        GemmNT q{};
        q.A = blkptr(w, nd.c, nd.c); q.lda = w.ld; q.a_tri = 2; q.Adiag = w.DX + (size_t)nd.c * TB * TB;
        q.B = wt + (size_t)(r - r_lo) * chunk + nd.off + (size_t)e * Rs * nd.s2; q.ldb = Rs;
        q.C = blkptr(w, nd.c, nd.a) + (size_t)ro * w.ld; q.ldc = w.ld;
        q.Ct = stage + (size_t)r * chunk + nd.off + (size_t)e * Rs * nd.s2; q.ldct = Rs;
        q.M = nd.s2; q.N = Rs; q.K = nd.s2; q.alpha = -1.0; q.beta = 0.0;
        ACE_TRY(launch_gemm_nt(q, ss[li++ % ns]));
      }
  ACE_TRY(join());
  shard_trace().level_mark(h, 2, st);
  if (!cx.emulate) {
    NcclApi& nc = nccl_api();
    ACE_NCCL(nc.AllGather(stage + (size_t)cx.rank * chunk, stage, chunk, ncclFloat64, cx.comm, st));
  }
  shard_trace().level_mark(h, 3, st);
  for (int r = 0; r < G; ++r)
    for (const Node& nd : nodes)
      for (int e = 0; e < 2; ++e) {
        const int sl = slice_of(r, e);
        const double* src = stage + (size_t)r * chunk + nd.off + (size_t)e * Rs * nd.s2;
        dim3 grid(Rs / 32, nd.s2 / 32);
        unpack_piece_kernel<<<grid, 256, 0, st>>>(src, Rs, nd.s2, blkptr(w, nd.a + sl * hs, nd.c), w.ld,
                                                  need_lower ? blkptr(w, nd.c, nd.a + sl * hs) : nullptr, w.ld);
        ACE_CUDA(cudaGetLastError());
      }
  shard_trace().level_mark(h, 4, st);
  return 0;
}

// Phase 2 after an incremental run: the background stream is joined and every rank's column panels of X travel to
// everybody (packed lower trapezoids, one grouped broadcast per panel), unpacked as X (lower) and U (upper).
inline int gather_inverse_sharded(const DenseWork& w, const ShardCtx& cx) {
  const int nb = w.nb, pb = w.panel_blocks, NP = shard_panels(nb, pb);
  cudaStream_t st = w.main;
  if (NP >= 2) ACE_CUDA(cudaStreamWaitEvent(st, w.ev_aux, 0));
  shard_trace().level_mark(1 << 20, 0, st);
  // staging = [world][chunk]: rank q's panels packed one after the other in chunk q (chunk = the largest rank total,
  // i.e. rank 0's: it owns the earliest, tallest panel of every cycle), exchanged with ONE ncclAllGather
  const int G = cx.world;
  std::vector<size_t> off(NP, 0), tot(G, 0);
  for (int c = 0; c < NP; ++c) {
    const int c0 = c * pb, c1 = std::min(c0 + pb, nb);
    off[c] = tot[c % G];
    tot[c % G] += (size_t)(nb - c1) * TB * (size_t)(c1 - c0) * TB;
  }
  size_t chunk = 0;
  for (int q = 0; q < G; ++q) chunk = std::max(chunk, tot[q]);
  double* stage = w.Bf;
  if ((size_t)G * chunk > (size_t)w.ld * w.ld) {
    set_error("gather_inverse_sharded: staging does not fit");
    return -2;
  }
  for (int c = 0; c < NP; ++c) {
    const int c0 = c * pb, c1 = std::min(c0 + pb, nb);
    const long mC = (long)(nb - c1) * TB, wC = (long)(c1 - c0) * TB;
    if (mC == 0 || !cx.mine(c)) continue;
    ACE_CUDA(cudaMemcpy2DAsync(stage + (size_t)(c % G) * chunk + off[c], sizeof(double) * mC, blkptr(w, c1, c0),
                               sizeof(double) * w.ld, sizeof(double) * mC, (size_t)wC, cudaMemcpyDeviceToDevice, st));
  }
  shard_trace().level_mark(1 << 20, 1, st);
  shard_trace().level_mark(1 << 20, 2, st);
  if (!cx.emulate && chunk > 0) {
    NcclApi& nc = nccl_api();
    ACE_NCCL(nc.AllGather(stage + (size_t)cx.rank * chunk, stage, chunk, ncclFloat64, cx.comm, st));
  }
  shard_trace().level_mark(1 << 20, 3, st);
  for (int c = 0; c < NP; ++c) {
    const int c0 = c * pb, c1 = std::min(c0 + pb, nb);
    const int mC = (nb - c1) * TB, wC = (c1 - c0) * TB;
    if (mC == 0) continue;
    dim3 grid(mC / 32, wC / 32);
    unpack_piece_kernel<<<grid, 256, 0, st>>>(stage + (size_t)(c % G) * chunk + off[c], mC, wC, blkptr(w, c1, c0), w.ld,
                                              blkptr(w, c0, c1), w.ld);
    ACE_CUDA(cudaGetLastError());
  }
  shard_trace().level_mark(1 << 20, 4, st);
  return 0;
}

// Phase 2 for a sharded fit: low levels redundantly, high levels split.
inline int trtri_merge_sharded(const DenseWork& w, const ShardCtx& cx) {
  for (int h = trtri_hmin(w); h < w.nb; h *= 2) {
    if (level_is_split(cx, h))
      ACE_TRY(trtri_level_sharded(w, 0, w.nb, h, w.main, w.Bf, cx, /*need_lower=*/true));  // X complete: posterior in factor form
    else {
      shard_trace().level_mark(-h, 0, w.main);
      ACE_TRY(trtri_level(w, 0, w.nb, h, w.main, w.Bf));
      for (int k = 1; k < 5; ++k) shard_trace().level_mark(-h, k, w.main);  // one phase only: the rest reads +0
    }
  }
  return 0;
}

// Phase 3 for a sharded fit: this rank's tiles of K^-1 (tile L of the lower_only order belongs to rank L mod G)
inline int uut_inverse_sharded(const DenseWork& w, const ShardCtx& cx) {
  const int r_lo = cx.emulate ? 0 : cx.rank, r_hi = cx.emulate ? cx.world : cx.rank + 1;
  for (int r = r_lo; r < r_hi; ++r) {
    GemmNT p{};
    p.A = w.A; p.lda = w.ld; p.a_tri = 1; p.Adiag = w.DU;
    p.B = w.A; p.ldb = w.ld; p.Bdiag = w.DU;
    p.C = w.Bf; p.ldc = w.ld; p.Ct = w.Bf; p.ldct = w.ld;
    p.M = p.N = p.K = w.nb * TB;
    p.alpha = 1.0; p.beta = 0.0; p.lower_only = 1;
    p.tile_first = r; p.tile_stride = cx.world;
    ACE_TRY(launch_gemm_nt(p, w.main));
  }
  return 0;
}

}  // namespace ace
