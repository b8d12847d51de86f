// shard_dense.cuh -- the dense part of ONE fit spread over the G GPUs of a node (first step).
//
// After the (still redundant) Cholesky every rank holds L.  The two n^3/3 phases that follow are sharded:
//
//  * triangular inverse: the bottom-up merge tree of chol.cuh is kept; its LOW levels (small nodes, ~6 % of the
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
// `emulate`: a single process plays all G ranks one after the other on one GPU (no NCCL).  Numerically this is
// the multi-GPU path bit for bit, which lets the single-GPU test tier cover it.
#pragma once
#include "chol.cuh"
#include "nccl_dyn.h"

namespace ace {

struct ShardCtx {
  int rank = 0, world = 1;
  bool emulate = false;
  ncclComm_t comm = nullptr;
  int h_min = 16;  // levels with child size h >= h_min (in 128-blocks) and h % (2*world) == 0 are split
  cudaEvent_t* events = nullptr;  // 4 * panels events of potrf_sharded (panel, first, step, copy)
  int* info_tmp = nullptr;
  bool mine(int panel) const { return emulate || panel % world == rank; }
};

inline int shard_panels(int nb, int pb) { return (nb + pb - 1) / pb; }

// Phase 1 for a sharded fit.  Needs w.Wp[0..1] (packed panels) and w.Wsmall.
inline int potrf_sharded(const DenseWork& w, const ShardCtx& cx) {
  const int nb = w.nb, pb = w.panel_blocks, NP = shard_panels(nb, pb);
  if (!w.Wp[0] || !cx.events) {
    set_error("potrf_sharded: packed panel buffers missing");
    return -2;
  }
  cudaEvent_t* ev_panel = cx.events;
  cudaEvent_t* ev_first = cx.events + NP;
  cudaEvent_t* ev_step = cx.events + 2 * NP;
  cudaEvent_t* ev_copy = cx.events + 3 * NP;
  NcclApi& nc = nccl_api();
  ACE_CUDA(cudaMemsetAsync(w.info, 0, sizeof(int), w.main));
  ACE_CUDA(cudaEventRecord(w.ev_upd[1], w.main));  // fork: side and aux join after everything queued on main
  ACE_CUDA(cudaStreamWaitEvent(w.side, w.ev_upd[1], 0));
  ACE_CUDA(cudaStreamWaitEvent(w.aux, w.ev_upd[1], 0));
  // panel J applied to panel c (both in panel units): A[c0:, c] -= L[c0:, J] * L[c rows, J]^T, operands packed
  auto apply = [&](int J, int c, cudaStream_t st) -> int {
    const int j0 = J * pb, c0 = c * pb, c1 = std::min(c0 + pb, nb);
    const long mJ = (long)(nb - j0) * TB;
    const double* pan = w.Wp[J & 1] + (size_t)(c0 - j0) * TB;
    GemmNT g{};
    g.A = pan; g.lda = mJ; g.B = pan; g.ldb = mJ; g.C = blkptr(w, c0, c0); g.ldc = w.ld;
    g.M = (nb - c0) * TB; g.N = (c1 - c0) * TB; g.K = (std::min(j0 + pb, nb) - j0) * TB;
    g.alpha = -1.0; g.beta = 1.0;
    return launch_gemm_nt(g, st);
  };
  for (int J = 0; J < NP; ++J) {
    const int j0 = J * pb, j1 = std::min(j0 + pb, nb);
    const long mJ = (long)(nb - j0) * TB, wJ = (long)(j1 - j0) * TB;
    double* Wp = w.Wp[J & 1];
    // ---- side stream: panel J becomes available in Wp[J & 1]
    if (J >= 2) {  // the buffer is free once panel J-2 has been applied everywhere and copied out
      ACE_CUDA(cudaStreamWaitEvent(w.side, ev_step[J - 2], 0));
      ACE_CUDA(cudaStreamWaitEvent(w.side, ev_copy[J - 2], 0));
    }
    if (cx.mine(J)) {
      ACE_TRY(potrf_rec(w, j0, j1, w.side));
      ACE_TRY(trtri_merge_range(w, j0, j1, w.side, w.Wsmall));  // X_JJ / U_JJ in place (early low merge levels)
      ACE_CUDA(cudaMemcpy2DAsync(Wp, sizeof(double) * mJ, blkptr(w, j0, j0), sizeof(double) * w.ld,
                                 sizeof(double) * wJ, (size_t)wJ, cudaMemcpyDeviceToDevice, w.side));
      if (j1 < nb) {
        GemmNT t{};
        t.A = blkptr(w, j1, j0); t.lda = w.ld;
        t.B = blkptr(w, j0, j0); t.ldb = w.ld; t.b_tri = 2; t.Bdiag = w.DX + (size_t)j0 * TB * TB;
        t.C = Wp + wJ; t.ldc = mJ;
        t.M = (nb - j1) * TB; t.N = (int)wJ; t.K = (int)wJ; t.alpha = 1.0; t.beta = 0.0;
        ACE_TRY(launch_gemm_nt(t, w.side));
      }
    }
    if (!cx.emulate) {
      const int root = J % cx.world;
      double* dx = w.DX + (size_t)j0 * TB * TB;
      double* du = w.DU + (size_t)j0 * TB * TB;
      double* dv = w.dvec + (size_t)j0 * TB;
      ACE_NCCL(nc.GroupStart());
      ACE_NCCL(nc.Broadcast(Wp, Wp, (size_t)mJ * wJ, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.Broadcast(dx, dx, (size_t)wJ * TB, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.Broadcast(du, du, (size_t)wJ * TB, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.Broadcast(dv, dv, (size_t)wJ, ncclFloat64, root, cx.comm, w.side));
      ACE_NCCL(nc.GroupEnd());
    }
    ACE_CUDA(cudaEventRecord(ev_panel[J], w.side));
    // ---- aux stream: the panel into this rank's A (the owner only lacks the solved rows below the diagonal block)
    ACE_CUDA(cudaStreamWaitEvent(w.aux, ev_panel[J], 0));
    {
      const long skip = cx.mine(J) ? wJ : 0;
      if (mJ > skip)
        ACE_CUDA(cudaMemcpy2DAsync(blkptr(w, j0, j0) + skip, sizeof(double) * w.ld, Wp + skip, sizeof(double) * mJ,
                                   sizeof(double) * (mJ - skip), (size_t)wJ, cudaMemcpyDeviceToDevice, w.aux));
    }
    ACE_CUDA(cudaEventRecord(ev_copy[J], w.aux));
    // ---- look-ahead: the next panel, if it is mine, gets panel J at once and on the high-priority stream
    const bool la = (J + 1 < NP) && cx.mine(J + 1);
    if (la) {
      if (J >= 1) ACE_CUDA(cudaStreamWaitEvent(w.side, ev_first[J - 1], 0));  // panels <= J-1 already applied to it
      ACE_TRY(apply(J, J + 1, w.side));
    }
    // ---- main stream: panel J applied to the rest of my panels, the soonest needed first
    ACE_CUDA(cudaStreamWaitEvent(w.main, ev_panel[J], 0));
    bool first = true;
    for (int c = J + 1; c < NP; ++c) {
      if (!cx.mine(c) || (la && c == J + 1)) continue;
      ACE_TRY(apply(J, c, w.main));
      if (first) {
        ACE_CUDA(cudaEventRecord(ev_first[J], w.main));
        first = false;
      }
    }
    if (first) ACE_CUDA(cudaEventRecord(ev_first[J], w.main));
    ACE_CUDA(cudaEventRecord(ev_step[J], w.main));
  }
  ACE_CUDA(cudaStreamWaitEvent(w.main, ev_panel[NP - 1], 0));
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
  // Wt[slice] = U11[slice rows, :] * L21^T    (all of them BEFORE any X21 overwrites L21)
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
        ACE_TRY(launch_gemm_nt(p, st));
      }
  // X21[:, slice] = -X22 * Wt[slice]^T, transposed copy (= U12[slice rows, :]) packed into the staging chunk
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
        ACE_TRY(launch_gemm_nt(q, st));
      }
  if (!cx.emulate) {
    NcclApi& nc = nccl_api();
    ACE_NCCL(nc.AllGather(stage + (size_t)cx.rank * chunk, stage, chunk, ncclFloat64, cx.comm, st));
  }
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
  return 0;
}

// Phase 2 for a sharded fit: low levels redundantly, high levels split.
inline int trtri_merge_sharded(const DenseWork& w, const ShardCtx& cx) {
  for (int h = trtri_hmin(w); h < w.nb; h *= 2) {
    if (level_is_split(cx, h))
      ACE_TRY(trtri_level_sharded(w, 0, w.nb, h, w.main, w.Bf, cx, /*need_lower=*/2 * h < w.nb));
    else
      ACE_TRY(trtri_level(w, 0, w.nb, h, w.main, w.Bf));
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
