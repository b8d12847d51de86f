// prep_functions.inl -- C ABI of the GPU-side preprocessing (prep_kernels.cuh).
// The control flow is the reference's (src/utilities_cpp.cpp:13-118) with its index quirks, exactly as in the host
// version (host_utils.inl: ace_normalize_train); the column statistics, the y moments and every transform run on the
// device, on device-resident copies of y, X, Z.

namespace ace {

static int prep_col_op(double* c, int n, double a, double b, int mode, cudaStream_t st) {
  col_affine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c, n, a, b, mode);
  ACE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ace

extern "C" {

int ace_normalize_train_gpu(double* y, double* X, double* Z, int n, int px, int pz, double* moments) {
  if (!y || !X || !Z || !moments || n < 2 || px < 1 || pz < 1) return usage("normalize_train_gpu: bad argument");
  ACE_TRY(check_device());
  ACE_CUDA(cudaSetDevice(g_device));
  cudaStream_t st = nullptr;  // legacy default stream: these calls are synchronous like the reference's
  const int nc = px + pz, R = 1 + nc;
  int npow2 = 1;
  while (npow2 < n) npow2 <<= 1;
  DBuf<double> dcols, dy, scratch, dstats;
  ACE_TRY(dcols.alloc((size_t)n * nc));
  ACE_TRY(dy.alloc((size_t)n + 2));
  ACE_TRY(scratch.alloc((size_t)npow2 * nc));
  ACE_TRY(dstats.alloc((size_t)PREP_STATS * nc));
  ACE_CUDA(cudaMemcpyAsync(dcols.p, X, sizeof(double) * (size_t)n * px, cudaMemcpyHostToDevice, st));
  ACE_CUDA(cudaMemcpyAsync(dcols.p + (size_t)n * px, Z, sizeof(double) * (size_t)n * pz, cudaMemcpyHostToDevice, st));
  ACE_CUDA(cudaMemcpyAsync(dy.p, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  col_sort_stats_kernel<<<nc, 1024, 0, st>>>(dcols.p, n, n, npow2, scratch.p, dstats.p);
  ACE_CUDA(cudaGetLastError());
  std::vector<double> hs((size_t)PREP_STATS * nc);
  ACE_CUDA(cudaMemcpyAsync(hs.data(), dstats.p, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, st));
  {  // y on the device while the host walks the columns
    const size_t ybytes = sizeof(double) * (size_t)n;
    const int in_smem = ybytes <= 200 * 1024;
    if (in_smem) ACE_CUDA(cudaFuncSetAttribute(y_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ybytes));
    y_moments_kernel<<<1, 1024, in_smem ? ybytes : 0, st>>>(dy.p, n, in_smem, dy.p + n);
    ACE_CUDA(cudaGetLastError());
  }
  ACE_CUDA(cudaStreamSynchronize(st));
  auto M = [&](int r, int c) -> double& { return moments[r + (size_t)R * c]; };
  for (int r = 0; r < R; ++r) {
    M(r, 0) = 0.0;
    M(r, 1) = 1.0;
    M(r, 2) = 0.0;
  }
  auto colp = [&](int i) -> double* { return dcols.p + (size_t)n * i; };
  // the extremes and middles of every column follow its transforms on the host (each transform is monotone), so
  // that later statistics (median, max |.|) need no second pass over the data
  std::vector<double> lo(nc), hi(nc), m1(nc), m2(nc);
  std::vector<int> nuniq(nc), isbinary(nc, 0);
  for (int i = 0; i < nc; ++i) {
    lo[i] = hs[(size_t)PREP_STATS * i + 0];
    hi[i] = hs[(size_t)PREP_STATS * i + 1];
    m1[i] = hs[(size_t)PREP_STATS * i + 2];
    m2[i] = hs[(size_t)PREP_STATS * i + 3];
    nuniq[i] = (int)hs[(size_t)PREP_STATS * i + 4];
  }
  auto track = [&](int i, double a, double b, int mode) {
    auto f = [&](double v) { return mode == 0 ? (v - a) / b : mode == 1 ? v - a : mode == 2 ? v / b : 0.0; };
    lo[i] = f(lo[i]); hi[i] = f(hi[i]); m1[i] = f(m1[i]); m2[i] = f(m2[i]);
  };
  for (int i = 0; i < nc; ++i) {  // src/utilities_cpp.cpp:27-66
    if (nuniq[i] == 2) {
      isbinary[i] = 1;
      M(i + 1, 2) = 1.0;
      if (lo[i] != 0) M(i, 0) = lo[i];                          // row i, not i+1: as in the reference
      if (hi[i] != 1) M(i, 1) = hi[i] - lo[i];
      const double a = M(i, 0), b = M(i, 1);
      ACE_TRY(prep_col_op(colp(i), n, a, b, 0, st));
      track(i, a, b, 0);
    } else if (nuniq[i] == 1) {
      if (i < px) {
        ACE_TRY(prep_col_op(colp(i), n, 0.0, 1.0, 3, st));
        track(i, 0.0, 1.0, 3);
      } else {
        if (i >= pz) return usage("normalize_train: constant Z column (the reference indexes out of bounds here)");
        ACE_TRY(prep_col_op(colp(px + i), n, 0.0, 1.0, 3, st));  // Z.col(i) with i >= px: as in the reference
        track(px + i, 0.0, 1.0, 3);
        nuniq[px + i] = 1;  // visited later in this loop: the reference recomputes its unique values then
      }
    }
  }
  // y: mean, centre, sd, scale happened on the device (the reference interleaves them with the column loops, but
  // they touch nothing else)
  std::vector<double> ym(2);
  ACE_CUDA(cudaMemcpy(ym.data(), dy.p + n, sizeof(double) * 2, cudaMemcpyDeviceToHost));
  M(0, 0) = ym[0];
  for (int i = 1; i <= nc; ++i) {                                // :71-82
    if (isbinary[i - 1] == 0) {
      const int c = i - 1;
      const double med = (n % 2) ? m2[c] : 0.5 * (m1[c] + m2[c]);
      M(i, 0) = med;
      ACE_TRY(prep_col_op(colp(c), n, med, 1.0, 1, st));
      track(c, med, 1.0, 1);
    }
  }
  M(0, 1) = ym[1];
  for (int i = 1; i <= px; ++i) {                                // :86-94
    if (isbinary[i - 1] == 0) {
      const int c = i - 1;
      const double mx = std::max(0.0, std::max(std::fabs(lo[c]), std::fabs(hi[c])));
      M(i, 1) = mx;
      ACE_TRY(prep_col_op(colp(c), n, 0.0, mx, 2, st));
      track(c, 0.0, mx, 2);
    }
  }
  for (int i = px + 1; i <= nc; ++i) {                           // :96-101  (index i-px-1 into isbinary: literal)
    if (isbinary[i - px - 1] == 0) {
      const int c = px + (i - px - 1);
      const double mx = std::max(0.0, std::max(std::fabs(lo[c]), std::fabs(hi[c])));
      M(i, 1) = mx;
      ACE_TRY(prep_col_op(colp(c), n, 0.0, mx, 2, st));
      track(c, 0.0, mx, 2);
    }
  }
  ACE_CUDA(cudaMemcpyAsync(X, dcols.p, sizeof(double) * (size_t)n * px, cudaMemcpyDeviceToHost, st));
  ACE_CUDA(cudaMemcpyAsync(Z, dcols.p + (size_t)n * px, sizeof(double) * (size_t)n * pz, cudaMemcpyDeviceToHost, st));
  ACE_CUDA(cudaMemcpyAsync(y, dy.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  ACE_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ace_normalize_test_gpu(double* X, double* Z, int n, int px, int pz, const double* moments) {
  if (!X || !Z || !moments || n < 1 || px < 1 || pz < 1) return usage("normalize_test_gpu: bad argument");
  ACE_TRY(check_device());
  ACE_CUDA(cudaSetDevice(g_device));
  cudaStream_t st = nullptr;
  const int nc = px + pz, R = 1 + nc;
  DBuf<double> dcols;
  ACE_TRY(dcols.alloc((size_t)n * nc));
  ACE_CUDA(cudaMemcpyAsync(dcols.p, X, sizeof(double) * (size_t)n * px, cudaMemcpyHostToDevice, st));
  ACE_CUDA(cudaMemcpyAsync(dcols.p + (size_t)n * px, Z, sizeof(double) * (size_t)n * pz, cudaMemcpyHostToDevice, st));
  for (int i = 0; i < nc; ++i)                                   // src/utilities_cpp.cpp:108-118
    ACE_TRY(prep_col_op(dcols.p + (size_t)n * i, n, moments[i + 1], moments[i + 1 + R], 0, st));
  ACE_CUDA(cudaMemcpyAsync(X, dcols.p, sizeof(double) * (size_t)n * px, cudaMemcpyDeviceToHost, st));
  ACE_CUDA(cudaMemcpyAsync(Z, dcols.p + (size_t)n * px, sizeof(double) * (size_t)n * pz, cudaMemcpyDeviceToHost, st));
  ACE_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
