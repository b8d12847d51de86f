// api_functions.inl -- per-function C-ABI entry points (one per reference export), marginal prediction,
// dense debug/bench hooks and the O(n) host preprocessing routines.  Included by ace_b200.cu.

namespace ace {

static int sync_stream(cudaStream_t st) {
  ACE_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// kernel build on host data (rectangular or symmetric)
static int kernmat_host(int kind, int sym, const double* X1, const double* X2, const double* Z1, const double* Z2,
                        int n1, int n2, int p, int Bz, const double* par, double* full, double* elements) {
  // Bz = 0 (no treatment basis: the nuisance term alone, B = 1) is what the reference computes for a Z with no columns
  if (!X1 || !X2 || (Bz > 0 && (!Z1 || !Z2)) || !par || !full) return usage("kernmat: null argument");
  if (n1 < 1 || n2 < 1 || p < 1 || Bz < 0) return usage("kernmat: bad dimensions");
  Core c;
  ACE_TRY(c.init(g_device, n1, p, Bz, kind, false, false));
  const int B = Bz + 1;
  const int n1p = round_up(n1, TB), n2p = sym ? n1p : round_up(n2, TB);
  DBuf<double> dX2, dZ2, dLZ2, K, cube;
  ACE_TRY(c.upload_data(nullptr, X1, Z1));
  ACE_TRY(c.upload_theta(par));
  const double *x2 = c.X.p, *z2 = c.Z.p, *lz2 = c.LZ.p;
  if (!sym) {
    ACE_TRY(dX2.alloc((size_t)n2p * p));
    ACE_TRY(dZ2.alloc((size_t)n2p * std::max(Bz, 1)));
    ACE_TRY(dLZ2.alloc((size_t)n2p * std::max(Bz, 1)));
    ACE_TRY(upload_matrix(dX2.p, n2p, n2p, X2, n2, p, c.st));
    ACE_TRY(upload_matrix(dZ2.p, n2p, n2p, Z2, n2, Bz, c.st));
    if (Bz > 0)
      logabs_kernel<<<(unsigned)(((size_t)n2p * Bz + 255) / 256), 256, 0, c.st>>>(dZ2.p, dLZ2.p, (size_t)n2p * Bz);
    ACE_CUDA(cudaGetLastError());
    x2 = dX2.p; z2 = dZ2.p; lz2 = dLZ2.p;
  }
  ACE_TRY(K.alloc((size_t)n1p * n2p));
  if (elements) ACE_TRY(cube.alloc((size_t)n1p * n2p * B));
  ACE_TRY(c.enqueue_prep());
  KernArgs a{};
  a.X1 = c.X.p; a.Z1 = c.Z.p; a.LZ1 = c.LZ.p; a.ld1 = n1p;
  a.X2 = x2; a.Z2 = z2; a.LZ2 = lz2; a.ld2 = n2p;
  a.n1 = n1; a.n2 = sym ? n1 : n2; a.n1_pad = n1p; a.n2_pad = n2p; a.p = p; a.B = B; a.tab = c.tab.p;
  a.K = K.p; a.ldk = n1p; a.cube = elements ? cube.p : nullptr; a.cube_slice = (long)n1p * n2p;
  a.sym = sym;
  ACE_TRY(launch_kernmat(a, kind, c.st));
  ACE_TRY(download_matrix(full, n1, a.n2, K.p, n1p, c.st));
  if (elements)
    for (int b = 0; b < B; ++b)
      ACE_TRY(download_matrix(elements + (size_t)b * n1 * a.n2, n1, a.n2, cube.p + (size_t)b * n1p * n2p, n1p, c.st));
  return sync_stream(c.st);
}

// shared by grad_*_cpp: gradients + stats from host inputs
static int grad_host(int kind, const double* y, const double* X, const double* Z, const double* invK,
                     const double* eigenval, const double* par, double* stats, unsigned B, double std_y, int n, int p,
                     double* gradients) {
  if (!y || !X || (B > 1 && !Z) || !invK || !eigenval || !par || !gradients) return usage("grad: null argument");
  if (n < 1 || p < 1 || B < 1) return usage("grad: bad dimensions");
  Core c;
  ACE_TRY(c.init(g_device, n, p, (int)B - 1, kind, false, true));
  DBuf<double> Kinv;
  ACE_TRY(Kinv.alloc((size_t)c.n_pad * c.n_pad));
  ACE_TRY(c.upload_data(y, X, Z));
  ACE_TRY(c.upload_theta(par));
  ACE_TRY(upload_matrix(Kinv.p, c.n_pad, c.n_pad, invK, n, n, c.st));
  std::vector<double> d(c.n_pad, 1.0);
  for (int i = 0; i < n; ++i) d[i] = std::sqrt(eigenval[i]);  // 2 sum log d = sum log eigenval
  ACE_CUDA(cudaMemcpyAsync(c.dvec.p, d.data(), sizeof(double) * c.n_pad, cudaMemcpyHostToDevice, c.st));
  ace_fit_config cfg;
  ace_fit_default_config(&cfg);
  cfg.kernel = kind;
  cfg.std_y = std_y;
  ACE_TRY(c.enqueue_prep());
  ACE_TRY(c.enqueue_alpha(Kinv.p, 0));
  ACE_TRY(c.enqueue_grad(Kinv.p));
  ACE_TRY(c.enqueue_finalize(Kinv.p, cfg, 0));
  ACE_TRY(c.fetch_scalars());
  ACE_CUDA(cudaMemcpy(gradients, c.grad.p, sizeof(double) * c.P, cudaMemcpyDeviceToHost));
  if (stats) {
    stats[0] = c.h_sc[SC_RMSE];
    stats[1] = c.h_sc[SC_EVID];
  }
  return 0;
}

}  // namespace ace

extern "C" {

int ace_kernmat_SE_cpp(const double* X1, const double* X2, const double* Z1, const double* Z2, int n1, int n2, int p,
                       int Bz, const double* par, double* full, double* elements) {
  return kernmat_host(0, 0, X1, X2, Z1, Z2, n1, n2, p, Bz, par, full, elements);
}
int ace_kernmat_Matern32_cpp(const double* X1, const double* X2, const double* Z1, const double* Z2, int n1, int n2,
                             int p, int Bz, const double* par, double* full, double* elements) {
  return kernmat_host(1, 0, X1, X2, Z1, Z2, n1, n2, p, Bz, par, full, elements);
}
int ace_kernmat_SE_symmetric_cpp(const double* X, const double* Z, int n, int p, int Bz, const double* par,
                                 double* full, double* elements) {
  return kernmat_host(0, 1, X, X, Z, Z, n, n, p, Bz, par, full, elements);
}
int ace_kernmat_Matern32_symmetric_cpp(const double* X, const double* Z, int n, int p, int Bz, const double* par,
                                       double* full, double* elements) {
  return kernmat_host(1, 1, X, X, Z, Z, n, n, p, Bz, par, full, elements);
}

int ace_invkernel_cpp(const double* pdmat, int n, double sigma, double* eigenval, double* inv) {
  if (!pdmat || n < 1) return usage("invkernel: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, true, false));
  ACE_TRY(upload_matrix(c.A.p, c.n_pad, c.n_pad, pdmat, n, n, c.st));
  add_diag_kernel<<<(n + 255) / 256, 256, 0, c.st>>>(c.A.p, c.n_pad, n, std::exp(sigma));
  if (c.n_pad > n) pad_identity_kernel<<<(c.n_pad - n + 255) / 256, 256, 0, c.st>>>(c.A.p, c.n_pad, n, c.n_pad);
  ACE_CUDA(cudaGetLastError());
  DenseWork w = c.dense(c.A.p, c.Bf.p);
  ACE_TRY(spd_inverse(w));
  if (inv) ACE_TRY(download_matrix(inv, n, n, c.Bf.p, c.n_pad, c.st));
  std::vector<double> d(n);
  ACE_CUDA(cudaMemcpyAsync(d.data(), c.dvec.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.st));
  ACE_TRY(c.fetch_scalars());
  if (eigenval)
    for (int i = 0; i < n; ++i) eigenval[i] = d[i] * d[i];
  if (c.h_info() > 0) {
    set_error("matrix not positive definite at pivot " + std::to_string(c.h_info()));
    return c.h_info();
  }
  return 0;
}

int ace_grad_SE_cpp(const double* y, const double* X, const double* Z, const double* Kfull, const double* K,
                    const double* invKmatn, const double* eigenval, const double* parameters, double* stats,
                    unsigned int B, double std_y, int n, int p, double* gradients) {
  (void)Kfull;
  (void)K;
  return grad_host(0, y, X, Z, invKmatn, eigenval, parameters, stats, B, std_y, n, p, gradients);
}
int ace_grad_Matern_cpp(const double* y, const double* X, const double* Z, const double* Kfull, const double* K,
                        const double* invKmatn, const double* eigenval, const double* parameters, double* stats,
                        unsigned int B, double std_y, int n, int p, double* gradients) {
  (void)Kfull;
  (void)K;
  return grad_host(1, y, X, Z, invKmatn, eigenval, parameters, stats, B, std_y, n, p, gradients);
}

int ace_stats_cpp(const double* y, const double* Kmat, const double* invKmatn, const double* eigenval, double mu,
                  double std_y, int n, double* out) {
  if (!y || !Kmat || !invKmatn || !eigenval || !out || n < 1) return usage("stats: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, false, true));
  DBuf<double> M;
  ACE_TRY(M.alloc((size_t)c.n_pad * c.n_pad));
  ACE_TRY(c.upload_data(y, nullptr, nullptr));
  std::vector<double> th(c.P, 0.0);
  th[1] = mu;
  ACE_TRY(c.upload_theta(th.data()));
  ACE_TRY(upload_matrix(M.p, c.n_pad, c.n_pad, invKmatn, n, n, c.st));
  std::vector<double> d(c.n_pad, 1.0);
  for (int i = 0; i < n; ++i) d[i] = std::sqrt(eigenval[i]);
  ACE_CUDA(cudaMemcpyAsync(c.dvec.p, d.data(), sizeof(double) * c.n_pad, cudaMemcpyHostToDevice, c.st));
  ACE_TRY(c.enqueue_alpha(M.p, 0));  // alpha = invK (y - mu)
  ACE_TRY(upload_matrix(M.p, c.n_pad, c.n_pad, Kmat, n, n, c.st));
  dim3 grid((c.n_pad + gv::ROWS - 1) / gv::ROWS, c.nchunks);
  gemv2_kernel<<<grid, gv::ROWS, 0, c.st>>>(M.p, c.n_pad, n, c.alpha.p, c.pu.p, c.ps.p, c.n_pad);
  reduce_partials_kernel<<<(c.n_pad + 255) / 256, 256, 0, c.st>>>(c.pu.p, c.nchunks, n, c.n_pad, c.Ka.p);
  stats_kernel<<<1, 1024, 0, c.st>>>(c.y.p, c.alpha.p, c.Ka.p, c.dvec.p, n, mu, std_y, c.sc.p);
  ACE_CUDA(cudaGetLastError());
  ACE_TRY(c.fetch_scalars());
  out[0] = c.h_sc[SC_RMSE];
  out[1] = c.h_sc[SC_EVID];
  return 0;
}

int ace_mu_solution_cpp(const double* y, const double* invKmat, int n, double* mu) {
  if (!y || !invKmat || !mu || n < 1) return usage("mu_solution: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, false, true));
  DBuf<double> M;
  ACE_TRY(M.alloc((size_t)c.n_pad * c.n_pad));
  ACE_TRY(c.upload_data(y, nullptr, nullptr));
  std::vector<double> th(c.P, 0.0);
  ACE_TRY(c.upload_theta(th.data()));
  ACE_TRY(upload_matrix(M.p, c.n_pad, c.n_pad, invKmat, n, n, c.st));
  ACE_TRY(c.enqueue_alpha(M.p, 0));
  ACE_TRY(c.fetch_scalars());
  *mu = 0.5 * c.h_sc[SC_SUM_U] / c.h_sc[SC_SUM_S];
  return 0;
}

int ace_pred_cpp(const double* y_X, double sigma, double mu, const double* invK_XX, const double* K_xX,
                 const double* K_xx, double mean_y, double std_y, int nx, int nX, double* map, double* ci,
                 double* var) {
  if (!y_X || !invK_XX || !K_xX || !K_xx || !map || !ci || !var || nx < 1 || nX < 1) return usage("pred: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, nX, 1, 1, 0, false, false));
  const int nx_pad = round_up(nx, TB);
  DBuf<double> Kinv, Kx, Kxx, kd;
  ACE_TRY(Kinv.alloc((size_t)c.n_pad * c.n_pad));
  ACE_TRY(Kx.alloc((size_t)nx_pad * c.n_pad));
  ACE_TRY(Kxx.alloc((size_t)nx_pad * nx_pad));
  ACE_TRY(kd.alloc(nx_pad));
  ACE_TRY(c.upload_data(y_X, nullptr, nullptr));
  ACE_TRY(upload_matrix(Kinv.p, c.n_pad, c.n_pad, invK_XX, nX, nX, c.st));
  ACE_TRY(upload_matrix(Kx.p, nx_pad, nx_pad, K_xX, nx, c.n_pad > 0 ? nX : 0, c.st));
  ACE_CUDA(cudaMemsetAsync(Kx.p + (size_t)nx_pad * nX, 0, sizeof(double) * (size_t)nx_pad * (c.n_pad - nX), c.st));
  ACE_TRY(upload_matrix(Kxx.p, nx_pad, nx_pad, K_xx, nx, nx, c.st));
  diag_extract_kernel<<<(nx + 255) / 256, 256, 0, c.st>>>(Kxx.p, nx_pad, nx, kd.p);
  ACE_CUDA(cudaGetLastError());
  PostOut o;
  ACE_TRY(posterior_rows(c, Kinv.p, Kx.p, kd.p, nx, nx_pad, mu, std::exp(sigma), o));
  std::vector<double> hm(nx), hv(nx);
  ACE_CUDA(cudaMemcpyAsync(hm.data(), o.map.p, sizeof(double) * nx, cudaMemcpyDeviceToHost, c.st));
  ACE_CUDA(cudaMemcpyAsync(hv.data(), o.var.p, sizeof(double) * nx, cudaMemcpyDeviceToHost, c.st));
  ACE_TRY(sync_stream(c.st));
  for (int i = 0; i < nx; ++i) {  // src/pred_cpp.cpp:20,26-29
    map[i] = mean_y + std_y * (hm[i] + mu);
    const double sd = std_y * std::sqrt(std::fabs(hv[i]));
    ci[i] = map[i] - 1.96 * sd;
    ci[i + nx] = map[i] + 1.96 * sd;
    var[i] = sd * sd;
  }
  return 0;
}

}  // extern "C"

namespace ace {

// common tail of the marginal prediction: Kx (nx_pad x n_pad), Cm (nx_pad x nx_pad, the marginal K_xx) on device
// Kinv == nullptr: factor form (posterior_rows_tri), T K_xX^T = W W^T
// subsets / S / avgS / counts: optional batch of S row subsets (nx x S flags, HOST memory) whose averages are
// evaluated on the sub-blocks of the one posterior covariance (ace_fit_predict_marginal_batch)
static int marginal_tail(Core& c, const double* Kinv, const double* Kx, double* Cm, const double* zx_dev, int nx,
                         int nx_pad, double mu, double std_y, double std_Z, int calculate_ate, const double* Zx_host,
                         double* map, double* ci, double* var, double* avg, const unsigned char* subsets = nullptr,
                         int S = 0, double* avgS = nullptr, int* counts = nullptr) {
  DBuf<double> kd, q, qS;
  DBuf<unsigned char> dsub;
  ACE_TRY(kd.alloc(nx_pad));
  ACE_TRY(q.alloc(4));
  std::vector<double> hqS((size_t)3 * std::max(S, 1), 0.0);
  if (S > 0) {
    calculate_ate = 1;
    ACE_TRY(qS.alloc((size_t)3 * S));
    ACE_TRY(dsub.alloc((size_t)nx * S));
    ACE_CUDA(cudaMemcpyAsync(dsub.p, subsets, (size_t)nx * S, cudaMemcpyHostToDevice, c.st));
  }
  diag_extract_kernel<<<(nx + 255) / 256, 256, 0, c.st>>>(Cm, nx_pad, nx, kd.p);
  ACE_CUDA(cudaGetLastError());
  PostOut o;
  if (Kinv != nullptr)
    ACE_TRY(posterior_rows(c, Kinv, Kx, kd.p, nx, nx_pad, mu, 0.0, o));
  else
    ACE_TRY(posterior_rows_tri(c, Kx, kd.p, nx, nx_pad, mu, 0.0, o));
  double hq[3] = {0, 0, 0};
  if (calculate_ate) {
    // C = K_m,xx - T K_m,xX^T  (src/pred_cpp.cpp:72), full block because the averages need it
    GemmNT g{};
    g.A = o.T.p; g.lda = nx_pad; g.B = (Kinv != nullptr) ? Kx : o.T.p; g.ldb = nx_pad; g.C = Cm; g.ldc = nx_pad;
    g.M = nx_pad; g.N = nx_pad; g.K = c.n_pad; g.alpha = -1.0; g.beta = 1.0;
    ACE_TRY(launch_gemm_nt(g, c.st));
    quadforms_kernel<<<1, 1024, 0, c.st>>>(Cm, nx_pad, nx, zx_dev, q.p);
    ACE_CUDA(cudaGetLastError());
    ACE_CUDA(cudaMemcpyAsync(hq, q.p, sizeof(double) * 3, cudaMemcpyDeviceToHost, c.st));
    if (S > 0) {
      quadforms_subset_kernel<<<S, 1024, 0, c.st>>>(Cm, nx_pad, nx, zx_dev, dsub.p, qS.p);
      ACE_CUDA(cudaGetLastError());
      ACE_CUDA(cudaMemcpyAsync(hqS.data(), qS.p, sizeof(double) * 3 * S, cudaMemcpyDeviceToHost, c.st));
    }
  }
  std::vector<double> hm(nx), hv(nx);
  ACE_CUDA(cudaMemcpyAsync(hm.data(), o.map.p, sizeof(double) * nx, cudaMemcpyDeviceToHost, c.st));
  ACE_CUDA(cudaMemcpyAsync(hv.data(), o.var.p, sizeof(double) * nx, cudaMemcpyDeviceToHost, c.st));
  ACE_TRY(sync_stream(c.st));
  for (int i = 0; i < nx; ++i) {  // src/pred_cpp.cpp:70,75-78
    map[i] = std_y * hm[i] / std_Z;
    const double sd = std_y * std::sqrt(std::fabs(hv[i])) / std_Z;
    ci[i] = map[i] - 1.96 * sd;
    ci[i + nx] = map[i] + 1.96 * sd;
    var[i] = sd * sd;
  }
  // averages over a row set (all rows: flags == nullptr), src/pred_cpp.cpp:86-110
  auto averages = [&](const unsigned char* flags, const double* quad, double* out, int* cnt) {
    double s = 0.0, sz = 0.0, dz = 0.0;
    int ns = 0;
    for (int i = 0; i < nx; ++i) {
      if (flags && !flags[i]) continue;
      ++ns;
      s += map[i];
      sz += Zx_host[i];
      dz += map[i] * Zx_host[i];
    }
    const double ate = s / (double)ns;
    const double ate_sd = std_y * std::sqrt(quad[0]) / (double)ns;
    const unsigned int ntx = (unsigned int)sz;
    const double att = dz / ntx;
    const double att_sd = std_y * std::sqrt(quad[1]) / ntx;
    const unsigned int nux = (unsigned int)ns - ntx;
    const double atu = (ate * (double)ns - att * ntx) / nux;
    const double atu_sd = std_y * std::sqrt(quad[2]) / nux;
    const double vals[3] = {ate, att, atu}, sds[3] = {ate_sd, att_sd, atu_sd};
    for (int k = 0; k < 3; ++k) {
      out[4 * k] = vals[k];
      out[4 * k + 1] = vals[k] - 1.96 * sds[k];
      out[4 * k + 2] = vals[k] + 1.96 * sds[k];
      out[4 * k + 3] = sds[k] * sds[k];
    }
    if (cnt) {
      cnt[0] = ns;
      cnt[1] = (int)ntx;
      cnt[2] = (int)nux;
    }
  };
  if (calculate_ate && avg) averages(nullptr, hq, avg, nullptr);
  for (int sidx = 0; sidx < S; ++sidx)
    averages(subsets + (size_t)sidx * nx, hqS.data() + 3 * sidx, avgS + 12 * sidx, counts ? counts + 3 * sidx : nullptr);
  return 0;
}

}  // namespace ace

extern "C" {

int ace_pred_marginal_cpp(const double* y_X, const double* Z_x, double sigma, double mu, const double* invK_XX,
                          const double* K_xX, const double* K_xx, double mean_y, double std_y, double std_Z,
                          int calculate_ate, int nx, int nX, int B, double* map, double* ci, double* var,
                          double* avg) {
  (void)sigma;
  (void)mean_y;
  if (!y_X || !Z_x || !invK_XX || !K_xX || !K_xx || !map || !ci || !var || nx < 1 || nX < 1 || B < 1)
    return usage("pred_marginal: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, nX, 1, 1, 0, false, false));
  const int nx_pad = round_up(nx, TB);
  DBuf<double> Kinv, Kx, Cm, cubeA, cubeB, zx;
  ACE_TRY(Kinv.alloc((size_t)c.n_pad * c.n_pad));
  ACE_TRY(Kx.alloc((size_t)nx_pad * c.n_pad));
  ACE_TRY(Cm.alloc((size_t)nx_pad * nx_pad));
  ACE_TRY(cubeA.alloc((size_t)nx * nX * B));
  ACE_TRY(cubeB.alloc((size_t)nx * nx * B));
  ACE_TRY(zx.alloc(nx_pad));
  ACE_TRY(c.upload_data(y_X, nullptr, nullptr));
  ACE_TRY(upload_matrix(Kinv.p, c.n_pad, c.n_pad, invK_XX, nX, nX, c.st));
  ACE_TRY(upload_matrix(zx.p, nx_pad, nx_pad, Z_x, nx, 1, c.st));
  ACE_CUDA(cudaMemcpyAsync(cubeA.p, K_xX, sizeof(double) * (size_t)nx * nX * B, cudaMemcpyHostToDevice, c.st));
  ACE_CUDA(cudaMemcpyAsync(cubeB.p, K_xx, sizeof(double) * (size_t)nx * nx * B, cudaMemcpyHostToDevice, c.st));
  // slices 1..B-1 (slice 0 when B == 1), src/pred_cpp.cpp:55-67
  DBuf<double> sA, sB;
  ACE_TRY(sA.alloc((size_t)nx * nX));
  ACE_TRY(sB.alloc((size_t)nx * nx));
  const int first = (B > 1) ? 1 : 0;
  cube_sum_kernel<<<(unsigned)(((size_t)nx * nX + 255) / 256), 256, 0, c.st>>>(cubeA.p, (size_t)nx * nX, first, B,
                                                                              (size_t)nx * nX, sA.p);
  cube_sum_kernel<<<(unsigned)(((size_t)nx * nx + 255) / 256), 256, 0, c.st>>>(cubeB.p, (size_t)nx * nx, first, B,
                                                                              (size_t)nx * nx, sB.p);
  ACE_CUDA(cudaGetLastError());
  ACE_CUDA(cudaMemsetAsync(Kx.p, 0, sizeof(double) * (size_t)nx_pad * c.n_pad, c.st));
  ACE_CUDA(cudaMemsetAsync(Cm.p, 0, sizeof(double) * (size_t)nx_pad * nx_pad, c.st));
  ACE_CUDA(cudaMemcpy2DAsync(Kx.p, sizeof(double) * nx_pad, sA.p, sizeof(double) * nx, sizeof(double) * nx, nX,
                             cudaMemcpyDeviceToDevice, c.st));
  ACE_CUDA(cudaMemcpy2DAsync(Cm.p, sizeof(double) * nx_pad, sB.p, sizeof(double) * nx, sizeof(double) * nx, nx,
                             cudaMemcpyDeviceToDevice, c.st));
  return marginal_tail(c, Kinv.p, Kx.p, Cm.p, zx.p, nx, nx_pad, mu, std_y, std_Z, calculate_ate, Z_x, map, ci, var,
                       avg);
}

static int fit_predict_marginal(ace_fit* f, const double* X2, const double* Z2, const double* dZ2, int nx,
                                double std_y, double std_Z, int calculate_ate, double* map, double* ci, double* var,
                                double* avg, const unsigned char* subsets, int S, double* avgS, int* counts) {
  if (!f || !X2 || !Z2 || !dZ2 || !map || !ci || !var || nx < 1) return usage("predict_marginal: bad argument");
  if (f->c.Bz < 1) return usage("predict_marginal: the fit has no treatment basis (Bz = 0)");
  Core& c = f->c;
  ACE_CUDA(cudaSetDevice(c.device));
  const int nx_pad = round_up(nx, TB);
  DBuf<double> dX2, dD, dLD, Kx, Cm, zx;
  ACE_TRY(dX2.alloc((size_t)nx_pad * c.p));
  ACE_TRY(dD.alloc((size_t)nx_pad * c.Bz));
  ACE_TRY(dLD.alloc((size_t)nx_pad * c.Bz));
  ACE_TRY(Kx.alloc((size_t)nx_pad * c.n_pad));
  ACE_TRY(Cm.alloc((size_t)nx_pad * nx_pad));
  ACE_TRY(zx.alloc(nx_pad));
  ACE_TRY(upload_matrix(dX2.p, nx_pad, nx_pad, X2, nx, c.p, c.st));
  ACE_TRY(upload_matrix(dD.p, nx_pad, nx_pad, dZ2, nx, c.Bz, c.st));
  ACE_TRY(upload_matrix(zx.p, nx_pad, nx_pad, Z2, nx, 1, c.st));  // Z_x = first basis column
  logabs_kernel<<<(unsigned)(((size_t)nx_pad * c.Bz + 255) / 256), 256, 0, c.st>>>(dD.p, dLD.p, (size_t)nx_pad * c.Bz);
  ACE_CUDA(cudaGetLastError());
  ACE_TRY(c.enqueue_prep());
  KernArgs a{};  // K_m,xX: rows = new points with the basis DERIVATIVE, columns = training points
  a.X1 = dX2.p; a.Z1 = dD.p; a.LZ1 = dLD.p; a.ld1 = nx_pad;
  a.X2 = c.X.p; a.Z2 = c.Z.p; a.LZ2 = c.LZ.p; a.ld2 = c.n_pad;
  a.n1 = nx; a.n2 = c.n; a.n1_pad = nx_pad; a.n2_pad = c.n_pad; a.p = c.p; a.B = c.B; a.tab = c.tab.p;
  a.K = Kx.p; a.ldk = nx_pad; a.skip0 = 1;
  ACE_TRY(launch_kernmat(a, c.kind, c.st));
  KernArgs s{};  // K_m,xx: symmetric over the new points
  s.X1 = s.X2 = dX2.p; s.Z1 = s.Z2 = dD.p; s.LZ1 = s.LZ2 = dLD.p; s.ld1 = s.ld2 = nx_pad;
  s.n1 = s.n2 = nx; s.n1_pad = s.n2_pad = nx_pad; s.p = c.p; s.B = c.B; s.tab = c.tab.p;
  s.K = Cm.p; s.ldk = nx_pad; s.sym = 1; s.skip0 = 1;
  ACE_TRY(launch_kernmat(s, c.kind, c.st));
  double hpar[2];
  ACE_CUDA(cudaMemcpyAsync(hpar, c.theta.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, c.st));
  ACE_TRY(sync_stream(c.st));
  if (!c.u_valid) ACE_TRY(c.ensure_full_inverse());
  return marginal_tail(c, c.u_valid ? nullptr : c.Bf.p, Kx.p, Cm.p, zx.p, nx, nx_pad, hpar[1], std_y, std_Z, calculate_ate, Z2, map, ci,
                       var, avg, subsets, S, avgS, counts);
}

int ace_fit_predict_marginal(ace_fit* f, const double* X2, const double* Z2, const double* dZ2, int nx, double mean_y,
                             double std_y, double std_Z, int calculate_ate, double* map, double* ci, double* var,
                             double* avg) {
  (void)mean_y;
  return fit_predict_marginal(f, X2, Z2, dZ2, nx, std_y, std_Z, calculate_ate, map, ci, var, avg, nullptr, 0, nullptr,
                              nullptr);
}

int ace_fit_predict_marginal_batch(ace_fit* f, const double* X2, const double* Z2, const double* dZ2, int nx,
                                   double mean_y, double std_y, double std_Z, const unsigned char* subsets, int S,
                                   double* map, double* ci, double* var, double* avg, int* counts) {
  (void)mean_y;
  if (!subsets || S < 1 || !avg) return usage("predict_marginal_batch: bad argument");
  double all[12];
  return fit_predict_marginal(f, X2, Z2, dZ2, nx, std_y, std_Z, 1, map, ci, var, all, subsets, S, avg, counts);
}

// ---------------------------------------------------------------------------------------------
// O(P) host arithmetic: clip + optimisers, in place like the reference
// ---------------------------------------------------------------------------------------------
void ace_norm_clip_cpp(int flag, double* grads, int P, double max_length) {  // src/utilities_cpp.cpp:121-129
  if (!flag || !grads) return;
  double ss = 0.0;
  for (int i = 0; i < P; ++i) ss += grads[i] * grads[i];
  const double L2 = std::sqrt(ss);
  if ((L2 > max_length) && std::isfinite(L2) && (L2 != 0.0))
    for (int i = 0; i < P; ++i) grads[i] = grads[i] / L2;
}

static int all_finite(const double* g, int P) {
  for (int i = 0; i < P; ++i)
    if (!std::isfinite(g[i])) return 0;
  return 1;
}

int ace_Nesterov_cpp(double lr, double momentum, double* nu, const double* grad, double* para, int P) {
  const int ok = all_finite(grad, P);  // src/optimizer_cpp.cpp:8-20
  for (int i = 0; i < P; ++i) {
    nu[i] = momentum * nu[i] + lr * grad[i];
    para[i] = para[i] + nu[i];
  }
  return ok;
}

int ace_Nadam_cpp(double iter, double lr, double beta1, double beta2, double eps, double* m, double* v,
                  const double* grad, double* para, int P) {
  const int ok = all_finite(grad, P);  // src/optimizer_cpp.cpp:23-42
  const double c1 = 1 - std::pow(beta1, iter), c2 = 1 - std::pow(beta2, iter);
  for (int i = 0; i < P; ++i) {
    m[i] = beta1 * m[i] + (1 - beta1) * grad[i];
    v[i] = beta2 * v[i] + (1 - beta2) * (grad[i] * grad[i]);
    para[i] = para[i] + lr * ((beta1 * m[i] + (1 - beta1) * grad[i]) / c1) / (std::sqrt(v[i] / c2) + eps);
  }
  return ok;
}

int ace_Adam_cpp(double iter, double lr, double beta1, double beta2, double eps, double* m, double* v,
                 const double* grad, double* para, int P) {
  const int ok = all_finite(grad, P);  // src/optimizer_cpp.cpp:45-63
  const double c1 = 1 - std::pow(beta1, iter), c2 = 1 - std::pow(beta2, iter);
  for (int i = 0; i < P; ++i) {
    m[i] = (beta1 * m[i]) + (1 - beta1) * grad[i];
    v[i] = beta2 * v[i] + (1 - beta2) * (grad[i] * grad[i]);
    para[i] = para[i] + lr * (m[i] / c1) / (std::sqrt(v[i] / c2) + eps);
  }
  return ok;
}

// ---------------------------------------------------------------------------------------------
// dense debug / bench hooks
// ---------------------------------------------------------------------------------------------
int ace_dbg_gemm_nt(const double* A, const double* B, double* C, int M, int N, int K, double alpha, double beta,
                    int lower_only) {
  if (!A || !B || !C) return usage("gemm: null argument");
  ACE_TRY(check_device());
  ACE_CUDA(cudaSetDevice(g_device));
  ACE_TRY(configure_kernels_once());
  DBuf<double> dA, dB, dC;
  ACE_TRY(dA.alloc((size_t)M * K));
  ACE_TRY(dB.alloc((size_t)N * K));
  ACE_TRY(dC.alloc((size_t)M * N));
  ACE_CUDA(cudaMemcpy(dA.p, A, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice));
  ACE_CUDA(cudaMemcpy(dB.p, B, sizeof(double) * (size_t)N * K, cudaMemcpyHostToDevice));
  ACE_CUDA(cudaMemcpy(dC.p, C, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice));
  GemmNT g{};
  g.A = dA.p; g.lda = M; g.B = dB.p; g.ldb = N; g.C = dC.p; g.ldc = M;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.lower_only = lower_only;
  ACE_TRY(launch_gemm_nt(g, nullptr));
  ACE_CUDA(cudaDeviceSynchronize());
  ACE_CUDA(cudaMemcpy(C, dC.p, sizeof(double) * (size_t)M * N, cudaMemcpyDeviceToHost));
  return 0;
}

int ace_dbg_spd_inverse(const double* A, int n, double* L, double* inv, double* diagL, double* ms3) {
  if (!A || n < 1) return usage("spd_inverse: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, true, false));
  ACE_TRY(upload_matrix(c.A.p, c.n_pad, c.n_pad, A, n, n, c.st));
  if (c.n_pad > n) pad_identity_kernel<<<(c.n_pad - n + 255) / 256, 256, 0, c.st>>>(c.A.p, c.n_pad, n, c.n_pad);
  ACE_CUDA(cudaGetLastError());
  DenseWork w = c.dense(c.A.p, c.Bf.p, /*allow_fused=*/false);  // L itself is an output here
  ACE_CUDA(cudaEventRecord(c.tev[0], c.st));
  ACE_TRY(potrf_blocked(w));
  ACE_CUDA(cudaEventRecord(c.tev[1], c.st));
  if (L) {
    ACE_TRY(download_matrix(L, n, n, c.A.p, c.n_pad, c.st));
    ACE_TRY(sync_stream(c.st));
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < j; ++i) L[i + (size_t)j * n] = 0.0;
  }
  ACE_CUDA(cudaEventRecord(c.tev[2], c.st));
  ACE_TRY(trtri_merge(w));
  ACE_CUDA(cudaEventRecord(c.tev[3], c.st));
  ACE_TRY(uut_inverse(w));
  ACE_CUDA(cudaEventRecord(c.tev[4], c.st));
  if (inv) ACE_TRY(download_matrix(inv, n, n, c.Bf.p, c.n_pad, c.st));
  if (diagL) ACE_CUDA(cudaMemcpyAsync(diagL, c.dvec.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.st));
  ACE_TRY(c.fetch_scalars());
  if (ms3) {
    float t;
    ACE_CUDA(cudaEventElapsedTime(&t, c.tev[0], c.tev[1])); ms3[0] = t;
    ACE_CUDA(cudaEventElapsedTime(&t, c.tev[2], c.tev[3])); ms3[1] = t;
    ACE_CUDA(cudaEventElapsedTime(&t, c.tev[3], c.tev[4])); ms3[2] = t;
  }
  if (c.h_info() > 0) return c.h_info();
  return 0;
}

// production schedule (fused diagonal-block kernel, fused panel TRSM, incremental inverse): inverse + diag(L)
int ace_dbg_spd_inverse_fused(const double* A, int n, double* inv, double* diagL) {
  if (!A || n < 1) return usage("spd_inverse_fused: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, true, false));
  ACE_TRY(upload_matrix(c.A.p, c.n_pad, c.n_pad, A, n, n, c.st));
  if (c.n_pad > n) pad_identity_kernel<<<(c.n_pad - n + 255) / 256, 256, 0, c.st>>>(c.A.p, c.n_pad, n, c.n_pad);
  ACE_CUDA(cudaGetLastError());
  DenseWork w = c.dense(c.A.p, c.Bf.p);
  ACE_TRY(spd_inverse(w));
  if (inv) ACE_TRY(download_matrix(inv, n, n, c.Bf.p, c.n_pad, c.st));
  if (diagL) ACE_CUDA(cudaMemcpyAsync(diagL, c.dvec.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.st));
  ACE_TRY(c.fetch_scalars());
  if (c.h_info() > 0) return c.h_info();
  return 0;
}

// the fused diagonal-block kernel alone on an n x n SPD matrix, n <= 128 * panel width: X = L^-1 assembled from its
// output layout (DX tiles on the diagonal, lower 128-blocks of A elsewhere), U likewise, diag(L), and the diagonal
// 128-tiles of L (the only part of L the layout keeps)
// SM clock stamps of the kernel's phases (8 per tile column: T0..T4 by warp 0 of CTA 0, F0..F2 by the factoring warp;
// then 4 per merge level and the end), 8 * (nt + 8) values; nt = 4 * ceil(n / 128) tile columns
static long long* g_diag_dbg_host = nullptr;
int ace_dbg_diag_block_timeline(long long* stamps, int count) {
  if (!stamps || !g_diag_dbg_host) return usage("dbg_diag_block_timeline: run ace_dbg_diag_block first");
  for (int i = 0; i < count && i < 8 * 40; ++i) stamps[i] = g_diag_dbg_host[i];
  return 0;
}

int ace_dbg_diag_block(const double* A, int n, double* X, double* U, double* diagL, double* Ldiag_tiles) {
  if (!A || n < 1) return usage("dbg_diag_block: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, true, false));
  if (c.n_pad / TB > c.panel_blocks || !c.diag_ws.p) return usage("dbg_diag_block: n exceeds one diagonal block");
  ACE_TRY(upload_matrix(c.A.p, c.n_pad, c.n_pad, A, n, n, c.st));
  if (c.n_pad > n) pad_identity_kernel<<<(c.n_pad - n + 255) / 256, 256, 0, c.st>>>(c.A.p, c.n_pad, n, c.n_pad);
  ACE_CUDA(cudaGetLastError());
  ACE_CUDA(cudaMemsetAsync(c.info.p, 0, sizeof(int), c.st));
  DenseWork w = c.dense(c.A.p, c.Bf.p);
  DBuf<long long> dbg;
  ACE_TRY(dbg.alloc(8 * 40));
  ACE_CUDA(cudaMemsetAsync(dbg.p, 0, sizeof(long long) * 8 * 40, c.st));
  w.diag_dbg = dbg.p;
  ACE_TRY(diag_block_factor_invert(w, 0, c.n_pad / TB, c.st));
  ACE_TRY(c.fetch_scalars());
  if (!g_diag_dbg_host) g_diag_dbg_host = new long long[8 * 40];
  ACE_CUDA(cudaMemcpy(g_diag_dbg_host, dbg.p, sizeof(long long) * 8 * 40, cudaMemcpyDeviceToHost));
  const size_t N = c.n_pad;
  std::vector<double> hA(N * N), hDX(N * TB), hDU(N * TB), hd(N);
  ACE_CUDA(cudaMemcpy(hA.data(), c.A.p, sizeof(double) * N * N, cudaMemcpyDeviceToHost));
  ACE_CUDA(cudaMemcpy(hDX.data(), c.DX.p, sizeof(double) * N * TB, cudaMemcpyDeviceToHost));
  ACE_CUDA(cudaMemcpy(hDU.data(), c.DU.p, sizeof(double) * N * TB, cudaMemcpyDeviceToHost));
  ACE_CUDA(cudaMemcpy(hd.data(), c.dvec.p, sizeof(double) * N, cudaMemcpyDeviceToHost));
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      const int bi = i / TB, bj = j / TB;
      const size_t t = (size_t)bi * TB * TB + (size_t)(j % TB) * TB + (i % TB);
      const double a = hA[i + (size_t)j * N];
      if (X) X[i + (size_t)j * n] = (bi == bj) ? hDX[t] : (bi > bj ? a : 0.0);
      if (U) U[i + (size_t)j * n] = (bi == bj) ? hDU[t] : (bi < bj ? a : 0.0);
      if (Ldiag_tiles) Ldiag_tiles[i + (size_t)j * n] = (bi == bj && i >= j) ? a : 0.0;
    }
  if (diagL) std::memcpy(diagL, hd.data(), sizeof(double) * n);
  if (c.h_info() > 0) return c.h_info();
  return 0;
}

int ace_dbg_set_trtri_max_h(int h) {
  dbg_trtri_max_h() = h;
  return 0;
}

// raw state after potrf + trtri: rawA (n_pad x n_pad: X lower / U upper), DX, DU (nb tiles each)
int ace_dbg_trtri_raw(const double* A, int n, double* rawA, double* DXo, double* DUo, double* rawBf) {
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, true, false));
  ACE_TRY(upload_matrix(c.A.p, c.n_pad, c.n_pad, A, n, n, c.st));
  if (c.n_pad > n) pad_identity_kernel<<<(c.n_pad - n + 255) / 256, 256, 0, c.st>>>(c.A.p, c.n_pad, n, c.n_pad);
  DenseWork w = c.dense(c.A.p, c.Bf.p, /*allow_fused=*/false);
  ACE_TRY(potrf_blocked(w));
  ACE_TRY(trtri_merge(w));
  ACE_CUDA(cudaStreamSynchronize(c.st));
  const size_t N = c.n_pad;
  ACE_CUDA(cudaMemcpy(rawA, c.A.p, sizeof(double) * N * N, cudaMemcpyDeviceToHost));
  ACE_CUDA(cudaMemcpy(DXo, c.DX.p, sizeof(double) * N * TB, cudaMemcpyDeviceToHost));
  ACE_CUDA(cudaMemcpy(DUo, c.DU.p, sizeof(double) * N * TB, cudaMemcpyDeviceToHost));
  if (rawBf) ACE_CUDA(cudaMemcpy(rawBf, c.Bf.p, sizeof(double) * N * N, cudaMemcpyDeviceToHost));
  return 0;
}

int ace_bench_dense(int n, int reps, double* ms3) {
  if (n < 1 || reps < 1 || !ms3) return usage("bench_dense: bad argument");
  Core c;
  ACE_TRY(c.init(g_device, n, 1, 1, 0, true, false));
  DenseWork w = c.dense(c.A.p, c.Bf.p);
  const char* sep = std::getenv("ACE_DENSE_SEPARATE");  // unset/1: potrf and trtri timed separately; 0: overlapped
  const bool separate = !(sep && std::atoi(sep) == 0);
  double acc[3] = {0, 0, 0};
  for (int r = 0; r <= reps; ++r) {
    const size_t total = (size_t)c.n_pad * c.n_pad;
    synth_spd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(c.A.p, c.n_pad, c.n_pad);
    ACE_CUDA(cudaGetLastError());
    PotrfTrace tr;
    const bool tracing = std::getenv("ACE_POTRF_TRACE") && r == reps;
    if (tracing) {
      w.trace = &tr;
      cudaEventCreate(&tr.t0);
      cudaEventRecord(tr.t0, c.st);
    }
    ACE_CUDA(cudaEventRecord(c.tev[0], c.st));
    if (separate) {
      ACE_TRY(potrf_blocked(w));
      if (tracing) {
        ACE_CUDA(cudaStreamSynchronize(c.st));
        auto at = [&](cudaEvent_t e) { float t = 0; cudaEventElapsedTime(&t, tr.t0, e); return t; };
        std::printf("J  panel_begin panel_end | upd_begin upda_end updb_end   (ms since start)\n");
        for (size_t k = 0; k < tr.panel_begin.size(); ++k) {
          std::printf("%2zu  %8.3f %8.3f |", k, at(tr.panel_begin[k]), at(tr.panel_end[k]));
          if (k < tr.upd_begin.size()) std::printf(" %8.3f %8.3f %8.3f", at(tr.upd_begin[k]), at(tr.upda_end[k]), k < tr.updb_end.size() ? at(tr.updb_end[k]) : -1.f);
          std::printf("\n");
        }
        w.trace = nullptr;
      }
      ACE_CUDA(cudaEventRecord(c.tev[1], c.st));
      ACE_TRY(trtri_merge(w));
    } else {  // production schedule: leading block's inverse overlapped with the potrf tail
      ACE_TRY(potrf_trtri(w));
      ACE_CUDA(cudaEventRecord(c.tev[1], c.st));
    }
    ACE_CUDA(cudaEventRecord(c.tev[2], c.st));
    ACE_TRY(uut_inverse(w));
    ACE_CUDA(cudaEventRecord(c.tev[3], c.st));
    ACE_TRY(c.fetch_scalars());
    if (c.h_info() > 0) return c.h_info();
    if (r > 0) {
      for (int k = 0; k < 3; ++k) {
        float t;
        ACE_CUDA(cudaEventElapsedTime(&t, c.tev[k], c.tev[k + 1]));
        acc[k] += t;
      }
    }
  }
  for (int k = 0; k < 3; ++k) ms3[k] = acc[k] / reps;
  return 0;
}

}  // extern "C"

#include "host_utils.inl"
#include "prep_functions.inl"
