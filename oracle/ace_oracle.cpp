// oracle/ace_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the reference's
// empirical-Bayes GP hot path (R package `ace` 0.4.1).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product path (additivecausalexpansion_b200/) never
// does and has no CPU fallback.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
//  and cannot be compiled here (needs R, Rcpp, RcppArmadillo).
// Trust comes from (i) an independent NumPy restatement (oracle/np_oracle.py)
// that must agree with this file, (ii) an extended-precision evaluation at small
// n, (iii) analytic identities that hold for the literal code (tests/).
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference).  All arrays are FP64 column-major like R / Armadillo; the
// literal loop structure and association order is kept wherever it is cheap,
// including the reference's quirks: the off-by-one
// length-scale index in every kernel build, the -0.25*9 Matern constant,
// y'alpha (not ybar'alpha) in the evidence, 0.5 in the mu solution,
// clip-to-unit-norm, stale inverse for the mu refresh.
//
// Third-party arithmetic the reference reaches through Armadillo
// (RcppArmadillo >= 0.8.200.0, DESCRIPTION:27, not vendored): LAPACK dsyevd
// (arma::eig_sym default "dc"), BLAS dsyrk/dgemm/dgemv.  Here they come from the
// OpenBLAS 0.3.30 bundled with SciPy in this image (symbols scipy_dsyevd_ ...),
// loaded at run time with dlopen; the path is handed in by oracle/__init__.py.

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <vector>

namespace {

typedef void (*dsyevd_fn)(const char*, const char*, const int*, double*, const int*, double*,
                          double*, const int*, int*, const int*, int*);
typedef void (*dgemm_fn)(const char*, const char*, const int*, const int*, const int*, const double*,
                         const double*, const int*, const double*, const int*, const double*, double*,
                         const int*);
typedef void (*dgemv_fn)(const char*, const int*, const int*, const double*, const double*, const int*,
                         const double*, const int*, const double*, double*, const int*);
typedef void (*dsyrk_fn)(const char*, const char*, const int*, const int*, const double*, const double*,
                         const int*, const double*, double*, const int*);
typedef void (*dpotrf_fn)(const char*, const int*, double*, const int*, int*);
typedef void (*dpotri_fn)(const char*, const int*, double*, const int*, int*);
typedef void (*setthr_fn)(int);
typedef int (*getthr_fn)(void);

dsyevd_fn p_dsyevd = nullptr;
dgemm_fn p_dgemm = nullptr;
dgemv_fn p_dgemv = nullptr;
dsyrk_fn p_dsyrk = nullptr;
dpotrf_fn p_dpotrf = nullptr;
dpotri_fn p_dpotri = nullptr;
setthr_fn p_setthr = nullptr;
getthr_fn p_getthr = nullptr;

inline double sgn(double x) {  // src/include/ace_kernel_utils.hpp:38-40
  return (double)((0 < x) - (x < 0));
}

// y = A x  (A m x n col-major), via BLAS dgemv like Armadillo's mat*vec.
void gemv_n(const double* A, int m, int n, const double* x, double* y) {
  const double one = 1.0, zero = 0.0;
  const int inc = 1;
  p_dgemv("N", &m, &n, &one, A, &m, x, &inc, &zero, y, &inc);
}

// src/include/ace_kernel_utils.hpp:7-20  uppertri2symmat
void uppertri2symmat(const double* matvec, size_t dim, double* out) {
  size_t cnt = 0;
  for (size_t r = 0; r < dim; r++) {
    for (size_t c = r; c < dim; c++) {
      out[r + dim * c] = out[c + dim * r] = matvec[cnt];
      cnt++;
    }
  }
}

// src/include/ace_kernel_utils.hpp:23-26  evid_grad = -0.5 * trace(Kaa * dK).
// Armadillo evaluates trace(A*B) without forming the product:
// sum_k dot(A.row(k), B.col(k)).
double evid_grad(const double* Kaa, const double* dK, size_t n) {
  double acc = 0.0;
  for (size_t k = 0; k < n; k++) {
    const double* bcol = dK + n * k;
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += Kaa[k + n * i] * bcol[i];
    acc += s;
  }
  return -0.5 * acc;
}

// src/include/ace_kernel_utils.hpp:29-31
double sigma_gradient(const double* Kaa, size_t n, double sigma) {
  double tr = 0.0;
  for (size_t i = 0; i < n; i++) tr += Kaa[i + n * i];
  return -0.5 * tr * std::exp(sigma);
}

// src/include/ace_kernel_utils.hpp:33-36
double logevidence(const double* y, const double* alpha, const double* eigenval, size_t n) {
  double sl = 0.0, dot = 0.0;
  for (size_t i = 0; i < n; i++) sl += std::log(eigenval[i]);
  for (size_t i = 0; i < n; i++) dot += y[i] * alpha[i];
  return -0.5 * ((double)n * std::log(2.0 * M_PI) + sl + dot);
}

bool all_finite(const double* g, size_t P) {
  for (size_t i = 0; i < P; i++)
    if (!std::isfinite(g[i])) return false;
  return true;
}

}  // namespace

extern "C" {

// Load BLAS/LAPACK entry points from the SciPy-bundled OpenBLAS.
int ace_oracle_init(const char* openblas_path, int nthreads) {
  void* h = dlopen(openblas_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    std::fprintf(stderr, "ace_oracle_init: dlopen failed: %s\n", dlerror());
    return -1;
  }
  p_dsyevd = (dsyevd_fn)dlsym(h, "scipy_dsyevd_");
  p_dgemm = (dgemm_fn)dlsym(h, "scipy_dgemm_");
  p_dgemv = (dgemv_fn)dlsym(h, "scipy_dgemv_");
  p_dsyrk = (dsyrk_fn)dlsym(h, "scipy_dsyrk_");
  p_dpotrf = (dpotrf_fn)dlsym(h, "scipy_dpotrf_");
  p_dpotri = (dpotri_fn)dlsym(h, "scipy_dpotri_");
  p_setthr = (setthr_fn)dlsym(h, "scipy_openblas_set_num_threads");
  p_getthr = (getthr_fn)dlsym(h, "scipy_openblas_get_num_threads");
  if (!p_dsyevd || !p_dgemm || !p_dgemv || !p_dsyrk || !p_dpotrf || !p_dpotri) return -2;
  if (nthreads > 0 && p_setthr) p_setthr(nthreads);
  return 0;
}

int ace_oracle_threads(void) { return p_getthr ? p_getthr() : 1; }

// ---------------------------------------------------------------------------
// Kernel builds
// ---------------------------------------------------------------------------

// src/kernel_SE_cpp.cpp:9-64  kernmat_SE_cpp (rectangular).
// full: n1 x n2; elements: n1 x n2 x B (B = Bz + 1).
void ace_oracle_kernmat_SE(const double* X1, const double* X2, const double* Z1, const double* Z2,
                           int n1, int n2, int p, int Bz, const double* par, double* full,
                           double* elements) {
  const size_t N1 = n1, N2 = n2, B = (size_t)Bz + 1;
  double* tmpX = elements;
  std::fill(tmpX, tmpX + N1 * N2 * B, 0.0);
  std::vector<double> tmprow(N2);
  for (size_t i = 0; i < (size_t)p; i++) {        // :27
    for (size_t r = 0; r < N1; r++) {             // :28
      for (size_t c = 0; c < N2; c++) {           // :30
        double d = X1[r + N1 * i] - X2[c + N2 * i];
        tmprow[c] = d * d;
      }
      for (size_t b = 0; b < B; b++) {            // :32-34  (index quirk Q1)
        const double w = std::exp(-par[1 + b + B * (i + 1)]);
        double* dst = tmpX + N1 * N2 * b + r;
        for (size_t c = 0; c < N2; c++) dst[N1 * c] += tmprow[c] * w;
      }
    }
  }
  for (size_t k = 0; k < N1 * N2; k++) tmpX[k] = std::exp(par[2] - tmpX[k]);  // :39
  std::memcpy(full, tmpX, sizeof(double) * N1 * N2);                          // :40
  for (size_t b = 1; b < B; b++) {                                            // :42
    double* sl = tmpX + N1 * N2 * b;
    for (size_t r = 0; r < N1; r++) {
      const double z1 = Z1[r + N1 * (b - 1)];
      if (z1 == 0) {                                                          // :44-47
        for (size_t c = 0; c < N2; c++) sl[r + N1 * c] = 0.0;
        continue;
      }
      for (size_t c = 0; c < N2; c++) {
        const double z2 = Z2[c + N2 * (b - 1)];
        if (z2 == 0) {                                                        // :49-51
          sl[r + N1 * c] = 0.0;
        } else {                                                              // :53
          sl[r + N1 * c] = (sgn(z1) * sgn(z2)) *
                           std::exp(par[2 + b] - sl[r + N1 * c] + std::log(std::abs(z1)) +
                                    std::log(std::abs(z2)));
        }
      }
    }
    for (size_t k = 0; k < N1 * N2; k++) full[k] += sl[k];                    // :60
  }
}

// src/kernel_SE_cpp.cpp:67-134  kernmat_SE_symmetric_cpp.
// The reference indexes the packed triangle with `unsigned int` (overflow at
// n >= 65536, :71,78); size_t is used here, sizes that overflow are never run.
void ace_oracle_kernmat_SE_sym(const double* X, const double* Z, int n_, int p, int Bz,
                               const double* par, double* full, double* elements) {
  const size_t n = n_, B = (size_t)Bz + 1;
  const size_t T = n * (n + 1) / 2;
  std::vector<double> tmpX(T * B, 0.0);  // :78  packed upper triangle, one column per term
  for (size_t i = 0; i < (size_t)p; i++) {  // :82
    size_t cnt = 0;
    for (size_t r = 0; r < n; r++) {
      for (size_t c = r; c < n; c++) {
        double d = X[r + n * i] - X[c + n * i];
        const double tmp = d * d;  // pow(.,2)
        for (size_t b = 0; b < B; b++) {
          tmpX[cnt + T * b] += tmp * std::exp(-par[1 + b + B * (i + 1)]);  // :89 (Q1)
        }
        cnt++;
      }
    }
  }
  for (size_t k = 0; k < T; k++) tmpX[k] = std::exp(par[2] - tmpX[k]);  // :96
  uppertri2symmat(tmpX.data(), n, elements);                            // :97
  for (size_t b = 1; b < B; b++) {                                      // :98
    size_t cnt = 0;
    double* col = tmpX.data() + T * b;
    for (size_t r = 0; r < n; r++) {
      const double zr = Z[r + n * (b - 1)];
      if (zr == 0) {  // :103-108
        for (size_t k = 0; k < n - r; k++) col[cnt + k] = 0.0;
        cnt += (n - r);
        continue;
      }
      for (size_t c = r; c < n; c++) {
        const double zc = Z[c + n * (b - 1)];
        if (zc == 0) {  // :111-115
          col[cnt] = 0.0;
          cnt++;
          continue;
        }
        col[cnt] = (sgn(zr) * sgn(zc)) *
                   std::exp(par[2 + b] - col[cnt] + std::log(std::abs(zr)) + std::log(std::abs(zc)));  // :119
        cnt++;
      }
    }
    uppertri2symmat(col, n, elements + n * n * b);      // :125
    for (size_t j = 0; j < T; j++) tmpX[j] += col[j];   // :126-128
  }
  uppertri2symmat(tmpX.data(), n, full);                // :130
}

// src/kernel_Matern_cpp.cpp:52-93  kernmat_Matern32_cpp (rectangular).
void ace_oracle_kernmat_Matern32(const double* X1, const double* X2, const double* Z1,
                                 const double* Z2, int n1, int n2, int p, int Bz, const double* par,
                                 double* full, double* elements) {
  const size_t N1 = n1, N2 = n2, B = (size_t)Bz + 1;
  double* tmpX = elements;
  std::fill(tmpX, tmpX + N1 * N2 * B, 0.0);
  std::vector<double> tmprow(N2);
  for (size_t i = 0; i < (size_t)p; i++) {  // :66
    for (size_t r = 0; r < N1; r++) {
      for (size_t c = 0; c < N2; c++) {
        double d = X1[r + N1 * i] - X2[c + N2 * i];
        tmprow[c] = d * d;
      }
      for (size_t b = 0; b < B; b++) {  // :71-73 (Q1)
        const double w = std::exp(-par[1 + b + B * (i + 1)]);
        double* dst = tmpX + N1 * N2 * b + r;
        for (size_t c = 0; c < N2; c++) dst[N1 * c] += tmprow[c] * w;
      }
    }
  }
  for (size_t k = 0; k < N1 * N2 * B; k++) tmpX[k] = std::sqrt(tmpX[k]);  // :76
  const double s3 = std::sqrt(3.0);
  for (size_t k = 0; k < N1 * N2; k++)  // :79
    tmpX[k] = (1 + s3 * tmpX[k]) * std::exp(par[2] - s3 * tmpX[k]);
  std::memcpy(full, tmpX, sizeof(double) * N1 * N2);  // :80
  for (size_t b = 1; b < B; b++) {                    // :81
    double* sl = tmpX + N1 * N2 * b;
    for (size_t r = 0; r < N1; r++) {
      const double z1 = Z1[r + N1 * (b - 1)];
      if (z1 == 0) {  // :83
        for (size_t c = 0; c < N2; c++) sl[r + N1 * c] = 0.0;
        continue;
      }
      for (size_t c = 0; c < N2; c++) {
        const double z2 = Z2[c + N2 * (b - 1)];
        if (z2 == 0) {  // :85
          sl[r + N1 * c] = 0.0;
          continue;
        }
        const double t = sl[r + N1 * c];
        sl[r + N1 * c] = (1 + s3 * t) * std::exp(par[2 + b] - s3 * t) * z1 * z2;  // :86
      }
    }
    for (size_t k = 0; k < N1 * N2; k++) full[k] += sl[k];  // :89
  }
}

// src/kernel_Matern_cpp.cpp:190-240  kernmat_Matern32_symmetric_cpp.
void ace_oracle_kernmat_Matern32_sym(const double* X, const double* Z, int n_, int p, int Bz,
                                     const double* par, double* full, double* elements) {
  const size_t n = n_, B = (size_t)Bz + 1;
  const size_t T = n * (n + 1) / 2;
  std::vector<double> tmpX(T * B, 0.0);
  for (size_t i = 0; i < (size_t)p; i++) {  // :203
    size_t cnt = 0;
    for (size_t r = 0; r < n; r++) {
      for (size_t c = r; c < n; c++) {
        double d = X[r + n * i] - X[c + n * i];
        const double tmp = d * d;
        for (size_t b = 0; b < B; b++) {
          tmpX[cnt + T * b] += tmp * std::exp(-par[1 + b + B * (i + 1)]);  // :210 (Q1)
        }
        cnt++;
      }
    }
  }
  for (size_t k = 0; k < T * B; k++) tmpX[k] = std::sqrt(tmpX[k]);  // :215
  const double s3 = std::sqrt(3.0);
  for (size_t k = 0; k < T; k++)  // :217
    tmpX[k] = (1 + s3 * tmpX[k]) * std::exp(par[2] - s3 * tmpX[k]);
  uppertri2symmat(tmpX.data(), n, elements);  // :218
  for (size_t b = 1; b < B; b++) {            // :219
    size_t cnt = 0;
    double* col = tmpX.data() + T * b;
    for (size_t r = 0; r < n; r++) {
      const double zr = Z[r + n * (b - 1)];
      if (zr == 0) {  // :223-225
        for (size_t k = 0; k < n - r; k++) col[cnt + k] = 0.0;
        cnt += n - r;
        continue;
      }
      for (size_t c = r; c < n; c++) {  // :226-229 (no test on Z(c,.): the product gives 0)
        const double t = col[cnt];
        col[cnt] = (1 + s3 * t) * std::exp(par[2 + b] - s3 * t) * zr * Z[c + n * (b - 1)];
        cnt++;
      }
    }
    uppertri2symmat(col, n, elements + n * n * b);     // :232
    for (size_t j = 0; j < T; j++) tmpX[j] += col[j];  // :233
  }
  uppertri2symmat(tmpX.data(), n, full);  // :235
}

// ---------------------------------------------------------------------------
// "Inverse" by eigendecomposition
// ---------------------------------------------------------------------------

// src/kernel_SE_cpp.cpp:137-157  invkernel_cpp.
// K: n x n (not modified); eigenval: n ascending; inv: n x n.
// Returns LAPACK info (the reference only prints on failure and carries on, :144-146).
int ace_oracle_invkernel(const double* K, int n_, double sigma, double* eigenval, double* inv) {
  const size_t n = n_;
  std::vector<double> V(K, K + n * n);  // by-value argument copy, :137
  const double es = std::exp(sigma);
  for (size_t i = 0; i < n; i++) V[i + n * i] += es;  // :142
  int info = 0, lwork = -1, liwork = -1, iwq = 0;
  double wq = 0;
  p_dsyevd("V", "U", &n_, V.data(), &n_, eigenval, &wq, &lwork, &iwq, &liwork, &info);
  lwork = (int)wq;
  liwork = iwq;
  std::vector<double> work((size_t)lwork);
  std::vector<int> iwork((size_t)liwork);
  p_dsyevd("V", "U", &n_, V.data(), &n_, eigenval, work.data(), &lwork, iwork.data(), &liwork, &info);
  if (info != 0) std::fprintf(stderr, "Eigenvalue decomp. not completed.\n");  // :145
  for (size_t i = 0; i < n; i++) {  // :150-152
    const double s = std::sqrt(eigenval[i]);
    double* col = V.data() + n * i;
    for (size_t r = 0; r < n; r++) col[r] = col[r] / s;
  }
  // :153  pdmat * pdmat.t()  -> Armadillo dispatches A*A' to dsyrk and mirrors.
  const double one = 1.0, zero = 0.0;
  p_dsyrk("U", "N", &n_, &n_, &one, V.data(), &n_, &zero, inv, &n_);
  for (size_t c = 0; c < n; c++)
    for (size_t r = c + 1; r < n; r++) inv[r + n * c] = inv[c + n * r];
  return info;
}

// CPU-favourable variant (NOT what the reference does): Cholesky dpotrf + dpotri.
// eigenval slot receives diag(L)^2 so that sum(log) is still log det.
int ace_oracle_invkernel_chol(const double* K, int n_, double sigma, double* eigenval, double* inv) {
  const size_t n = n_;
  std::memcpy(inv, K, sizeof(double) * n * n);
  const double es = std::exp(sigma);
  for (size_t i = 0; i < n; i++) inv[i + n * i] += es;
  int info = 0;
  p_dpotrf("L", &n_, inv, &n_, &info);
  if (info != 0) return info;
  for (size_t i = 0; i < n; i++) eigenval[i] = inv[i + n * i] * inv[i + n * i];
  p_dpotri("L", &n_, inv, &n_, &info);
  for (size_t c = 0; c < n; c++)
    for (size_t r = c + 1; r < n; r++) inv[c + n * r] = inv[r + n * c];
  return info;
}

// ---------------------------------------------------------------------------
// Gradients and statistics
// ---------------------------------------------------------------------------

// src/kernel_SE_cpp.cpp:161-188 evid_scale_gradients + :192-243 grad_SE_cpp.
// Kcube: n x n x B.  stats[2] written in place.  grad: P = 2 + B + B*p.
void ace_oracle_grad_SE(const double* y, const double* X, const double* Kfull, const double* Kcube,
                        const double* invK, const double* eigenval, const double* par, int n_, int p_,
                        int B_, double std_y, double* stats, double* grad) {
  const size_t n = n_, p = p_, B = B_;
  const size_t P = 2 + B + B * p;
  std::fill(grad, grad + P, 0.0);
  std::vector<double> ybar(n), alpha(n), tmpK(n * n);
  for (size_t i = 0; i < n; i++) ybar[i] = y[i] - par[1];  // :212
  gemv_n(invK, n_, n_, ybar.data(), alpha.data());         // :215
  for (size_t c = 0; c < n; c++)                           // :218
    for (size_t r = 0; r < n; r++) tmpK[r + n * c] = invK[r + n * c] - alpha[r] * alpha[c];
  grad[0] = sigma_gradient(tmpK.data(), n, par[0]);        // :221
  for (size_t b = 0; b < B; b++)                           // :224-227
    grad[2 + b] = evid_grad(tmpK.data(), Kcube + n * n * b, n);
  // :230-231 -> evid_scale_gradients(X, tmpK, K, L = par[2+B ..], B)   (:161-188)
  {
    std::vector<double> tmpX(n * n), dK(n * n);
    const double* L = par + 2 + B;
    for (size_t i = 0; i < p; i++) {
      for (size_t r = 0; r < n; r++)  // :173-175
        for (size_t k = 0; k < n; k++) {
          double d = X[r + n * i] - X[k + n * i];
          tmpX[k + n * r] = d * d;
        }
      for (size_t b = 0; b < B; b++) {  // :176-179: the `const arma::mat& dK` parameter materialises
        const double e = std::exp(-L[b + B * i]);
        const double* Kb = Kcube + n * n * b;
        for (size_t k = 0; k < n * n; k++) dK[k] = Kb[k] * tmpX[k] * e;
        grad[2 + B + b + B * i] = evid_grad(tmpK.data(), dK.data(), n);
      }
    }
  }
  {  // :234  mu gradient
    std::vector<double> t(n);
    gemv_n(invK, n_, n_, ybar.data(), t.data());
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += t[i];
    grad[1] = s;
  }
  {  // :238 RMSE, :240 evidence
    std::vector<double> Ka(n);
    gemv_n(Kfull, n_, n_, alpha.data(), Ka.data());
    double ss = 0.0;
    for (size_t i = 0; i < n; i++) {
      double r = ybar[i] - Ka[i];
      ss += r * r;
    }
    stats[0] = std_y * std::sqrt(ss) / std::sqrt((double)n);
    stats[1] = logevidence(y, alpha.data(), eigenval, n);
  }
}

// src/kernel_Matern_cpp.cpp:340-377 evid_scale_Matern32_gradients + :420-467 grad_Matern_cpp.
void ace_oracle_grad_Matern(const double* y, const double* X, const double* Kfull, const double* Kcube,
                            const double* invK, const double* eigenval, const double* par, int n_,
                            int p_, int B_, double std_y, double* stats, double* grad) {
  const size_t n = n_, p = p_, B = B_;
  const size_t P = 2 + B + B * p;
  std::fill(grad, grad + P, 0.0);
  std::vector<double> ybar(n), alpha(n), tmpK(n * n);
  for (size_t i = 0; i < n; i++) ybar[i] = y[i] - par[1];  // :441
  gemv_n(invK, n_, n_, ybar.data(), alpha.data());         // :442
  for (size_t c = 0; c < n; c++)                           // :443
    for (size_t r = 0; r < n; r++) tmpK[r + n * c] = invK[r + n * c] - alpha[r] * alpha[c];
  grad[0] = sigma_gradient(tmpK.data(), n, par[0]);        // :446
  for (size_t b = 0; b < B; b++)                           // :449-452
    grad[2 + b] = evid_grad(tmpK.data(), Kcube + n * n * b, n);
  {  // :455 -> evid_scale_Matern32_gradients (:340-377)
    const double* L = par + 2 + B;
    std::vector<double> tmpX(n * n * B, 0.0), tmpX2(n * n), dK(n * n), tmprow(n);
    for (size_t i = 0; i < p; i++) {  // :351-359  distance rebuilt with the GRADIENT indexing (Q2)
      for (size_t r = 0; r < n; r++) {
        for (size_t c = 0; c < n; c++) {
          double d = X[r + n * i] - X[c + n * i];
          tmprow[c] = d * d;
        }
        for (size_t b = 0; b < B; b++) {
          const double w = std::exp(-L[b + B * i]);
          double* dst = tmpX.data() + n * n * b + r;
          for (size_t c = 0; c < n; c++) dst[n * c] += tmprow[c] * w;
        }
      }
    }
    for (size_t b = 0; b < B; b++) {  // :362-364
      double* sl = tmpX.data() + n * n * b;
      const double* Kb = Kcube + n * n * b;
      for (size_t k = 0; k < n * n; k++) sl[k] = Kb[k] / (1 + std::sqrt(3 * sl[k]));
    }
    for (size_t i = 0; i < p; i++) {  // :367-375
      for (size_t r = 0; r < n; r++)
        for (size_t k = 0; k < n; k++) {
          double d = X[r + n * i] - X[k + n * i];
          tmpX2[k + n * r] = d * d;
        }
      for (size_t b = 0; b < B; b++) {
        const double* sl = tmpX.data() + n * n * b;
        for (size_t k = 0; k < n * n; k++) dK[k] = sl[k] * tmpX2[k];
        // - 0.25 * 9 * trace(Kaa * (.)) * exp(-L)   (:373)
        const double tr = -2.0 * evid_grad(tmpK.data(), dK.data(), n);
        grad[2 + B + b + B * i] = -0.25 * 9 * tr * std::exp(-L[b + B * i]);
      }
    }
  }
  grad[1] = 0;  // :458
  {             // :461-463
    std::vector<double> Ka(n);
    gemv_n(Kfull, n_, n_, alpha.data(), Ka.data());
    double ss = 0.0;
    for (size_t i = 0; i < n; i++) {
      double r = ybar[i] - Ka[i];
      ss += r * r;
    }
    stats[0] = std_y * std::sqrt(ss) / std::sqrt((double)n);
    stats[1] = logevidence(y, alpha.data(), eigenval, n);
  }
}

// src/stats_cpp.cpp:9-32
void ace_oracle_stats(const double* y, const double* Kmat, const double* invK, const double* eigenval,
                      double mu, double std_y, int n_, double* stats) {
  const size_t n = n_;
  std::vector<double> ybar(n), alpha(n), Ka(n);
  for (size_t i = 0; i < n; i++) ybar[i] = y[i] - mu;  // :22
  gemv_n(invK, n_, n_, ybar.data(), alpha.data());     // :23
  gemv_n(Kmat, n_, n_, alpha.data(), Ka.data());
  double ss = 0.0;
  for (size_t i = 0; i < n; i++) {
    double r = ybar[i] - Ka[i];
    ss += r * r;
  }
  stats[0] = std_y * std::sqrt(ss) / std::sqrt((double)n);  // :26
  stats[1] = logevidence(y, alpha.data(), eigenval, n);     // :29
}

// src/utilities_cpp.cpp:6-10
double ace_oracle_mu_solution(const double* y, const double* invK, int n_) {
  const size_t n = n_;
  std::vector<double> t(n);
  gemv_n(invK, n_, n_, y, t.data());
  double s = 0.0, a = 0.0;
  for (size_t i = 0; i < n; i++) s += t[i];
  for (size_t k = 0; k < n * n; k++) a += invK[k];
  return 0.5 * s / a;
}

// src/utilities_cpp.cpp:121-129
void ace_oracle_norm_clip(int flag, double* grads, int P, double max_length) {
  if (flag) {
    double ss = 0.0;
    for (int i = 0; i < P; i++) ss += grads[i] * grads[i];
    const double L2 = std::sqrt(ss);
    if ((L2 > max_length) & std::isfinite(L2) & (L2 != 0)) {
      for (int i = 0; i < P; i++) grads[i] = grads[i] / L2;
    }
  }
}

// src/optimizer_cpp.cpp:8-20
int ace_oracle_Nesterov(double lr, double momentum, double* nu, const double* grad, double* para, int P) {
  const bool ok = all_finite(grad, P);
  for (int i = 0; i < P; i++) {
    nu[i] = momentum * nu[i] + lr * grad[i];
    para[i] = para[i] + nu[i];
  }
  return ok ? 1 : 0;
}

// src/optimizer_cpp.cpp:23-42
int ace_oracle_Nadam(double iter, double lr, double beta1, double beta2, double eps, double* m,
                     double* v, const double* grad, double* para, int P) {
  const bool ok = all_finite(grad, P);
  const double c1 = 1 - std::pow(beta1, iter), c2 = 1 - std::pow(beta2, iter);
  for (int i = 0; i < P; i++) {
    m[i] = beta1 * m[i] + (1 - beta1) * grad[i];
    v[i] = beta2 * v[i] + (1 - beta2) * (grad[i] * grad[i]);
    para[i] = para[i] + lr * ((beta1 * m[i] + (1 - beta1) * grad[i]) / c1) / (std::sqrt(v[i] / c2) + eps);
  }
  return ok ? 1 : 0;
}

// src/optimizer_cpp.cpp:45-63
int ace_oracle_Adam(double iter, double lr, double beta1, double beta2, double eps, double* m, double* v,
                    const double* grad, double* para, int P) {
  const bool ok = all_finite(grad, P);
  const double c1 = 1 - std::pow(beta1, iter), c2 = 1 - std::pow(beta2, iter);
  for (int i = 0; i < P; i++) {
    m[i] = (beta1 * m[i]) + (1 - beta1) * grad[i];
    v[i] = beta2 * v[i] + (1 - beta2) * (grad[i] * grad[i]);
    para[i] = para[i] + lr * (m[i] / c1) / (std::sqrt(v[i] / c2) + eps);
  }
  return ok ? 1 : 0;
}

// ---------------------------------------------------------------------------
// Posterior
// ---------------------------------------------------------------------------

// src/pred_cpp.cpp:8-34  pred_cpp.   K_xX: nx x nX, K_xx: nx x nx.
// map: nx, ci: nx x 2, var: nx.
void ace_oracle_pred(const double* y_X, double sigma, double mu, const double* invK, const double* K_xX,
                     const double* K_xx, double mean_y, double std_y, int nx_, int nX_, double* map,
                     double* ci, double* var) {
  const size_t nx = nx_, nX = nX_;
  std::vector<double> tmp(nx * nX), yb(nX), t(nx), C(K_xx, K_xx + nx * nx);
  const double one = 1.0, zero = 0.0, mone = -1.0;
  p_dgemm("N", "N", &nx_, &nX_, &nX_, &one, K_xX, &nx_, invK, &nX_, &zero, tmp.data(), &nx_);  // :19
  for (size_t i = 0; i < nX; i++) yb[i] = y_X[i] - mu;
  gemv_n(tmp.data(), nx_, nX_, yb.data(), t.data());
  for (size_t i = 0; i < nx; i++) map[i] = mean_y + std_y * (t[i] + mu);  // :20
  p_dgemm("N", "T", &nx_, &nx_, &nX_, &mone, tmp.data(), &nx_, K_xX, &nx_, &one, C.data(), &nx_);  // :22
  const double es = std::exp(sigma);
  for (size_t i = 0; i < nx; i++) C[i + nx * i] += es;  // :23
  for (size_t i = 0; i < nx; i++) {
    const double sd = std_y * std::sqrt(std::abs(C[i + nx * i]));  // :26
    ci[i] = map[i] - 1.96 * sd;                                    // :27
    ci[i + nx] = map[i] + 1.96 * sd;                               // :28
    var[i] = sd * sd;                                              // :29
  }
}

// src/pred_cpp.cpp:37-126  pred_marginal_cpp.
// K_xX cube: nx x nX x B, K_xx cube: nx x nx x B.  avg[9] = {ate map, ci lo, ci hi, var,
// att ..., atu ...} laid out as 3 groups of {map, ci0, ci1, var} -> 12 doubles.
void ace_oracle_pred_marginal(const double* y_X, const double* Z_x, double sigma, double mu,
                              const double* invK, const double* K_xX, const double* K_xx, double mean_y,
                              double std_y, double std_Z, int calculate_ate, int nx_, int nX_, int B_,
                              double* map, double* ci, double* var, double* avg) {
  (void)sigma;
  (void)mean_y;
  const size_t nx = nx_, nX = nX_, B = B_;
  std::vector<double> KmxX(nx * nX), Kmxx(nx * nx), tmp(nx * nX), yb(nX), t(nx);
  if (B > 1) {  // :55-63
    std::memcpy(KmxX.data(), K_xX + nx * nX * 1, sizeof(double) * nx * nX);
    std::memcpy(Kmxx.data(), K_xx + nx * nx * 1, sizeof(double) * nx * nx);
    for (size_t b = 2; b < B; b++) {
      for (size_t k = 0; k < nx * nX; k++) KmxX[k] += K_xX[nx * nX * b + k];
      for (size_t k = 0; k < nx * nx; k++) Kmxx[k] += K_xx[nx * nx * b + k];
    }
  } else {  // :64-67
    std::memcpy(KmxX.data(), K_xX, sizeof(double) * nx * nX);
    std::memcpy(Kmxx.data(), K_xx, sizeof(double) * nx * nx);
  }
  const double one = 1.0, zero = 0.0, mone = -1.0;
  p_dgemm("N", "N", &nx_, &nX_, &nX_, &one, KmxX.data(), &nx_, invK, &nX_, &zero, tmp.data(), &nx_);  // :69
  for (size_t i = 0; i < nX; i++) yb[i] = y_X[i] - mu;
  // :70  std_y * tmp * (y_X - mu) / std_Z   evaluated left to right: (std_y*tmp)*(yb) / std_Z
  gemv_n(tmp.data(), nx_, nX_, yb.data(), t.data());
  for (size_t i = 0; i < nx; i++) map[i] = std_y * t[i] / std_Z;
  p_dgemm("N", "T", &nx_, &nx_, &nX_, &mone, tmp.data(), &nx_, KmxX.data(), &nx_, &one, Kmxx.data(), &nx_);  // :72
  for (size_t i = 0; i < nx; i++) {
    const double sd = std_y * std::sqrt(std::abs(Kmxx[i + nx * i])) / std_Z;  // :75
    ci[i] = map[i] - 1.96 * sd;
    ci[i + nx] = map[i] + 1.96 * sd;
    var[i] = sd * sd;
  }
  if (!calculate_ate) return;
  // :86-92 ATE
  double s = 0.0;
  for (size_t i = 0; i < nx; i++) s += map[i];
  const double ate = s / (double)nx;
  double acc = 0.0;
  for (size_t k = 0; k < nx * nx; k++) acc += Kmxx[k];
  double ate_sd = std_y * std::sqrt(acc) / (double)nx;
  avg[0] = ate;
  avg[1] = ate - 1.96 * ate_sd;
  avg[2] = ate + 1.96 * ate_sd;
  avg[3] = ate_sd * ate_sd;
  // :95-101 ATT
  double sz = 0.0;
  for (size_t i = 0; i < nx; i++) sz += Z_x[i];
  const unsigned int ntx = (unsigned int)sz;
  double dz = 0.0;
  for (size_t i = 0; i < nx; i++) dz += map[i] * Z_x[i];
  const double att = dz / ntx;
  std::vector<double> kz(nx);
  gemv_n(Kmxx.data(), nx_, nx_, Z_x, kz.data());
  double q = 0.0;
  for (size_t i = 0; i < nx; i++) q += kz[i] * Z_x[i];
  double att_sd = std_y * std::sqrt(q) / ntx;
  avg[4] = att;
  avg[5] = att - 1.96 * att_sd;
  avg[6] = att + 1.96 * att_sd;
  avg[7] = att_sd * att_sd;
  // :104-110 ATU
  const unsigned int nux = (unsigned int)nx - ntx;
  const double atu = (ate * (double)nx - att * ntx) / nux;
  std::vector<double> u(nx);
  for (size_t i = 0; i < nx; i++) u[i] = (Z_x[i] == 0) ? 1.0 : 0.0;
  gemv_n(Kmxx.data(), nx_, nx_, u.data(), kz.data());
  q = 0.0;
  for (size_t i = 0; i < nx; i++) q += kz[i] * u[i];
  double atu_sd = std_y * std::sqrt(q) / nux;
  avg[8] = atu;
  avg[9] = atu - 1.96 * atu_sd;
  avg[10] = atu + 1.96 * atu_sd;
  avg[11] = atu_sd * atu_sd;
}

// ---------------------------------------------------------------------------
// Natural cubic spline basis (input generation; "bit-exact on basis")
// ---------------------------------------------------------------------------

static std::vector<double> unique_sorted(const double* knots, int K) {
  // arma::unique returns sorted unique values (src/ncs_basis_cpp.cpp:65-68).
  std::vector<double> k(knots, knots + K);
  std::sort(k.begin(), k.end());
  k.erase(std::unique(k.begin(), k.end()), k.end());
  return k;
}

// Number of unique knots = number of basis columns.
int ace_oracle_ncs_ncol(const double* knots, int K) { return (int)unique_sorted(knots, K).size(); }

// src/ncs_basis_cpp.cpp:5-28 generate_ncs_matrix + :61-79 ncs_basis.  design: n x K.
void ace_oracle_ncs_basis(const double* x, int n_, const double* knots_in, int Kin, double* design) {
  const std::vector<double> knots = unique_sorted(knots_in, Kin);
  const size_t n = n_, K = knots.size();
  std::vector<double> d(n * K);
  // arma::pow(x - k, 3) is evaluated element by element with std::pow (Armadillo's eop_aux::pow); v * v * v rounds
  // twice and differs from it by 1 ulp for a sizeable fraction of inputs
  auto cube = [](double v) { return std::pow(v, 3.0); };
  for (size_t r = 0; r < n; r++)  // :15
    d[r + n * (K - 1)] = (x[r] > knots[K - 1] ? 1.0 : 0.0) * cube(x[r] - knots[K - 1]);
  for (size_t i = 0; i + 1 < K; i++) {  // :16-19
    for (size_t r = 0; r < n; r++) {
      double v = (x[r] > knots[i] ? 1.0 : 0.0) * cube(x[r] - knots[i]) - d[r + n * (K - 1)];
      d[r + n * i] = v / (knots[K - 1] - knots[i]);
    }
  }
  for (size_t r = 0; r < n; r++) d[r + n * (K - 1)] = 0.0;  // :21
  for (size_t r = 0; r < n; r++) design[r] = x[r];          // :76
  for (size_t i = 0; i + 2 < K; i++)                        // :23-25
    for (size_t r = 0; r < n; r++) design[r + n * (1 + i)] = d[r + n * i] - d[r + n * (K - 2)];
  for (size_t r = 0; r < n; r++) design[r + n * (K - 1)] = -d[r + n * (K - 2)];  // :26
}

// src/ncs_basis_cpp.cpp:30-59 + :82-99  ncs_basis_deriv.
void ace_oracle_ncs_basis_deriv(const double* x, int n_, const double* knots_in, int Kin, double* design) {
  const std::vector<double> knots = unique_sorted(knots_in, Kin);
  const size_t n = n_, K = knots.size();
  std::vector<double> d(n * K);
  auto sq = [](double v) { return std::pow(v, 2.0); };  // arma::pow(x - k, 2), as above
  for (size_t r = 0; r < n; r++)  // :44
    d[r + n * (K - 1)] = 3 * (x[r] > knots[K - 1] ? 1.0 : 0.0) * sq(x[r] - knots[K - 1]);
  for (size_t i = 0; i + 1 < K; i++) {  // :45-48
    for (size_t r = 0; r < n; r++) {
      double v = 3 * (x[r] > knots[i] ? 1.0 : 0.0) * sq(x[r] - knots[i]) - d[r + n * (K - 1)];
      d[r + n * i] = v / (knots[K - 1] - knots[i]);
    }
  }
  for (size_t r = 0; r < n; r++) d[r + n * (K - 1)] = 0.0;  // :51
  for (size_t r = 0; r < n; r++) design[r] = 1.0;           // :95
  for (size_t i = 0; i + 2 < K; i++)                        // :53-55
    for (size_t r = 0; r < n; r++) design[r + n * (1 + i)] = d[r + n * i] - d[r + n * (K - 2)];
  for (size_t r = 0; r < n; r++) design[r + n * (K - 1)] = -d[r + n * (K - 2)];  // :56
}

// ---------------------------------------------------------------------------
// R-level sequencing: Kernel$para_update + ace.train loop
// ---------------------------------------------------------------------------

// One Kernel$para_update (R/kernel_SE_R6.R:40-62, R/kernel_Matern32_R6.R:39-60) with
// Optim$update (R/optimizer_classes.R:22-31,54-63,85-92).
//   kernel: 0 = SE, 1 = Matern32.   optimizer: 0 = Nadam, 1 = Adam, 2 = Nesterov (GD/NAG).
//   use_chol: 0 = literal dsyevd inverse, 1 = CPU-favourable dpotrf/dpotri variant.
//   par[P], m[P], v[P] (v unused for Nesterov, m = nu) updated in place.
//   Kfull, Kcube, invK, eigenval: caller-provided work/state buffers (the R6 fields).
//   grad_out[P]: gradients after clipping (as R sees them).  stats[2].
//   tsec[4] (optional): seconds spent in build / inverse / gradient / rest.
// Returns 1 if gradients were finite, 0 otherwise (R would stop()).
int ace_oracle_para_update(int iter, const double* y, const double* X, const double* Z, int n, int p,
                           int Bz, int kernel, int optimizer, int use_chol, double lr, double beta1,
                           double beta2, double momentum, int norm_clip, double clip_at, double std_y,
                           double* par, double* m, double* v, double* Kfull, double* Kcube, double* invK,
                           double* eigenval, double* grad_out, double* stats, double* tsec);

}  // extern "C"

#include <chrono>
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

int ace_oracle_para_update(int iter, const double* y, const double* X, const double* Z, int n, int p,
                           int Bz, int kernel, int optimizer, int use_chol, double lr, double beta1,
                           double beta2, double momentum, int norm_clip, double clip_at, double std_y,
                           double* par, double* m, double* v, double* Kfull, double* Kcube, double* invK,
                           double* eigenval, double* grad_out, double* stats, double* tsec) {
  const int B = Bz + 1;
  const int P = 2 + B + B * p;
  double t0 = now_s();
  // getinv_kernel: kernel_mat_sym then invkernel_cpp (R/kernel_SE_R6.R:33-39)
  if (kernel == 0)
    ace_oracle_kernmat_SE_sym(X, Z, n, p, Bz, par, Kfull, Kcube);
  else
    ace_oracle_kernmat_Matern32_sym(X, Z, n, p, Bz, par, Kfull, Kcube);
  double t1 = now_s();
  if (use_chol)
    ace_oracle_invkernel_chol(Kfull, n, par[0], eigenval, invK);
  else
    ace_oracle_invkernel(Kfull, n, par[0], eigenval, invK);
  double t2 = now_s();
  if (iter == 1) par[1] = ace_oracle_mu_solution(y, invK, n);  // :45
  stats[0] = stats[1] = 0.0;
  if (kernel == 0)
    ace_oracle_grad_SE(y, X, Kfull, Kcube, invK, eigenval, par, n, p, B, std_y, stats, grad_out);
  else
    ace_oracle_grad_Matern(y, X, Kfull, Kcube, invK, eigenval, par, n, p, B, std_y, stats, grad_out);
  double t3 = now_s();
  // Optim$update: clip (in place on the gradient vector) then step
  ace_oracle_norm_clip(norm_clip, grad_out, P, clip_at);
  int ok;
  if (optimizer == 0)
    ok = ace_oracle_Nadam((double)iter, lr, beta1, beta2, 1e-8, m, v, grad_out, par, P);
  else if (optimizer == 1)
    ok = ace_oracle_Adam((double)iter, lr, beta1, beta2, 1e-8, m, v, grad_out, par, P);
  else
    ok = ace_oracle_Nesterov(lr, momentum, m, grad_out, par, P);
  // mean_solution with the PRE-update inverse (:54, quirk Q6)
  par[1] = ace_oracle_mu_solution(y, invK, n);
  double t4 = now_s();
  if (tsec) {
    tsec[0] = t1 - t0;
    tsec[1] = t2 - t1;
    tsec[2] = t3 - t2;
    tsec[3] = t4 - t3;
  }
  return ok;
}

// ace.train loop (R/main_ace.R:213-235): stats_out is 2 x (maxiter + 2) col-major, column 0 = 0.
// Returns the number of iterations run (iter), or -iter if gradients became non-finite at iter.
// After the loop, column iter+1 holds get_train_stats (R/kernel_SE_R6.R:63-74) computed with a
// LOCAL inverse that is not stored: invK keeps the inverse of the last iteration (quirk Q6).
int ace_oracle_train(const double* y, const double* X, const double* Z, int n, int p, int Bz, int kernel,
                     int optimizer, int use_chol, int maxiter, double tol, double lr, double beta1,
                     double beta2, double momentum, int norm_clip, double clip_at, double std_y,
                     double* par, double* m, double* v, double* invK, double* stats_out) {
  const size_t N = n, B = (size_t)Bz + 1, P = 2 + B + B * (size_t)p;
  std::vector<double> Kfull(N * N), Kcube(N * N * B), eig(N), grad(P);
  std::fill(stats_out, stats_out + 2 * ((size_t)maxiter + 2), 0.0);
  int iter = 1;
  for (iter = 1; iter <= maxiter; iter++) {
    double st[2];
    int ok = ace_oracle_para_update(iter, y, X, Z, n, p, Bz, kernel, optimizer, use_chol, lr, beta1, beta2,
                                    momentum, norm_clip, clip_at, std_y, par, m, v, Kfull.data(),
                                    Kcube.data(), invK, eig.data(), grad.data(), st, nullptr);
    if (!ok) return -iter;
    stats_out[0 + 2 * iter] = st[0];
    stats_out[1 + 2 * iter] = st[1];
    const double change = std::abs(stats_out[1 + 2 * iter] - stats_out[1 + 2 * (iter - 1)]);
    if ((change < tol) && (iter > 3)) break;
  }
  if (iter > maxiter) iter = maxiter;  // R's for leaves iter == maxiter when exhausted
  // get_train_stats: rebuild K with the final parameters, local inverse
  {
    std::vector<double> inv2(N * N);
    if (kernel == 0)
      ace_oracle_kernmat_SE_sym(X, Z, n, p, Bz, par, Kfull.data(), Kcube.data());
    else
      ace_oracle_kernmat_Matern32_sym(X, Z, n, p, Bz, par, Kfull.data(), Kcube.data());
    if (use_chol)
      ace_oracle_invkernel_chol(Kfull.data(), n, par[0], eig.data(), inv2.data());
    else
      ace_oracle_invkernel(Kfull.data(), n, par[0], eig.data(), inv2.data());
    double st[2];
    ace_oracle_stats(y, Kfull.data(), inv2.data(), eig.data(), par[1], std_y, n, st);
    stats_out[0 + 2 * (iter + 1)] = st[0];
    stats_out[1 + 2 * (iter + 1)] = st[1];
  }
  return iter;
}

}  // extern "C"
