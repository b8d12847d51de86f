// Harness for integration/src/ace_b200_shim.cpp: the Rcpp shim is compiled against the stand-in headers of
// oracle/miniarma (R / Rcpp / Armadillo are not installed in the build container), linked with libace_b200.so, and
// called like R would call the exported functions.  argv[1] = "gpu" additionally runs the device entry points.
#include <RcppArmadillo.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include "ace_b200.h"

namespace miniarma {
Blas& blas() {
  static Blas b;
  return b;
}
}  // namespace miniarma

bool Nadam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, arma::vec& m, arma::vec& v,
               const arma::vec& grad, arma::vec& para);
void norm_clip_cpp(bool flag, arma::vec& grads, double max_length);
arma::mat ncs_basis(arma::colvec x, arma::vec knots);
arma::mat normalize_train(arma::vec& y, arma::mat& X, arma::mat& Z);
Rcpp::List kernmat_SE_symmetric_cpp(const arma::mat& X, const arma::mat& Z, const arma::vec& parameters);
Rcpp::List invkernel_cpp(arma::mat pdmat, const double& sigma);
arma::vec grad_SE_cpp(const arma::vec& y, const arma::mat& X, const arma::mat& Z, const arma::mat& Kfull,
                      const arma::cube& K, const arma::mat& invKmatn, const arma::vec& eigenval,
                      const arma::vec& parameters, arma::vec& stats, const unsigned int& B, double std_y);
SEXP ace_fit_create_R(const arma::vec& y, const arma::mat& X, const arma::mat& Z, const arma::vec& parameters, int kernel,
                      int optimizer, double learning_rate, double beta1, double beta2, double momentum, bool norm_clip,
                      double clip_at, double std_y, int device);
arma::vec ace_fit_para_update_R(SEXP handle, int iter);
arma::vec ace_fit_get_parameters_R(SEXP handle);
Rcpp::List ace_fit_predict_R(SEXP handle, const arma::mat& X2, const arma::mat& Z2, double mean_y, double std_y);

#define REQUIRE(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
  // ---- host-side routines (no GPU needed)
  arma::vec m(3), v(3), g(3), par(3);
  m.zeros(); v.zeros();
  g[0] = 0.5; g[1] = -0.25; g[2] = 0.0;
  par[0] = 1; par[1] = 2; par[2] = 3;
  REQUIRE(Nadam_cpp(1.0, 0.01, 0.9, 0.999, 1e-8, m, v, g, par));
  REQUIRE(par[0] > 1.0 && par[1] < 2.0 && par[2] == 3.0);
  REQUIRE(std::fabs(m[0] - 0.05) < 1e-15);  // in place on the caller's vector, like the reference's arma::vec&
  arma::vec gc(2);
  gc[0] = 3.0; gc[1] = 4.0;
  norm_clip_cpp(true, gc, 1.0);
  REQUIRE(std::fabs(std::sqrt(gc[0] * gc[0] + gc[1] * gc[1]) - 1.0) < 1e-15);
  arma::vec x(5), kn(3);
  for (int i = 0; i < 5; ++i) x[i] = -0.8 + 0.4 * i;
  kn[0] = -1; kn[1] = 0.1; kn[2] = 1;
  arma::mat Bm = ncs_basis(x, kn);
  REQUIRE(Bm.n_rows == 5 && Bm.n_cols == 3 && Bm(2, 0) == x[2]);
  arma::vec yy(6);
  arma::mat XX(6, 1), ZZ(6, 1);
  for (int i = 0; i < 6; ++i) { yy[i] = i * i; XX(i, 0) = std::sin(i + 1.0); ZZ(i, 0) = i % 2; }
  arma::mat mo = normalize_train(yy, XX, ZZ);
  REQUIRE(mo.n_rows == 3 && mo.n_cols == 3 && std::fabs(mo(0, 0) - 55.0 / 6.0) < 1e-14 && mo(2, 2) == 1.0);
  std::printf("OK host devices=%d\n", ace_device_count());
  if (argc < 2 || std::strcmp(argv[1], "gpu") != 0) return 0;

  // ---- device entry points: per-function exports and the fit handle, as the R6 class calls them
  const int n = 60, p = 2, Bz = 1, B = 2, P = 2 + B + B * p;
  arma::mat X(n, p), Z(n, Bz);
  arma::vec y(n), th(P);
  for (int i = 0; i < n; ++i) {
    X(i, 0) = std::sin(0.37 * i); X(i, 1) = std::cos(0.91 * i); Z(i, 0) = (i % 3 == 0) ? 0.0 : std::cos(0.11 * i);
    y[i] = std::sin(0.5 * i) + 0.1 * std::cos(3.0 * i);
  }
  th[0] = std::log(0.3); th[1] = 0.0; th[2] = 0.0; th[3] = 0.0;
  for (int i = 4; i < P; ++i) th[i] = std::log(2.0);
  Rcpp::List kl = kernmat_SE_symmetric_cpp(X, Z, th);
  const arma::mat& K = kl["full"].m;
  REQUIRE(K.n_rows == (arma::uword)n && K(3, 7) == K(7, 3) && kl["elements"].c.n_slices == 2);
  REQUIRE(kl["elements"].c(0, 5, 1) == 0.0);  // exact zero where z == 0
  Rcpp::List il = invkernel_cpp(K, th[0]);
  const arma::mat& inv = il["inv"].m;
  double worst = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += inv(i, k) * (K(k, j) + (k == j ? std::exp(th[0]) : 0.0));
      worst = std::fmax(worst, std::fabs(s - (i == j ? 1.0 : 0.0)));
    }
  REQUIRE(worst < 1e-9);
  arma::vec stats(2);
  stats.zeros();
  arma::vec gr = grad_SE_cpp(y, X, Z, K, kl["elements"].c, inv, il["eigenval"].m, th, stats, B, 1.0);
  REQUIRE(gr.n_elem == (arma::uword)P && std::isfinite(stats[1]) && stats[1] != 0.0);
  SEXP h = ace_fit_create_R(y, X, Z, th, ACE_KERNEL_SE, ACE_OPT_NADAM, 0.01, 0.9, 0.999, 0.0, true, 1.0, 1.0, 0);
  arma::vec out = ace_fit_para_update_R(h, 2);  // iter != 1: mu is not replaced first, so the evidence matches grad_SE_cpp's
  REQUIRE(std::fabs(out[1] - stats[1]) <= 1e-9 * std::fabs(stats[1]));
  arma::vec th2 = ace_fit_get_parameters_R(h);
  REQUIRE(th2.n_elem == (arma::uword)P && th2[0] != th[0]);
  arma::mat X2(4, p), Z2(4, Bz);
  for (int i = 0; i < 4; ++i) { X2(i, 0) = 0.1 * i; X2(i, 1) = -0.2 * i; Z2(i, 0) = 0.3; }
  Rcpp::List pl = ace_fit_predict_R(h, X2, Z2, 0.0, 1.0);
  REQUIRE(pl["map"].m.n_elem == 4 && pl["var"].m[0] > 0.0);
  ace_fit_destroy(static_cast<ace_fit*>(h));
  std::printf("OK gpu evidence %.12f\n", out[1]);
  return 0;
}
