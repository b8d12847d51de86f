// ace_b200_shim.cpp -- drop-in replacement for the bodies of the reference's native sources
// (src/kernel_SE_cpp.cpp, kernel_Matern_cpp.cpp, stats_cpp.cpp, pred_cpp.cpp, optimizer_cpp.cpp, utilities_cpp.cpp,
// ncs_basis_cpp.cpp): the same 19 exported functions with the same signatures (src/RcppExports.cpp:10-299), each a
// marshalling call into the C ABI of libace_b200.so (include/ace_b200.h), plus the device-resident fit handle the
// R6 kernel classes use (integration/R/kernel_*_R6.R).  Remove the seven reference .cpp files, add this one, run
// Rcpp::compileAttributes(): RcppExports.cpp / RcppExports.R regenerate with the 19 original entries + the handle's.
//
// Armadillo objects are dense and column-major, so .memptr() is exactly what the ABI takes; `const arma::mat&`
// arguments alias R memory (no copy), results are fresh allocations, and the in-place arguments of the reference
// (stats, m, v, nu, para, grads, y / X / Z of normalize_*) are written in place here too.
//
// This file is compiled in the build container only against the stand-in headers of oracle/miniarma (R, Rcpp and
// Armadillo are not installed there): tests/test_integration_shim.py builds it that way and calls through it.
// [[Rcpp::depends("RcppArmadillo")]]
#include <RcppArmadillo.h>

#include "ace_b200.h"

using namespace Rcpp;

static const char* kNotFinite =
    "Some gradients are not finite, NaN, or NA. Often this is due to too large learning rates.";

static void check(int status, const char* where) {
  if (status == 0) return;
  if (status == ACE_ERR_NOT_FINITE) Rcpp::stop(kNotFinite);  // R/optimizer_classes.R:26-29
  Rcpp::stop("%s: status %d: %s", where, status, ace_last_error());
}

// ------------------------------------------------------------------------------------------- kernel builds
typedef int (*kern_rect_fn)(const double*, const double*, const double*, const double*, int, int, int, int,
                            const double*, double*, double*);
typedef int (*kern_sym_fn)(const double*, const double*, int, int, int, const double*, double*, double*);

static Rcpp::List kern_rect(kern_rect_fn fn, const char* name, const arma::mat& X1, const arma::mat& X2,
                            const arma::mat& Z1, const arma::mat& Z2, const arma::vec& parameters) {
  const int n1 = X1.n_rows, n2 = X2.n_rows, p = X2.n_cols, Bz = Z1.n_cols;
  arma::mat Kfull(n1, n2);
  arma::cube K(n1, n2, Bz + 1);
  check(fn(X1.memptr(), X2.memptr(), Z1.memptr(), Z2.memptr(), n1, n2, p, Bz, parameters.memptr(), Kfull.memptr(),
           K.memptr()), name);
  return Rcpp::List::create(Rcpp::Named("full") = Kfull, Rcpp::Named("elements") = K);
}

static Rcpp::List kern_sym(kern_sym_fn fn, const char* name, const arma::mat& X, const arma::mat& Z,
                           const arma::vec& parameters) {
  const int n = X.n_rows, p = X.n_cols, Bz = Z.n_cols;
  arma::mat Kfull(n, n);
  arma::cube Ks(n, n, Bz + 1);
  check(fn(X.memptr(), Z.memptr(), n, p, Bz, parameters.memptr(), Kfull.memptr(), Ks.memptr()), name);
  return Rcpp::List::create(Rcpp::Named("full") = Kfull, Rcpp::Named("elements") = Ks);
}

// [[Rcpp::export]]
Rcpp::List kernmat_SE_cpp(const arma::mat& X1, const arma::mat& X2, const arma::mat& Z1, const arma::mat& Z2,
                          const arma::vec& parameters) {
  return kern_rect(ace_kernmat_SE_cpp, "kernmat_SE_cpp", X1, X2, Z1, Z2, parameters);
}

// [[Rcpp::export]]
Rcpp::List kernmat_SE_symmetric_cpp(const arma::mat& X, const arma::mat& Z, const arma::vec& parameters) {
  return kern_sym(ace_kernmat_SE_symmetric_cpp, "kernmat_SE_symmetric_cpp", X, Z, parameters);
}

// [[Rcpp::export]]
Rcpp::List kernmat_Matern32_cpp(const arma::mat& X1, const arma::mat& X2, const arma::mat& Z1, const arma::mat& Z2,
                                const arma::vec& parameters) {
  return kern_rect(ace_kernmat_Matern32_cpp, "kernmat_Matern32_cpp", X1, X2, Z1, Z2, parameters);
}

// [[Rcpp::export]]
Rcpp::List kernmat_Matern32_symmetric_cpp(const arma::mat& X, const arma::mat& Z, const arma::vec& parameters) {
  return kern_sym(ace_kernmat_Matern32_symmetric_cpp, "kernmat_Matern32_symmetric_cpp", X, Z, parameters);
}

// ------------------------------------------------------------------------------------------- inverse
// [[Rcpp::export]]
Rcpp::List invkernel_cpp(arma::mat pdmat, const double& sigma) {
  const int n = pdmat.n_cols;
  arma::vec eigval(n);
  arma::mat inv(n, n);
  const int s = ace_invkernel_cpp(pdmat.memptr(), n, sigma, eigval.memptr(), inv.memptr());
  if (s > 0 && s != ACE_ERR_NOT_FINITE) {
    // not positive definite (s = 1-based pivot): the reference prints and carries on with NaNs, which surface as
    // "gradients are not finite" one call later (src/kernel_SE_cpp.cpp:144-152, quirk Q10)
    Rcout << "Eigenvalue decomp. not completed." << std::endl;
    inv.fill(arma::datum::nan);
    eigval.fill(arma::datum::nan);
  } else {
    check(s, "invkernel_cpp");
  }
  // eigenval = diag(L)^2 of the Cholesky factor: same sum(log(.)), the only use the package makes of it
  return Rcpp::List::create(Rcpp::Named("eigenval") = eigval, Rcpp::Named("inv") = inv);
}

// ------------------------------------------------------------------------------------------- gradients, statistics
// [[Rcpp::export]]
arma::vec grad_SE_cpp(const arma::vec& y, const arma::mat& X, const arma::mat& Z, const arma::mat& Kfull,
                      const arma::cube& K, const arma::mat& invKmatn, const arma::vec& eigenval,
                      const arma::vec& parameters, arma::vec& stats, const unsigned int& B, double std_y) {
  arma::vec g(parameters.n_elem);
  check(ace_grad_SE_cpp(y.memptr(), X.memptr(), Z.memptr(), Kfull.memptr(), K.memptr(), invKmatn.memptr(),
                        eigenval.memptr(), parameters.memptr(), stats.memptr(), B, std_y, (int)X.n_rows,
                        (int)X.n_cols, g.memptr()), "grad_SE_cpp");
  return g;  // stats was written in place, like the reference's arma::vec&
}

// [[Rcpp::export]]
arma::vec grad_Matern_cpp(const arma::vec& y, const arma::mat& X, const arma::mat& Z, arma::mat& Kfull, arma::cube& K,
                          arma::mat& invKmatn, arma::vec& eigenval, const arma::vec& parameters, arma::vec& stats,
                          const unsigned int& B, double std_y) {
  arma::vec g(parameters.n_elem);
  check(ace_grad_Matern_cpp(y.memptr(), X.memptr(), Z.memptr(), Kfull.memptr(), K.memptr(), invKmatn.memptr(),
                            eigenval.memptr(), parameters.memptr(), stats.memptr(), B, std_y, (int)X.n_rows,
                            (int)X.n_cols, g.memptr()), "grad_Matern_cpp");
  return g;
}

// [[Rcpp::export]]
arma::rowvec stats_cpp(const arma::colvec& y, const arma::mat& Kmat, const arma::mat& invKmatn,
                       const arma::vec& eigenval, const double mu, double std_y = 1) {
  arma::rowvec out(2);
  check(ace_stats_cpp(y.memptr(), Kmat.memptr(), invKmatn.memptr(), eigenval.memptr(), mu, std_y, (int)y.n_rows,
                      out.memptr()), "stats_cpp");
  return out;
}

// [[Rcpp::export]]
double mu_solution_cpp(arma::colvec& y, arma::mat& invKmat) {
  double mu = 0.0;
  check(ace_mu_solution_cpp(y.memptr(), invKmat.memptr(), (int)y.n_rows, &mu), "mu_solution_cpp");
  return mu;
}

// ------------------------------------------------------------------------------------------- posterior
// [[Rcpp::export]]
Rcpp::List pred_cpp(const arma::vec& y_X, const double sigma, const double mu, const arma::mat& invK_XX,
                    arma::mat& K_xX, arma::mat K_xx, double mean_y, double std_y) {
  const int nx = K_xx.n_rows, nX = invK_XX.n_rows;
  arma::vec map(nx), var(nx);
  arma::mat ci(nx, 2);
  check(ace_pred_cpp(y_X.memptr(), sigma, mu, invK_XX.memptr(), K_xX.memptr(), K_xx.memptr(), mean_y, std_y, nx, nX,
                     map.memptr(), ci.memptr(), var.memptr()), "pred_cpp");
  return Rcpp::List::create(Rcpp::Named("map") = map, Rcpp::Named("ci") = ci, Rcpp::Named("var") = var);
}

static Rcpp::List average_entry(const double* a4) {  // {map, ci lo, ci hi, var} -> list(map, ci, var)
  arma::vec ci(2);
  ci[0] = a4[1];
  ci[1] = a4[2];
  return Rcpp::List::create(Rcpp::Named("map") = a4[0], Rcpp::Named("ci") = ci, Rcpp::Named("var") = a4[3]);
}

// [[Rcpp::export]]
Rcpp::List pred_marginal_cpp(const arma::vec& y_X, const arma::colvec& Z_x, const double sigma, const double mu,
                             const arma::mat& invK_XX, const arma::cube& K_xX, const arma::cube& K_xx,
                             const double& mean_y, const double& std_y, const double& std_Z, bool calculate_ate) {
  const int nx = K_xx.n_rows, nX = invK_XX.n_rows, B = K_xx.n_slices;
  arma::vec map(nx), var(nx);
  arma::mat ci(nx, 2);
  double avg[12] = {0};
  check(ace_pred_marginal_cpp(y_X.memptr(), Z_x.memptr(), sigma, mu, invK_XX.memptr(), K_xX.memptr(), K_xx.memptr(),
                              mean_y, std_y, std_Z, calculate_ate ? 1 : 0, nx, nX, B, map.memptr(), ci.memptr(),
                              var.memptr(), avg), "pred_marginal_cpp");
  if (!calculate_ate)
    return Rcpp::List::create(Rcpp::Named("map") = map, Rcpp::Named("ci") = ci, Rcpp::Named("var") = var);
  return Rcpp::List::create(Rcpp::Named("map") = map, Rcpp::Named("ci") = ci, Rcpp::Named("var") = var,
                            Rcpp::Named("ate") = average_entry(avg), Rcpp::Named("att") = average_entry(avg + 4),
                            Rcpp::Named("atu") = average_entry(avg + 8));  // src/pred_cpp.cpp:112-123
}

// ------------------------------------------------------------------------------------------- optimisers, clip
// [[Rcpp::export]]
bool Nesterov_cpp(double learn_rate, double momentum, arma::vec& nu, const arma::vec& grad, arma::vec& para) {
  return ace_Nesterov_cpp(learn_rate, momentum, nu.memptr(), grad.memptr(), para.memptr(), (int)para.n_elem) != 0;
}

// [[Rcpp::export]]
bool Nadam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, arma::vec& m, arma::vec& v,
               const arma::vec& grad, arma::vec& para) {
  return ace_Nadam_cpp(iter, learn_rate, beta1, beta2, eps, m.memptr(), v.memptr(), grad.memptr(), para.memptr(),
                       (int)para.n_elem) != 0;
}

// [[Rcpp::export]]
bool Adam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, arma::vec& m, arma::vec& v,
              const arma::vec& grad, arma::vec& para) {
  return ace_Adam_cpp(iter, learn_rate, beta1, beta2, eps, m.memptr(), v.memptr(), grad.memptr(), para.memptr(),
                      (int)para.n_elem) != 0;
}

// [[Rcpp::export]]
void norm_clip_cpp(bool flag, arma::vec& grads, double max_length) {
  ace_norm_clip_cpp(flag ? 1 : 0, grads.memptr(), (int)grads.n_elem, max_length);
}

// ------------------------------------------------------------------------------------------- bases, normalisation
// [[Rcpp::export]]
arma::mat ncs_basis(arma::colvec x, arma::vec knots) {
  const int K = ace_ncs_basis(x.memptr(), (int)x.n_elem, knots.memptr(), (int)knots.n_elem, NULL);
  if (K < 0) check(K, "ncs_basis");
  arma::mat design(x.n_elem, K);
  const int s = ace_ncs_basis(x.memptr(), (int)x.n_elem, knots.memptr(), (int)knots.n_elem, design.memptr());
  if (s < 0) check(s, "ncs_basis");
  return design;
}

// [[Rcpp::export]]
arma::mat ncs_basis_deriv(arma::colvec x, arma::vec knots) {
  const int K = ace_ncs_basis_deriv(x.memptr(), (int)x.n_elem, knots.memptr(), (int)knots.n_elem, NULL);
  if (K < 0) check(K, "ncs_basis_deriv");
  arma::mat design(x.n_elem, K);
  const int s = ace_ncs_basis_deriv(x.memptr(), (int)x.n_elem, knots.memptr(), (int)knots.n_elem, design.memptr());
  if (s < 0) check(s, "ncs_basis_deriv");
  return design;
}

// [[Rcpp::export]]
arma::mat normalize_train(arma::vec& y, arma::mat& X, arma::mat& Z) {
  arma::mat moments(1 + X.n_cols + Z.n_cols, 3);
  check(ace_normalize_train(y.memptr(), X.memptr(), Z.memptr(), (int)X.n_rows, (int)X.n_cols, (int)Z.n_cols,
                            moments.memptr()), "normalize_train");
  return moments;
}

// [[Rcpp::export]]
void normalize_test(arma::mat& X, arma::mat& Z, const arma::mat& moments) {
  check(ace_normalize_test(X.memptr(), Z.memptr(), (int)X.n_rows, (int)X.n_cols, (int)Z.n_cols, moments.memptr()),
        "normalize_test");
}

// ------------------------------------------------------------------------------------------- device-resident handle
// One external pointer per R6 kernel object: K, K^-1 and the optimiser moments stay in HBM between calls.
static void fit_finalizer(ace_fit* h) { ace_fit_destroy(h); }
typedef Rcpp::XPtr<ace_fit, Rcpp::PreserveStorage, fit_finalizer, true> FitPtr;

// [[Rcpp::export]]
SEXP ace_fit_create_R(const arma::vec& y, const arma::mat& X, const arma::mat& Z, const arma::vec& parameters,
                      int kernel, int optimizer, double learning_rate, double beta1, double beta2, double momentum,
                      bool norm_clip, double clip_at, double std_y, int device) {
  ace_fit_config cfg;
  ace_fit_default_config(&cfg);
  cfg.kernel = kernel; cfg.optimizer = optimizer; cfg.learning_rate = learning_rate; cfg.beta1 = beta1;
  cfg.beta2 = beta2; cfg.momentum = momentum; cfg.norm_clip = norm_clip ? 1 : 0; cfg.clip_at = clip_at;
  cfg.std_y = std_y; cfg.device = device;
  ace_fit* h = NULL;
  check(ace_fit_create(&h, y.memptr(), X.memptr(), Z.memptr(), (int)X.n_rows, (int)X.n_cols, (int)Z.n_cols,
                       parameters.memptr(), &cfg), "ace_fit_create");
  return FitPtr(h, true);
}

// [[Rcpp::export]]
arma::vec ace_fit_para_update_R(SEXP handle, int iter) {
  FitPtr h(handle);
  arma::vec out(3);  // RMSE, log-evidence, gradient norm (after clipping)
  const int s = ace_fit_para_update(h.get(), iter, out.memptr(), out.memptr() + 2);
  if (s > 0) Rcpp::stop(kNotFinite);  // non-finite gradients, or a non-positive pivot (quirk Q10: same message)
  check(s, "ace_fit_para_update");
  Rcpp::checkUserInterrupt();  // between iterations, on R's main thread
  return out;
}

// [[Rcpp::export]]
arma::vec ace_fit_get_parameters_R(SEXP handle) {
  FitPtr h(handle);
  int n, p, B, P;
  check(ace_fit_dims(h.get(), &n, &p, &B, &P), "ace_fit_dims");
  arma::vec par(P);
  check(ace_fit_get_parameters(h.get(), par.memptr()), "ace_fit_get_parameters");
  return par;
}

// [[Rcpp::export]]
void ace_fit_set_parameters_R(SEXP handle, const arma::vec& parameters) {
  FitPtr h(handle);
  check(ace_fit_set_parameters(h.get(), parameters.memptr()), "ace_fit_set_parameters");
}

// [[Rcpp::export]]
arma::mat ace_fit_get_invKmatn_R(SEXP handle) {
  FitPtr h(handle);
  int n, p, B, P;
  check(ace_fit_dims(h.get(), &n, &p, &B, &P), "ace_fit_dims");
  arma::mat inv(n, n);
  check(ace_fit_get_invKmatn(h.get(), inv.memptr()), "ace_fit_get_invKmatn");
  return inv;
}

// [[Rcpp::export]]
arma::rowvec ace_fit_get_train_stats_R(SEXP handle) {
  FitPtr h(handle);
  arma::rowvec out(2);
  check(ace_fit_get_train_stats(h.get(), out.memptr()), "ace_fit_get_train_stats");
  return out;
}

// [[Rcpp::export]]
Rcpp::List ace_fit_predict_R(SEXP handle, const arma::mat& X2, const arma::mat& Z2, double mean_y, double std_y) {
  FitPtr h(handle);
  const int nx = X2.n_rows;
  arma::vec map(nx), var(nx);
  arma::mat ci(nx, 2);
  check(ace_fit_predict(h.get(), X2.memptr(), Z2.memptr(), nx, mean_y, std_y, map.memptr(), ci.memptr(), var.memptr()),
        "ace_fit_predict");
  return Rcpp::List::create(Rcpp::Named("map") = map, Rcpp::Named("ci") = ci, Rcpp::Named("var") = var);
}

// [[Rcpp::export]]
Rcpp::List ace_fit_predict_marginal_R(SEXP handle, const arma::mat& X2, const arma::mat& Z2, const arma::mat& dZ2,
                                      double mean_y, double std_y, double std_Z, bool calculate_ate) {
  FitPtr h(handle);
  const int nx = X2.n_rows;
  arma::vec map(nx), var(nx);
  arma::mat ci(nx, 2);
  double avg[12] = {0};
  check(ace_fit_predict_marginal(h.get(), X2.memptr(), Z2.memptr(), dZ2.memptr(), nx, mean_y, std_y, std_Z,
                                 calculate_ate ? 1 : 0, map.memptr(), ci.memptr(), var.memptr(), avg),
        "ace_fit_predict_marginal");
  if (!calculate_ate)
    return Rcpp::List::create(Rcpp::Named("map") = map, Rcpp::Named("ci") = ci, Rcpp::Named("var") = var);
  return Rcpp::List::create(Rcpp::Named("map") = map, Rcpp::Named("ci") = ci, Rcpp::Named("var") = var,
                            Rcpp::Named("ate") = average_entry(avg), Rcpp::Named("att") = average_entry(avg + 4),
                            Rcpp::Named("atu") = average_entry(avg + 8));
}
