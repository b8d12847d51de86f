// ref_harness.cpp -- C entry points around the REFERENCE'S OWN functions, test infrastructure only.
//
// oracle/Makefile (target _ref) compiles the reference's native sources unmodified, where they lie under
// /root/reference/src, against the stand-in headers in oracle/miniarma/ and links them with this file into
// oracle/_ref/libace_ref.so.  The exported names and signatures are those of the oracle restatement
// (oracle/ace_oracle.cpp), so oracle/__init__.py can drive either library with the same ctypes code:
// `oracle.using_reference()` swaps the library.  Nothing of the product links or loads this.
#include <RcppArmadillo.h>

#include <dlfcn.h>

#include <cstdio>

// ---- declarations of the reference's exported functions (src/RcppExports.cpp:10-299 declares the same) ----
Rcpp::List kernmat_SE_cpp(const arma::mat& X1, const arma::mat& X2, const arma::mat& Z1, const arma::mat& Z2,
                          const arma::vec& parameters);
Rcpp::List kernmat_SE_symmetric_cpp(const arma::mat& X, const arma::mat& Z, const arma::vec& parameters);
Rcpp::List kernmat_Matern32_cpp(const arma::mat& X1, const arma::mat& X2, const arma::mat& Z1, const arma::mat& Z2,
                                const arma::vec& parameters);
Rcpp::List kernmat_Matern32_symmetric_cpp(const arma::mat& X, const arma::mat& Z, const arma::vec& parameters);
Rcpp::List invkernel_cpp(arma::mat pdmat, const double& sigma);
arma::vec grad_SE_cpp(const arma::vec& y, const arma::mat& X, const arma::mat& Z, const arma::mat& Kfull,
                      const arma::cube& K, const arma::mat& invKmatn, const arma::vec& eigenval,
                      const arma::vec& parameters, arma::vec& stats, const unsigned int& B, double std_y);
arma::vec grad_Matern_cpp(const arma::vec& y, const arma::mat& X, const arma::mat& Z, arma::mat& Kfull, arma::cube& K,
                          arma::mat& invKmatn, arma::vec& eigenval, const arma::vec& parameters, arma::vec& stats,
                          const unsigned int& B, double std_y);
arma::rowvec stats_cpp(const arma::colvec& y, const arma::mat& Kmat, const arma::mat& invKmatn,
                       const arma::vec& eigenval, const double mu, double std_y);
double mu_solution_cpp(arma::colvec& y, arma::mat& invKmat);
arma::mat normalize_train(arma::vec& y, arma::mat& X, arma::mat& Z);
void normalize_test(arma::mat& X, arma::mat& Z, const arma::mat& moments);
void norm_clip_cpp(bool flag, arma::vec& grads, double max_length);
bool Nesterov_cpp(double learn_rate, double momentum, arma::vec& nu, const arma::vec& grad, arma::vec& para);
bool Nadam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, arma::vec& m, arma::vec& v,
               const arma::vec& grad, arma::vec& para);
bool Adam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, arma::vec& m, arma::vec& v,
              const arma::vec& grad, arma::vec& para);
Rcpp::List pred_cpp(const arma::vec& y_X, const double sigma, const double mu, const arma::mat& invK_XX,
                    arma::mat& K_xX, arma::mat K_xx, double mean_y, double std_y);
Rcpp::List pred_marginal_cpp(const arma::vec& y_X, const arma::colvec& Z_x, const double sigma, const double mu,
                             const arma::mat& invK_XX, const arma::cube& K_xX, const arma::cube& K_xx,
                             const double& mean_y, const double& std_y, const double& std_Z, bool calculate_ate);
arma::mat ncs_basis(arma::colvec x, arma::vec knots);
arma::mat ncs_basis_deriv(arma::colvec x, arma::vec knots);

namespace miniarma {
Blas& blas() {
  static Blas b;
  return b;
}
}  // namespace miniarma

namespace {
typedef void (*setthr_fn)(int);
typedef int (*getthr_fn)(void);
setthr_fn p_setthr = nullptr;
getthr_fn p_getthr = nullptr;

arma::mat in_mat(const double* p, size_t r, size_t c) {
  arma::mat m(r, c);
  if (r * c > 0) std::memcpy(m.mem, p, sizeof(double) * r * c);
  return m;
}
arma::vec in_vec(const double* p, size_t n) {
  arma::vec v(n);
  if (n) std::memcpy(v.mem, p, sizeof(double) * n);
  return v;
}
arma::cube in_cube(const double* p, size_t r, size_t c, size_t s) {
  arma::cube q(r, c, s);
  if (r * c * s > 0) std::memcpy(q.mem, p, sizeof(double) * r * c * s);
  return q;
}
void out(double* dst, const arma::Mat& m) {
  if (m.n_elem) std::memcpy(dst, m.mem, sizeof(double) * m.n_elem);
}
void out_kern(const Rcpp::List& l, double* full, double* elements) {
  out(full, l["full"].m);
  const arma::cube& c = l["elements"].c;
  if (c.n_elem) std::memcpy(elements, c.mem, sizeof(double) * c.n_elem);
}
}  // namespace

extern "C" {

int ace_oracle_init(const char* openblas_path, int nthreads) {
  void* h = dlopen(openblas_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    std::fprintf(stderr, "ace_ref init: dlopen failed: %s\n", dlerror());
    return -1;
  }
  miniarma::Blas& b = miniarma::blas();
  b.dgemm = (miniarma::dgemm_fn)dlsym(h, "scipy_dgemm_");
  b.dgemv = (miniarma::dgemv_fn)dlsym(h, "scipy_dgemv_");
  b.dsyrk = (miniarma::dsyrk_fn)dlsym(h, "scipy_dsyrk_");
  b.dsyevd = (miniarma::dsyevd_fn)dlsym(h, "scipy_dsyevd_");
  p_setthr = (setthr_fn)dlsym(h, "scipy_openblas_set_num_threads");
  p_getthr = (getthr_fn)dlsym(h, "scipy_openblas_get_num_threads");
  if (!b.dgemm || !b.dgemv || !b.dsyrk || !b.dsyevd) return -2;
  if (nthreads > 0 && p_setthr) p_setthr(nthreads);
  return 0;
}
int ace_oracle_threads(void) { return p_getthr ? p_getthr() : 1; }
int ace_oracle_is_reference(void) { return 1; }

void ace_oracle_kernmat_SE(const double* X1, const double* X2, const double* Z1, const double* Z2, int n1, int n2, int p,
                           int Bz, const double* par, double* full, double* elements) {
  const size_t B = Bz + 1, P = 2 + B + B * p;
  out_kern(kernmat_SE_cpp(in_mat(X1, n1, p), in_mat(X2, n2, p), in_mat(Z1, n1, Bz), in_mat(Z2, n2, Bz), in_vec(par, P)),
           full, elements);
}
void ace_oracle_kernmat_Matern32(const double* X1, const double* X2, const double* Z1, const double* Z2, int n1, int n2,
                                 int p, int Bz, const double* par, double* full, double* elements) {
  const size_t B = Bz + 1, P = 2 + B + B * p;
  out_kern(kernmat_Matern32_cpp(in_mat(X1, n1, p), in_mat(X2, n2, p), in_mat(Z1, n1, Bz), in_mat(Z2, n2, Bz),
                                in_vec(par, P)),
           full, elements);
}
void ace_oracle_kernmat_SE_sym(const double* X, const double* Z, int n, int p, int Bz, const double* par, double* full,
                               double* elements) {
  const size_t B = Bz + 1, P = 2 + B + B * p;
  out_kern(kernmat_SE_symmetric_cpp(in_mat(X, n, p), in_mat(Z, n, Bz), in_vec(par, P)), full, elements);
}
void ace_oracle_kernmat_Matern32_sym(const double* X, const double* Z, int n, int p, int Bz, const double* par,
                                     double* full, double* elements) {
  const size_t B = Bz + 1, P = 2 + B + B * p;
  out_kern(kernmat_Matern32_symmetric_cpp(in_mat(X, n, p), in_mat(Z, n, Bz), in_vec(par, P)), full, elements);
}

int ace_oracle_invkernel(const double* K, int n, double sigma, double* eigenval, double* inv) {
  Rcpp::List l = invkernel_cpp(in_mat(K, n, n), sigma);
  out(eigenval, l["eigenval"].m);
  out(inv, l["inv"].m);
  return 0;
}

void ace_oracle_grad_SE(const double* y, const double* X, const double* Kfull, const double* Kcube, const double* invK,
                        const double* eigenval, const double* par, int n, int p, int B, double std_y, double* stats,
                        double* grad) {
  const size_t P = 2 + (size_t)B + (size_t)B * p;
  arma::vec st(2);
  st.zeros();
  const unsigned int Bu = (unsigned int)B;
  arma::mat Zdummy(n, B > 1 ? B - 1 : 1);
  arma::vec g = grad_SE_cpp(in_vec(y, n), in_mat(X, n, p), Zdummy, in_mat(Kfull, n, n), in_cube(Kcube, n, n, B),
                            in_mat(invK, n, n), in_vec(eigenval, n), in_vec(par, P), st, Bu, std_y);
  out(grad, g);
  out(stats, st);
}
void ace_oracle_grad_Matern(const double* y, const double* X, const double* Kfull, const double* Kcube,
                            const double* invK, const double* eigenval, const double* par, int n, int p, int B,
                            double std_y, double* stats, double* grad) {
  const size_t P = 2 + (size_t)B + (size_t)B * p;
  arma::vec st(2);
  st.zeros();
  const unsigned int Bu = (unsigned int)B;
  arma::mat Zdummy(n, B > 1 ? B - 1 : 1), Kf = in_mat(Kfull, n, n), iK = in_mat(invK, n, n);
  arma::cube Kc = in_cube(Kcube, n, n, B);
  arma::vec ev = in_vec(eigenval, n);
  arma::vec g = grad_Matern_cpp(in_vec(y, n), in_mat(X, n, p), Zdummy, Kf, Kc, iK, ev, in_vec(par, P), st, Bu, std_y);
  out(grad, g);
  out(stats, st);
}

void ace_oracle_stats(const double* y, const double* Kmat, const double* invK, const double* eigenval, double mu,
                      double std_y, int n, double* stats) {
  out(stats, stats_cpp(in_vec(y, n), in_mat(Kmat, n, n), in_mat(invK, n, n), in_vec(eigenval, n), mu, std_y));
}
double ace_oracle_mu_solution(const double* y, const double* invK, int n) {
  arma::vec yy = in_vec(y, n);
  arma::mat iK = in_mat(invK, n, n);
  return mu_solution_cpp(yy, iK);
}
void ace_oracle_norm_clip(int flag, double* grads, int P, double max_length) {
  arma::vec g = in_vec(grads, P);
  norm_clip_cpp(flag != 0, g, max_length);
  out(grads, g);
}
int ace_oracle_Nesterov(double lr, double momentum, double* nu, const double* grad, double* para, int P) {
  arma::vec n_ = in_vec(nu, P), p_ = in_vec(para, P);
  const bool ok = Nesterov_cpp(lr, momentum, n_, in_vec(grad, P), p_);
  out(nu, n_);
  out(para, p_);
  return ok ? 1 : 0;
}
int ace_oracle_Nadam(double iter, double lr, double beta1, double beta2, double eps, double* m, double* v,
                     const double* grad, double* para, int P) {
  arma::vec m_ = in_vec(m, P), v_ = in_vec(v, P), p_ = in_vec(para, P);
  const bool ok = Nadam_cpp(iter, lr, beta1, beta2, eps, m_, v_, in_vec(grad, P), p_);
  out(m, m_);
  out(v, v_);
  out(para, p_);
  return ok ? 1 : 0;
}
int ace_oracle_Adam(double iter, double lr, double beta1, double beta2, double eps, double* m, double* v,
                    const double* grad, double* para, int P) {
  arma::vec m_ = in_vec(m, P), v_ = in_vec(v, P), p_ = in_vec(para, P);
  const bool ok = Adam_cpp(iter, lr, beta1, beta2, eps, m_, v_, in_vec(grad, P), p_);
  out(m, m_);
  out(v, v_);
  out(para, p_);
  return ok ? 1 : 0;
}

void ace_oracle_pred(const double* y_X, double sigma, double mu, const double* invK, const double* K_xX,
                     const double* K_xx, double mean_y, double std_y, int nx, int nX, double* map, double* ci,
                     double* var) {
  arma::mat KxX = in_mat(K_xX, nx, nX);
  Rcpp::List l = pred_cpp(in_vec(y_X, nX), sigma, mu, in_mat(invK, nX, nX), KxX, in_mat(K_xx, nx, nx), mean_y, std_y);
  out(map, l["map"].m);
  out(ci, l["ci"].m);
  out(var, l["var"].m);
}
void ace_oracle_pred_marginal(const double* y_X, const double* Z_x, double sigma, double mu, const double* invK,
                              const double* K_xX, const double* K_xx, double mean_y, double std_y, double std_Z,
                              int calculate_ate, int nx, int nX, int B, double* map, double* ci, double* var,
                              double* avg) {
  Rcpp::List l = pred_marginal_cpp(in_vec(y_X, nX), in_vec(Z_x, nx), sigma, mu, in_mat(invK, nX, nX),
                                   in_cube(K_xX, nx, nX, B), in_cube(K_xx, nx, nx, B), mean_y, std_y, std_Z,
                                   calculate_ate != 0);
  out(map, l["map"].m);
  out(ci, l["ci"].m);
  out(var, l["var"].m);
  if (calculate_ate) {
    const char* names[3] = {"ate", "att", "atu"};
    for (int k = 0; k < 3; ++k) {
      const Rcpp::List& s = l[names[k]].l;
      avg[4 * k] = s["map"].d;
      avg[4 * k + 1] = s["ci"].m.mem[0];
      avg[4 * k + 2] = s["ci"].m.mem[1];
      avg[4 * k + 3] = s["var"].d;
    }
  }
}

int ace_oracle_ncs_ncol(const double* knots, int K) { return (int)arma::unique(in_vec(knots, K)).n_elem; }
void ace_oracle_ncs_basis(const double* x, int n, const double* knots, int K, double* design) {
  out(design, ncs_basis(in_vec(x, n), in_vec(knots, K)));
}
void ace_oracle_ncs_basis_deriv(const double* x, int n, const double* knots, int K, double* design) {
  out(design, ncs_basis_deriv(in_vec(x, n), in_vec(knots, K)));
}

// normalize_train / normalize_test mutate their arguments in place (src/utilities_cpp.cpp:13-118)
void ace_oracle_normalize_train(double* y, double* X, double* Z, int n, int px, int pz, double* moments) {
  arma::vec yy = in_vec(y, n);
  arma::mat XX = in_mat(X, n, px), ZZ = in_mat(Z, n, pz);
  arma::mat mo = normalize_train(yy, XX, ZZ);
  out(y, yy);
  out(X, XX);
  out(Z, ZZ);
  out(moments, mo);
}
void ace_oracle_normalize_test(double* X, double* Z, int n, int px, int pz, const double* moments) {
  arma::mat XX = in_mat(X, n, px), ZZ = in_mat(Z, n, pz);
  normalize_test(XX, ZZ, in_mat(moments, 1 + px + pz, 3));
  out(X, XX);
  out(Z, ZZ);
}

}  // extern "C"
