// RcppArmadillo.h -- STAND-IN for the RcppArmadillo / Rcpp headers, test infrastructure only.
//
// The reference's native sources (/root/reference/src/*.cpp) include <RcppArmadillo.h>; neither R, Rcpp nor
// Armadillo exist in the build container.  This header implements exactly the slice of the two APIs those seven
// files use, so that the reference's OWN source files compile UNMODIFIED (oracle/Makefile target `_ref`) and can be
// executed to pin the oracle restatement (oracle/ace_oracle.cpp) and the CUDA path against the reference's literal
// loops, index arithmetic and quirks.  What it is NOT: Armadillo.  Differences that matter numerically:
//   * every operator is evaluated eagerly into a temporary (Armadillo fuses element-wise expressions; the
//     sequence of rounded operations per element is the same, so values agree bit for bit as long as the compiler
//     does not contract to FMA: build with -O2 and no -march, like R does);
//   * A * B goes to BLAS dgemm / dgemv (SciPy's bundled OpenBLAS, loaded with dlopen like the oracle), A * A.t()
//     to dsyrk + mirror, trace(A * B) is the O(n^2) sum Armadillo's op_trace specialisation computes, eig_sym is
//     LAPACK dsyevd ("dc", Armadillo's default);
//   * pow(X, k) is element-wise std::pow (Armadillo: eop_aux::pow);  stddev uses n - 1;  unique() sorts.
// Only double matrices exist; uvec / umat results (x > k, x == 0) are 0/1 doubles.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <cstdio>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

// ---------------------------------------------------------------------------------------------- BLAS / LAPACK hooks
namespace miniarma {
typedef void (*dgemm_fn)(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                         const int*, const double*, const int*, const double*, double*, const int*);
typedef void (*dgemv_fn)(const char*, const int*, const int*, const double*, const double*, const int*, const double*,
                         const int*, const double*, double*, const int*);
typedef void (*dsyrk_fn)(const char*, const char*, const int*, const int*, const double*, const double*, const int*,
                         const double*, double*, const int*);
typedef void (*dsyevd_fn)(const char*, const char*, const int*, double*, const int*, double*, double*, const int*, int*,
                          const int*, int*);
struct Blas {
  dgemm_fn dgemm = nullptr;
  dgemv_fn dgemv = nullptr;
  dsyrk_fn dsyrk = nullptr;
  dsyevd_fn dsyevd = nullptr;
};
Blas& blas();  // defined in the harness (oracle/ref_harness.cpp), filled by ace_oracle_init
}  // namespace miniarma

namespace arma {

typedef unsigned long long uword;
struct Mat;
struct Col;
struct Row;
struct subview;
struct diagview;

struct TransView {
  const Mat& m;
};

struct Mat {
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  double* mem = nullptr;
  std::vector<double> store;  // empty for a Mat that aliases cube memory
  bool owner = true;

  Mat() {}
  Mat(uword r, uword c) { init(r, c); }
  Mat(const Mat& o) {
    init(o.n_rows, o.n_cols);
    if (n_elem) std::memcpy(mem, o.mem, sizeof(double) * n_elem);
  }
  Mat(Mat&& o) noexcept { steal(o); }
  Mat(double* alias, uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(alias), owner(false) {}
  Mat(const subview& s);
  Mat(const diagview& d);
  virtual ~Mat() {}

  void init(uword r, uword c) {
    n_rows = r; n_cols = c; n_elem = r * c;
    store.assign(n_elem, 0.0);  // Armadillo leaves it uninitialised; zero is a valid "uninitialised"
    mem = store.data();
    owner = true;
  }
  void steal(Mat& o) {
    if (o.owner) {
      store = std::move(o.store);
      n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem;
      mem = store.data();
      owner = true;
      o.mem = nullptr; o.n_rows = o.n_cols = o.n_elem = 0;
    } else {
      init(o.n_rows, o.n_cols);
      if (n_elem) std::memcpy(mem, o.mem, sizeof(double) * n_elem);
    }
  }
  virtual void fix_shape(uword& r, uword& c) const { (void)r; (void)c; }
  Mat& assign(const Mat& o) {
    if (this == &o) return *this;
    uword r = o.n_rows, c = o.n_cols;
    fix_shape(r, c);
    if (!owner) {
      if (r * c != n_elem) throw std::logic_error("miniarma: size mismatch assigning into a cube slice");
    } else if (r != n_rows || c != n_cols) {
      init(r, c);
    }
    if (n_elem) std::memmove(mem, o.mem, sizeof(double) * n_elem);
    return *this;
  }
  Mat& operator=(const Mat& o) { return assign(o); }
  Mat& operator=(Mat&& o) {
    uword r = o.n_rows, c = o.n_cols;
    fix_shape(r, c);
    if (owner && o.owner && r == o.n_rows && c == o.n_cols) {
      steal(o);
      return *this;
    }
    return assign(o);
  }

  double& operator()(uword i) { return mem[i]; }
  const double& operator()(uword i) const { return mem[i]; }
  double& operator[](uword i) { return mem[i]; }
  const double& operator[](uword i) const { return mem[i]; }
  double& operator()(uword r, uword c) { return mem[r + n_rows * c]; }
  const double& operator()(uword r, uword c) const { return mem[r + n_rows * c]; }

  double* memptr() { return mem; }
  const double* memptr() const { return mem; }
  uword size() const { return n_elem; }
  Mat& zeros() { std::fill(mem, mem + n_elem, 0.0); return *this; }
  Mat& ones() { std::fill(mem, mem + n_elem, 1.0); return *this; }
  Mat& fill(double v) { std::fill(mem, mem + n_elem, v); return *this; }
  bool is_sorted(const char* dir) const {
    const bool asc = std::string(dir) == "ascend";
    for (uword i = 1; i < n_elem; ++i)
      if (asc ? (mem[i] < mem[i - 1]) : (mem[i] > mem[i - 1])) return false;
    return true;
  }
  TransView t() const { return TransView{*this}; }

  inline subview col(uword c);
  inline const subview col(uword c) const;
  inline subview row(uword r);
  inline const subview row(uword r) const;
  inline subview rows(uword a, uword b);
  inline const subview rows(uword a, uword b) const;
  inline subview submat(uword r1, uword c1, uword r2, uword c2);
  inline diagview diag();
  inline const diagview diag() const;

  Mat& operator+=(const Mat& o) { for (uword i = 0; i < n_elem; ++i) mem[i] += o.mem[i]; return *this; }
  Mat& operator-=(const Mat& o) { for (uword i = 0; i < n_elem; ++i) mem[i] -= o.mem[i]; return *this; }
};

struct Col : Mat {
  Col() { n_cols = 1; }
  explicit Col(uword n) : Mat(n, 1) {}
  Col(const Mat& m) : Mat(m.n_elem, 1) { if (n_elem) std::memcpy(mem, m.mem, sizeof(double) * n_elem); }
  Col(const Col& o) : Mat(static_cast<const Mat&>(o)) {}
  Col(const subview& s);
  Col(const diagview& d);
  void fix_shape(uword& r, uword& c) const override { r = r * c; c = 1; }
  Col& operator=(const Mat& o) { assign(o); return *this; }
  Col& operator=(const Col& o) { assign(o); return *this; }
  inline Col& operator=(const subview& s);
};

struct Row : Mat {
  Row() { n_rows = 1; }
  explicit Row(uword n) : Mat(1, n) {}
  Row(const Mat& m) : Mat(1, m.n_elem) { if (n_elem) std::memcpy(mem, m.mem, sizeof(double) * n_elem); }
  Row(const Row& o) : Mat(static_cast<const Mat&>(o)) {}
  void fix_shape(uword& r, uword& c) const override { c = r * c; r = 1; }
  Row& operator=(const Mat& o) { assign(o); return *this; }
  Row& operator=(const Row& o) { assign(o); return *this; }
};

typedef Mat mat;
typedef Col vec;
typedef Col colvec;
typedef Row rowvec;
typedef Col uvec;

// rectangular view into a Mat (col, row, rows, submat)
struct subview {
  Mat* m;
  uword r0, c0, nr, nc;
  double& at(uword i, uword j) const { return m->mem[(r0 + i) + m->n_rows * (c0 + j)]; }
  Mat eval() const {
    Mat out(nr, nc);
    for (uword j = 0; j < nc; ++j)
      for (uword i = 0; i < nr; ++i) out(i, j) = at(i, j);
    return out;
  }
  void set(const Mat& v) const {
    if (v.n_elem != nr * nc) throw std::logic_error("miniarma: subview assignment size mismatch");
    uword k = 0;
    for (uword j = 0; j < nc; ++j)
      for (uword i = 0; i < nr; ++i) at(i, j) = v.mem[k++];
  }
  const subview& operator=(const Mat& v) const { set(v); return *this; }
  const subview& operator=(const subview& s) const { set(s.eval()); return *this; }
  const subview& operator+=(const Mat& v) const {
    uword k = 0;
    for (uword j = 0; j < nc; ++j)
      for (uword i = 0; i < nr; ++i) at(i, j) += v.mem[k++];
    return *this;
  }
  const subview& operator-=(double s) const { for (uword j = 0; j < nc; ++j) for (uword i = 0; i < nr; ++i) at(i, j) -= s; return *this; }
  const subview& operator/=(double s) const { for (uword j = 0; j < nc; ++j) for (uword i = 0; i < nr; ++i) at(i, j) /= s; return *this; }
  void zeros() const { fill(0.0); }
  void ones() const { fill(1.0); }
  void fill(double v) const { for (uword j = 0; j < nc; ++j) for (uword i = 0; i < nr; ++i) at(i, j) = v; }
};

struct diagview {
  Mat* m;
  uword n() const { return std::min(m->n_rows, m->n_cols); }
  const diagview& operator+=(double s) const { for (uword i = 0; i < n(); ++i) (*m)(i, i) += s; return *this; }
};

inline Mat::Mat(const subview& s) { Mat t = s.eval(); steal(t); }
inline Mat::Mat(const diagview& d) { init(d.n(), 1); for (uword i = 0; i < n_elem; ++i) mem[i] = (*d.m)(i, i); }
inline Col::Col(const subview& s) : Mat(s.nr * s.nc, 1) { Mat t = s.eval(); std::memcpy(mem, t.mem, sizeof(double) * n_elem); }
inline Col::Col(const diagview& d) : Mat(d) {}
inline Col& Col::operator=(const subview& s) { assign(s.eval()); return *this; }

inline subview Mat::col(uword c) { return subview{this, 0, c, n_rows, 1}; }
inline const subview Mat::col(uword c) const { return subview{const_cast<Mat*>(this), 0, c, n_rows, 1}; }
inline subview Mat::row(uword r) { return subview{this, r, 0, 1, n_cols}; }
inline const subview Mat::row(uword r) const { return subview{const_cast<Mat*>(this), r, 0, 1, n_cols}; }
inline subview Mat::rows(uword a, uword b) { return subview{this, a, 0, b - a + 1, n_cols}; }
inline const subview Mat::rows(uword a, uword b) const { return subview{const_cast<Mat*>(this), a, 0, b - a + 1, n_cols}; }
inline subview Mat::submat(uword r1, uword c1, uword r2, uword c2) { return subview{this, r1, c1, r2 - r1 + 1, c2 - c1 + 1}; }
inline diagview Mat::diag() { return diagview{this}; }
inline const diagview Mat::diag() const { return diagview{const_cast<Mat*>(this)}; }

struct Cube {
  uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0;
  std::vector<double> store;
  double* mem = nullptr;
  std::vector<std::unique_ptr<Mat>> sl;
  Cube() {}
  Cube(uword r, uword c, uword s) { init(r, c, s); }
  Cube(const Cube& o) { init(o.n_rows, o.n_cols, o.n_slices); if (n_elem) std::memcpy(mem, o.mem, sizeof(double) * n_elem); }
  Cube& operator=(const Cube& o) {
    if (this == &o) return *this;
    if (o.n_rows != n_rows || o.n_cols != n_cols || o.n_slices != n_slices) init(o.n_rows, o.n_cols, o.n_slices);
    if (n_elem) std::memcpy(mem, o.mem, sizeof(double) * n_elem);
    return *this;
  }
  void init(uword r, uword c, uword s) {
    n_rows = r; n_cols = c; n_slices = s; n_elem = r * c * s;
    store.assign(n_elem, 0.0);
    mem = store.data();
    sl.clear();
    for (uword b = 0; b < s; ++b) sl.emplace_back(new Mat(mem + r * c * b, r, c));
  }
  Cube& zeros() { std::fill(mem, mem + n_elem, 0.0); return *this; }
  double* memptr() { return mem; }
  const double* memptr() const { return mem; }
  Mat& slice(uword b) { return *sl[b]; }
  const Mat& slice(uword b) const { return *sl[b]; }
  double& operator()(uword r, uword c, uword s) { return mem[r + n_rows * c + n_rows * n_cols * s]; }
  const double& operator()(uword r, uword c, uword s) const { return mem[r + n_rows * c + n_rows * n_cols * s]; }
};
typedef Cube cube;

namespace datum {
const double pi = 3.14159265358979323846264338327950288;
const double nan = std::numeric_limits<double>::quiet_NaN();
}

// ---------------------------------------------------------------------------------------------- element-wise algebra
#define MINIARMA_EW1(NAME, EXPR)                                  \
  inline Mat NAME(const Mat& a) {                                 \
    Mat o(a.n_rows, a.n_cols);                                    \
    for (uword i = 0; i < a.n_elem; ++i) { const double x = a.mem[i]; o.mem[i] = (EXPR); } \
    return o;                                                     \
  }
MINIARMA_EW1(exp, std::exp(x))
MINIARMA_EW1(log, std::log(x))
MINIARMA_EW1(sqrt, std::sqrt(x))
MINIARMA_EW1(abs, std::abs(x))
MINIARMA_EW1(operator-, -x)
#undef MINIARMA_EW1
inline Mat pow(const Mat& a, double k) {
  Mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = std::pow(a.mem[i], k);  // Armadillo: eop_aux::pow -> std::pow
  return o;
}
inline Cube sqrt(const Cube& a) {
  Cube o(a.n_rows, a.n_cols, a.n_slices);
  for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = std::sqrt(a.mem[i]);
  return o;
}

#define MINIARMA_EW2(OP, EXPR)                                                                        \
  inline Mat operator OP(const Mat& a, const Mat& b) {                                                \
    if (a.n_elem != b.n_elem) throw std::logic_error("miniarma: element-wise size mismatch");         \
    Mat o(a.n_rows, a.n_cols);                                                                        \
    for (uword i = 0; i < a.n_elem; ++i) { const double x = a.mem[i], y = b.mem[i]; o.mem[i] = (EXPR); } \
    return o;                                                                                         \
  }
MINIARMA_EW2(+, x + y)
MINIARMA_EW2(-, x - y)
MINIARMA_EW2(%, x * y)
MINIARMA_EW2(/, x / y)
#undef MINIARMA_EW2

#define MINIARMA_SC(OP, EXPR_MS, EXPR_SM)                                                      \
  inline Mat operator OP(const Mat& a, double s) {                                             \
    Mat o(a.n_rows, a.n_cols);                                                                 \
    for (uword i = 0; i < a.n_elem; ++i) { const double x = a.mem[i]; o.mem[i] = (EXPR_MS); }  \
    return o;                                                                                  \
  }                                                                                            \
  inline Mat operator OP(double s, const Mat& a) {                                             \
    Mat o(a.n_rows, a.n_cols);                                                                 \
    for (uword i = 0; i < a.n_elem; ++i) { const double x = a.mem[i]; o.mem[i] = (EXPR_SM); }  \
    return o;                                                                                  \
  }
MINIARMA_SC(+, x + s, s + x)
MINIARMA_SC(-, x - s, s - x)
MINIARMA_SC(*, x * s, s * x)
MINIARMA_SC(/, x / s, s / x)
#undef MINIARMA_SC

inline Mat operator>(const Mat& a, double s) {
  Mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = (a.mem[i] > s) ? 1.0 : 0.0;
  return o;
}
inline Mat operator==(const Mat& a, double s) {
  Mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = (a.mem[i] == s) ? 1.0 : 0.0;
  return o;
}

// ---------------------------------------------------------------------------------------------- products
// A * B stays lazy so that trace(A * B) can take Armadillo's O(n^2) route; everything else evaluates it with BLAS.
struct Prod {
  const Mat& a;
  const Mat& b;
  bool tb;  // B transposed
  Mat eval() const {
    const int M = (int)a.n_rows, K = (int)a.n_cols;
    const int N = (int)(tb ? b.n_rows : b.n_cols), Kb = (int)(tb ? b.n_cols : b.n_rows);
    if (K != Kb) throw std::logic_error("miniarma: matrix product size mismatch");
    Mat o((uword)M, (uword)N);
    const double one = 1.0, zero = 0.0;
    const int ione = 1;
    miniarma::Blas& bl = miniarma::blas();
    if (M == 0 || N == 0) return o;
    if (tb && &a == &b) {  // A * A.t(): Armadillo's syrk path, result mirrored (exactly symmetric)
      bl.dsyrk("U", "N", &M, &K, &one, a.mem, &M, &zero, o.mem, &M);
      for (int j = 0; j < M; ++j)
        for (int i = j + 1; i < M; ++i) o.mem[i + (size_t)M * j] = o.mem[j + (size_t)M * i];
      return o;
    }
    if (!tb && N == 1) {
      bl.dgemv("N", &M, &K, &one, a.mem, &M, b.mem, &ione, &zero, o.mem, &ione);
      return o;
    }
    const int ldb = (int)b.n_rows;
    bl.dgemm("N", tb ? "T" : "N", &M, &N, &K, &one, a.mem, &M, b.mem, &ldb, &zero, o.mem, &M);
    return o;
  }
  operator Mat() const { return eval(); }
};
inline Prod operator*(const Mat& a, const Mat& b) { return Prod{a, b, false}; }
inline Prod operator*(const Mat& a, const TransView& t) { return Prod{a, t.m, true}; }

// Armadillo's op_trace for a product: sum_k sum_i A(k,i) B(i,k), two running sums, no n x n temporary
inline double trace(const Prod& p) {
  const Mat& A = p.a;
  const Mat& B = p.b;
  if (p.tb) return 0.0 / 0.0;  // not used by the reference
  const uword N = std::min(A.n_rows, B.n_cols), K = A.n_cols;
  double acc1 = 0.0, acc2 = 0.0;
  for (uword k = 0; k < N; ++k) {
    const double* bcol = B.mem + B.n_rows * k;
    uword i, j;
    for (i = 0, j = 1; j < K; i += 2, j += 2) {
      acc1 += A(k, i) * bcol[i];
      acc2 += A(k, j) * bcol[j];
    }
    if (i < K) acc1 += A(k, i) * bcol[i];
  }
  return acc1 + acc2;
}
inline double trace(const Mat& A) {
  double s = 0.0;
  for (uword i = 0; i < std::min(A.n_rows, A.n_cols); ++i) s += A(i, i);
  return s;
}

// ---------------------------------------------------------------------------------------------- reductions
// Armadillo's arrayops::accumulate: two running sums over pairs of elements
inline double accu(const Mat& a) {
  double acc1 = 0.0, acc2 = 0.0;
  uword i, j;
  for (i = 0, j = 1; j < a.n_elem; i += 2, j += 2) {
    acc1 += a.mem[i];
    acc2 += a.mem[j];
  }
  if (i < a.n_elem) acc1 += a.mem[i];
  return acc1 + acc2;
}
inline double sum(const Mat& a) { return accu(a); }  // the reference only sums vectors
inline double mean(const Mat& a) { return accu(a) / (double)a.n_elem; }
inline double dot(const Mat& a, const Mat& b) {
  if (a.n_elem != b.n_elem) throw std::logic_error("miniarma: dot size mismatch");
  double acc1 = 0.0, acc2 = 0.0;
  uword i, j;
  for (i = 0, j = 1; j < a.n_elem; i += 2, j += 2) {
    acc1 += a.mem[i] * b.mem[i];
    acc2 += a.mem[j] * b.mem[j];
  }
  if (i < a.n_elem) acc1 += a.mem[i] * b.mem[i];
  return acc1 + acc2;
}
inline double norm(const Mat& a) { return std::sqrt(dot(a, a)); }
inline double min(const Mat& a) { return *std::min_element(a.mem, a.mem + a.n_elem); }
inline double max(const Mat& a) { return *std::max_element(a.mem, a.mem + a.n_elem); }
inline double median(const Mat& a) {
  std::vector<double> t(a.mem, a.mem + a.n_elem);
  std::sort(t.begin(), t.end());
  const size_t n = t.size();
  return (n % 2) ? t[n / 2] : 0.5 * (t[n / 2 - 1] + t[n / 2]);
}
inline double stddev(const Mat& a) {  // Armadillo's op_var::direct_var, norm_type 0 (n - 1)
  const uword n = a.n_elem;
  const double acc1 = mean(a);
  double acc2 = 0.0, acc3 = 0.0;
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) {
    const double ti = acc1 - a.mem[i], tj = acc1 - a.mem[j];
    acc2 += ti * ti + tj * tj;
    acc3 += ti + tj;
  }
  if (i < n) {
    const double ti = acc1 - a.mem[i];
    acc2 += ti * ti;
    acc3 += ti;
  }
  return std::sqrt((acc2 - acc3 * acc3 / (double)n) / (double)(n - 1));
}
inline Col unique(const Mat& a) {
  std::vector<double> t(a.mem, a.mem + a.n_elem);
  std::sort(t.begin(), t.end());
  t.erase(std::unique(t.begin(), t.end()), t.end());
  Col o((uword)t.size());
  std::copy(t.begin(), t.end(), o.mem);
  return o;
}
inline Col sort(const Mat& a, const char* dir) {
  Col o(a);
  if (std::string(dir) == "ascend") std::sort(o.mem, o.mem + o.n_elem);
  else std::sort(o.mem, o.mem + o.n_elem, [](double x, double y) { return x > y; });
  return o;
}
inline bool is_finite(const Mat& a) {
  for (uword i = 0; i < a.n_elem; ++i)
    if (!std::isfinite(a.mem[i])) return false;
  return true;
}
inline bool is_finite(double x) { return std::isfinite(x); }

template <typename T>
struct conv_to {
  static T from(const Mat& m) { return T(m); }
};

// eig_sym(eigval, eigvec, X): LAPACK dsyevd (Armadillo's default method "dc"); eigvec may alias X
inline bool eig_sym(Col& eigval, Mat& eigvec, const Mat& X) {
  const int n = (int)X.n_rows;
  if (&eigvec != &X) eigvec = X;
  eigval = Col((uword)n);
  int info = 0, lwork = -1, liwork = -1, iwq = 0;
  double wq = 0.0;
  miniarma::blas().dsyevd("V", "U", &n, eigvec.mem, &n, eigval.mem, &wq, &lwork, &iwq, &liwork, &info);
  lwork = (int)wq;
  liwork = iwq;
  std::vector<double> work((size_t)std::max(lwork, 1));
  std::vector<int> iwork((size_t)std::max(liwork, 1));
  miniarma::blas().dsyevd("V", "U", &n, eigvec.mem, &n, eigval.mem, work.data(), &lwork, iwork.data(), &liwork, &info);
  return info == 0;
}

}  // namespace arma

// ---------------------------------------------------------------------------------------------- Rcpp slice
namespace Rcpp {

struct Value;
struct List {
  std::vector<std::pair<std::string, std::shared_ptr<Value>>> items;
  template <typename... Args>
  static List create(const Args&... args);
  const Value& operator[](const std::string& name) const {
    for (auto& it : items)
      if (it.first == name) return *it.second;
    throw std::out_of_range("Rcpp::List: no element named " + name);
  }
};

struct Value {
  int kind = 0;  // 0 scalar, 1 matrix / vector, 2 cube, 3 list
  double d = 0.0;
  arma::Mat m;
  arma::Cube c;
  List l;
  Value(double x) : kind(0), d(x) {}
  Value(const arma::Mat& x) : kind(1), m(x) {}   // a copy, like Rcpp::wrap
  Value(const arma::Cube& x) : kind(2), c(x) {}
  Value(const List& x) : kind(3), l(x) {}
};

struct NamedValue {
  std::string name;
  std::shared_ptr<Value> v;
};
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  template <typename T>
  NamedValue operator=(const T& x) const { return NamedValue{name, std::make_shared<Value>(x)}; }
};
struct Placeholder {
  Named operator()(const char* n) const { return Named(n); }
};
static const Placeholder _;

template <typename... Args>
List List::create(const Args&... args) {
  List out;
  const NamedValue arr[] = {args...};
  for (const NamedValue& nv : arr) out.items.emplace_back(nv.name, nv.v);
  return out;
}

inline void checkUserInterrupt() {}
static std::ostream& Rcout = std::cout;

// Rcpp::stop(fmt, ...): an R error; here a C++ exception carrying the formatted message
struct exception : std::runtime_error {
  explicit exception(const std::string& m) : std::runtime_error(m) {}
};
template <typename... Args>
[[noreturn]] inline void stop(const char* fmt, Args... args) {
  char buf[1024];
  if (sizeof...(args) == 0) std::snprintf(buf, sizeof buf, "%s", fmt);
  else std::snprintf(buf, sizeof buf, fmt, args...);
  throw exception(buf);
}

// external pointers: SEXP is an opaque handle; XPtr<T, Storage, Finalizer> wraps a T* (the finaliser runs when R
// collects the object; here: never -- the harness destroys what it creates)
}  // namespace Rcpp
typedef void* SEXP;
namespace Rcpp {
template <typename T>
struct PreserveStorage {};
template <typename T>
void standard_delete_finalizer(T* p) { delete p; }
template <typename T, template <class> class Storage = PreserveStorage, void Finalizer(T*) = standard_delete_finalizer<T>,
          bool finalizeOnExit = false>
struct XPtr {
  T* p;
  explicit XPtr(T* ptr, bool set_delete_finalizer = true) : p(ptr) { (void)set_delete_finalizer; }
  XPtr(SEXP s) : p(static_cast<T*>(s)) {}
  operator SEXP() const { return static_cast<SEXP>(p); }
  T* get() const { return p; }
  T* operator->() const { return p; }
};

}  // namespace Rcpp
