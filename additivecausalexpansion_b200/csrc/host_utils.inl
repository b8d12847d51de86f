// host_utils.inl -- the O(n) preprocessing routines that live in the same DLL as the hot path in the
// reference (ncs_basis, ncs_basis_deriv, normalize_train, normalize_test).  They are one-off host
// work in the reference too (out of GPU scope); they are here so that the library is a
// complete stand-in for the reference's registered routines.  Written from the behaviour of
// src/ncs_basis_cpp.cpp and src/utilities_cpp.cpp:13-118, including the literal index quirks.

namespace ace {

static std::vector<double> sorted_unique(const double* v, int n) {
  std::vector<double> k(v, v + n);
  std::sort(k.begin(), k.end());
  k.erase(std::unique(k.begin(), k.end()), k.end());
  return k;
}

// truncated-power natural-cubic-spline design (order 3) or its derivative (order 2, factor 3)
static int ncs_design(const double* x, int n, const double* knots_in, int nknots, double* design, bool deriv) {
  if (!x || !knots_in || n < 0 || nknots < 2) return usage("ncs_basis: bad argument");
  const std::vector<double> kn = sorted_unique(knots_in, nknots);  // src/ncs_basis_cpp.cpp:65-68
  const int K = (int)kn.size();
  if (!design) return K;
  if (K < 2) return usage("ncs_basis: need at least two distinct knots");
  std::vector<double> dk((size_t)n * K);
  auto tp = [&](double xv, double c) {
    const double ind = (xv > c) ? 1.0 : 0.0, t = xv - c;
    // arma::pow(x - k, 3) / arma::pow(x - k, 2) (src/ncs_basis_cpp.cpp:15-17,44-46) are element-wise std::pow calls in
    // Armadillo (eop_aux::pow); t * t * t would round twice and differ by 1 ulp in a sizeable fraction of inputs
    return deriv ? 3 * ind * std::pow(t, 2.0) : ind * std::pow(t, 3.0);
  };
  for (int r = 0; r < n; ++r) dk[r + (size_t)n * (K - 1)] = tp(x[r], kn[K - 1]);
  for (int i = 0; i < K - 1; ++i)
    for (int r = 0; r < n; ++r) {
      const double v = tp(x[r], kn[i]) - dk[r + (size_t)n * (K - 1)];
      dk[r + (size_t)n * i] = v / (kn[K - 1] - kn[i]);
    }
  for (int r = 0; r < n; ++r) design[r] = deriv ? 1.0 : x[r];
  for (int i = 0; i < K - 2; ++i)
    for (int r = 0; r < n; ++r) design[r + (size_t)n * (1 + i)] = dk[r + (size_t)n * i] - dk[r + (size_t)n * (K - 2)];
  for (int r = 0; r < n; ++r) design[r + (size_t)n * (K - 1)] = -dk[r + (size_t)n * (K - 2)];
  return K;
}

// Armadillo's arrayops::accumulate: two running sums over the even / odd elements (what arma::mean, arma::sum
// and arma::accu of a contiguous vector reduce to)
static double arma_accumulate(const double* x, int n) {
  double acc1 = 0.0, acc2 = 0.0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) {
    acc1 += x[i];
    acc2 += x[j];
  }
  if (i < n) acc1 += x[i];
  return acc1 + acc2;
}

static double col_median(const double* c, int n) {
  std::vector<double> t(c, c + n);
  std::sort(t.begin(), t.end());
  return (n % 2) ? t[n / 2] : 0.5 * (t[n / 2 - 1] + t[n / 2]);
}

}  // namespace ace

extern "C" {

int ace_ncs_basis(const double* x, int n, const double* knots, int nknots, double* design) {
  return ncs_design(x, n, knots, nknots, design, false);
}

int ace_ncs_basis_deriv(const double* x, int n, const double* knots, int nknots, double* design) {
  return ncs_design(x, n, knots, nknots, design, true);
}

int ace_normalize_train(double* y, double* X, double* Z, int n, int px, int pz, double* moments) {
  if (!y || !X || !Z || !moments || n < 2 || px < 1 || pz < 1) return usage("normalize_train: bad argument");
  const int R = 1 + px + pz;
  auto M = [&](int r, int c) -> double& { return moments[r + (size_t)R * c]; };
  for (int r = 0; r < R; ++r) {
    M(r, 0) = 0.0;
    M(r, 1) = 1.0;
    M(r, 2) = 0.0;
  }
  std::vector<int> isbinary(px + pz, 0);
  auto col = [&](int i) -> double* { return i < px ? X + (size_t)n * i : Z + (size_t)n * (i - px); };
  for (int i = 0; i < px + pz; ++i) {  // src/utilities_cpp.cpp:27-66
    double* c = col(i);
    const std::vector<double> u = sorted_unique(c, n);
    if (u.size() == 2) {
      isbinary[i] = 1;
      M(i + 1, 2) = 1.0;
      if (u.front() != 0) M(i, 0) = u.front();                 // row i, not i+1: as in the reference
      if (u.back() != 1) M(i, 1) = u.back() - u.front();
      for (int r = 0; r < n; ++r) c[r] = (c[r] - M(i, 0)) / M(i, 1);
    } else if (u.size() == 1) {
      if (i < px) {
        for (int r = 0; r < n; ++r) c[r] = 0.0;
      } else {
        if (i >= pz) return usage("normalize_train: constant Z column (the reference indexes out of bounds here)");
        double* zc = Z + (size_t)n * i;                        // Z.col(i) with i >= px: as in the reference
        for (int r = 0; r < n; ++r) zc[r] = 0.0;
      }
    }
  }
  M(0, 0) = arma_accumulate(y, n) / n;                          // :69 mean(y)
  for (int r = 0; r < n; ++r) y[r] -= M(0, 0);
  for (int i = 1; i <= px + pz; ++i) {                         // :71-82
    if (isbinary[i - 1] == 0) {
      double* c = col(i - 1);
      M(i, 0) = col_median(c, n);
      for (int r = 0; r < n; ++r) c[r] -= M(i, 0);
    }
  }
  {                                                            // :84-85  arma::stddev, n-1 form (op_var::direct_var)
    const double mean = arma_accumulate(y, n) / n;
    double a2 = 0.0, a3 = 0.0;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
      const double ti = mean - y[i], tj = mean - y[j];
      a2 += ti * ti + tj * tj;
      a3 += ti + tj;
    }
    if (i < n) {
      const double ti = mean - y[i];
      a2 += ti * ti;
      a3 += ti;
    }
    M(0, 1) = std::sqrt((a2 - a3 * a3 / n) / (n - 1));
    for (int r = 0; r < n; ++r) y[r] /= M(0, 1);
  }
  for (int i = 1; i <= px; ++i) {                              // :86-94
    if (isbinary[i - 1] == 0) {
      double* c = col(i - 1);
      double mx = 0.0;
      for (int r = 0; r < n; ++r) mx = std::max(mx, std::fabs(c[r]));
      M(i, 1) = mx;
      for (int r = 0; r < n; ++r) c[r] /= mx;
    }
  }
  for (int i = px + 1; i <= px + pz; ++i) {                    // :96-101  (index i-px-1 into isbinary: literal)
    if (isbinary[i - px - 1] == 0) {
      double* c = Z + (size_t)n * (i - px - 1);
      double mx = 0.0;
      for (int r = 0; r < n; ++r) mx = std::max(mx, std::fabs(c[r]));
      M(i, 1) = mx;
      for (int r = 0; r < n; ++r) c[r] = c[r] / mx;
    }
  }
  return 0;
}

int ace_normalize_test(double* X, double* Z, int n, int px, int pz, const double* moments) {
  if (!X || !Z || !moments) return usage("normalize_test: bad argument");
  const int R = 1 + px + pz;
  for (int i = 0; i < px; ++i)                                 // src/utilities_cpp.cpp:108-118
    for (int r = 0; r < n; ++r)
      X[r + (size_t)n * i] = (X[r + (size_t)n * i] - moments[i + 1]) / moments[i + 1 + R];
  for (int i = 0; i < pz; ++i)
    for (int r = 0; r < n; ++r)
      Z[r + (size_t)n * i] = (Z[r + (size_t)n * i] - moments[i + 1 + px]) / moments[i + 1 + px + R];
  return 0;
}

}  // extern "C"
