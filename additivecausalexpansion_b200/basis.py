"""Treatment-basis classes with the reference's R6 interface (R/spline_linear_R6.R, spline_square_R6.R,
spline_ns_R6.R, spline_B_R6.R): fields `B`, `dB`, methods `dim()`, `trainbasis(Z, n_knots)`,
`testbasis(Znew)`.  The basis VALUES are inputs of the hot path (O(n*K) host work in the reference too).

* knots: base-R `quantile(Z, probs = seq(n_knots)/(n_knots+1))` (type 7), boundary knots (-1, 1).
* ns ("cubic", "ncs", "ns" all map here, R/utilities.R:27-28): the package's own truncated-power design,
  evaluated by the library's `ace_ncs_basis` (bit-exactness against the oracle is tested).
* B: the reference calls splines2::bSpline / dbs (un-vendored third party, no pinned version in
  DESCRIPTION:22).  Restated here with the published Cox-de Boor recursion on the clamped knot vector
  (intercept = FALSE drops the first function).  PARITY UNPINNED for this class.
"""
from __future__ import annotations

import numpy as np

from . import api


def quantile7(x, probs):
    """R's default quantile (type 7): index = 1 + (n-1) p; (1-h) x[lo] + h x[hi]."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = x.size
    out = []
    for pr in np.atleast_1d(probs):
        index = (n - 1) * float(pr)
        lo = int(np.floor(index))
        hi = int(np.ceil(index))
        h = index - lo
        q = x[lo]
        if index > lo and x[hi] != q:
            q = (1 - h) * q + h * x[hi]
        out.append(q)
    return np.array(out)


def knot_interval_index(z, knots):
    """Index of the knot interval each z falls in (right-continuous, last interval closed): the
    'knot indices' that must be bit-exact between implementations."""
    k = np.asarray(knots, dtype=np.float64)
    idx = np.searchsorted(k, np.asarray(z, dtype=np.float64).ravel(), side="right") - 1
    return np.clip(idx, 0, k.size - 2).astype(np.int64)


def bspline_design(x, interior, boundary=(-1.0, 1.0), degree=3, deriv=0):
    """Cox-de Boor B-spline design without intercept: len(interior) + degree columns."""
    x = np.asarray(x, dtype=np.float64).ravel()
    a, b = boundary
    interior = np.asarray(interior, dtype=np.float64).ravel()
    t = np.concatenate([[a] * (degree + 1), interior, [b] * (degree + 1)])
    nb = interior.size + degree + 1  # including the intercept function
    # degree-0 functions on each knot span (right end closed on the last non-empty span)
    span = knot_interval_index(x, t[degree:degree + interior.size + 2]) + degree
    N = np.zeros((x.size, t.size - 1))
    N[np.arange(x.size), span] = 1.0
    dN = np.zeros_like(N)
    for k in range(1, degree + 1):
        Nk = np.zeros((x.size, t.size - 1 - k))
        dNk = np.zeros_like(Nk)
        for i in range(t.size - 1 - k):
            d1, d2 = t[i + k] - t[i], t[i + k + 1] - t[i + 1]
            if d1 > 0:
                Nk[:, i] += (x - t[i]) / d1 * N[:, i]
                dNk[:, i] += N[:, i] / d1 + (x - t[i]) / d1 * dN[:, i]
            if d2 > 0:
                Nk[:, i] += (t[i + k + 1] - x) / d2 * N[:, i + 1]
                dNk[:, i] += -N[:, i + 1] / d2 + (t[i + k + 1] - x) / d2 * dN[:, i + 1]
        N, dN = Nk, dNk
    out = dN if deriv else N
    return np.asfortranarray(out[:, 1:nb])


class _Basis:
    B = None
    dB = None

    def dim(self):
        return self.B.shape[1] + 1  # basis plus nuisance term m()


class linear_spline(_Basis):
    """R/spline_linear_R6.R (also the "binary" basis)."""

    def trainbasis(self, Z, n_knots=None, verbose=False):
        z = np.asarray(Z, dtype=np.float64).ravel()
        self.B = np.asfortranarray(z.reshape(-1, 1).copy())
        self.dB = np.ones((z.size, 1), order="F")
        return self.B

    def testbasis(self, Znew=None):
        if Znew is None:
            return {"B": self.B, "dB": self.dB}
        z = np.asarray(Znew, dtype=np.float64).ravel()
        return {"B": np.asfortranarray(z.reshape(-1, 1).copy()), "dB": np.ones((z.size, 1), order="F")}


class square_spline(_Basis):
    """R/spline_square_R6.R."""

    @staticmethod
    def _mk(z):
        return (np.asfortranarray(np.stack([z, z ** 2], axis=1)),
                np.asfortranarray(np.stack([np.ones_like(z), 2 * z], axis=1)))

    def trainbasis(self, Z, n_knots=None, verbose=False):
        self.B, self.dB = self._mk(np.asarray(Z, dtype=np.float64).ravel())
        return self.B

    def testbasis(self, Znew=None):
        if Znew is None:
            return {"B": self.B, "dB": self.dB}
        B, dB = self._mk(np.asarray(Znew, dtype=np.float64).ravel())
        return {"B": B, "dB": dB}


class ns_spline(_Basis):
    """R/spline_ns_R6.R: natural cubic spline, knots = type-7 quantiles + boundary (-1, 1)."""

    myknots = None

    def trainbasis(self, Z, n_knots, verbose=False):
        z = np.asarray(Z, dtype=np.float64).ravel()
        ik = quantile7(z, np.arange(1, n_knots + 1) / (n_knots + 1)) if n_knots > 0 else np.array([])
        self.myknots = np.concatenate([ik, [-1.0, 1.0]])
        self.B = api.ncs_basis(z, self.myknots)
        self.dB = api.ncs_basis_deriv(z, self.myknots)
        return self.B

    def testbasis(self, Znew=None):
        if Znew is None:
            return {"B": self.B, "dB": self.dB}
        z = np.asarray(Znew, dtype=np.float64).ravel()
        return {"B": api.ncs_basis(z, self.myknots), "dB": api.ncs_basis_deriv(z, self.myknots)}


class B_spline(_Basis):
    """R/spline_B_R6.R.  `m` is the spline ORDER (degree = m - 1).  Through ace.train the reference passes
    `verbose` in the position of `m` (quirk Q9: degree 0 when verbose=TRUE); ace_train() reproduces that
    call, direct users of this class get the documented cubic default."""

    knots = None
    degree = 3

    def trainbasis(self, Z, n_knots, m=4, verbose=False):
        z = np.asarray(Z, dtype=np.float64).ravel()
        self.degree = int(m) - 1
        if self.degree < 0:
            raise ValueError("'degree' must be a nonnegative integer.")  # what splines2 raises for m = FALSE
        self.knots = quantile7(z, np.arange(1, n_knots + 1) / (n_knots + 1))
        self.B = bspline_design(z, self.knots, (-1.0, 1.0), self.degree, 0)
        self.dB = bspline_design(z, self.knots, (-1.0, 1.0), self.degree, 1)
        return self.B

    def testbasis(self, Znew=None):
        if Znew is None:
            return {"B": self.B, "dB": self.dB}
        z = np.asarray(Znew, dtype=np.float64).ravel()
        return {"B": bspline_design(z, self.knots, (-1.0, 1.0), self.degree, 0),
                "dB": bspline_design(z, self.knots, (-1.0, 1.0), self.degree, 1)}


def set_basis(basis, isuniv=True):
    """R/utilities.R:23-30."""
    if basis in ("binary", "linear"):
        return linear_spline()
    if isuniv and basis == "B":
        return B_spline()
    if isuniv and basis == "square":
        return square_spline()
    if isuniv:
        return ns_spline()  # "cubic", "ncs", "ns", anything else
    raise ValueError("multivariate Z only supports the linear basis")
