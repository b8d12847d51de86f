"""CPU oracle for the ACE empirical-Bayes GP hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``additivecausalexpansion_b200``) never does; it has no CPU fallback.

``ace_oracle.cpp`` is a literal C++ restatement of the reference's native files
(file:line cited per function there); this module is the ctypes face of it with the
reference's own function names and argument order (``R/RcppExports.R:4-78``).
The reference has no tests or golden vectors of its own.  What pins this restatement:
``oracle/_ref/libace_ref.so`` -- the reference's OWN native sources (``/root/reference/src/*.cpp``) compiled
unmodified against a small stand-in for the RcppArmadillo headers (``oracle/miniarma/``; R, Rcpp and Armadillo do
not exist in the build container), driven through the same ctypes code (``using_reference()``);
``tests/test_oracle_vs_reference.py`` compares the two function by function and ``tests/golden/ref_golden.npz``
carries reference-generated vectors to machines that do not have ``/root/reference``.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)


def _find_openblas() -> str:
    import scipy

    pats = [
        os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"),
        os.path.join(os.path.dirname(scipy.__file__), ".libs", "libscipy_openblas-*.so"),
    ]
    for p in pats:
        hits = sorted(glob.glob(p))
        if hits:
            return os.path.abspath(hits[0])
    raise RuntimeError("SciPy's bundled OpenBLAS not found")


def build(force: bool = False) -> str:
    """Compile oracle/libace_oracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "libace_oracle.so")
    src = os.path.join(_HERE, "ace_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libace_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib(nthreads: int = 0):
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.ace_oracle_init.argtypes = [C.c_char_p, C.c_int]
        L.ace_oracle_init.restype = C.c_int
        rc = L.ace_oracle_init(_find_openblas().encode(), int(nthreads))
        if rc != 0:
            raise RuntimeError(f"ace_oracle_init failed ({rc})")
        L.ace_oracle_mu_solution.restype = C.c_double
        L.ace_oracle_threads.restype = C.c_int
        _LIB = L
    elif nthreads:
        _LIB.ace_oracle_init(_find_openblas().encode(), int(nthreads))
    return _LIB


def threads() -> int:
    return int(lib().ace_oracle_threads())


# --------------------------------------------------------------------------- the compiled reference (oracle/_ref)
REFERENCE_SRC = "/root/reference/src"
_REF_LIB = None


def reference_so() -> str:
    return os.path.join(_HERE, "_ref", "libace_ref.so")


def build_ref(force: bool = False):
    """Compile the reference's own sources (where they lie under /root/reference/src) + oracle/ref_harness.cpp
    against oracle/miniarma into oracle/_ref/libace_ref.so.  Returns the path, or None when the reference sources
    are not on this machine and no prebuilt library travelled here."""
    so = reference_so()
    if os.path.isdir(REFERENCE_SRC):
        deps = [os.path.join(_HERE, "ref_harness.cpp"), os.path.join(_HERE, "miniarma", "RcppArmadillo.h")]
        stale = (not os.path.exists(so)) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps)
        if force or stale:
            subprocess.check_call(["make", "-C", _HERE, "_ref"] + (["-B"] if force else []),
                                  stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


def reference_available() -> bool:
    return build_ref() is not None


def ref_lib(nthreads: int = 0):
    global _REF_LIB
    if _REF_LIB is None:
        so = build_ref()
        if so is None:
            raise RuntimeError("oracle/_ref/libace_ref.so is not available (the reference sources are not here)")
        L = C.CDLL(so)
        L.ace_oracle_init.argtypes = [C.c_char_p, C.c_int]
        L.ace_oracle_init.restype = C.c_int
        rc = L.ace_oracle_init(_find_openblas().encode(), int(nthreads))
        if rc != 0:
            raise RuntimeError(f"ace_ref init failed ({rc})")
        L.ace_oracle_mu_solution.restype = C.c_double
        L.ace_oracle_threads.restype = C.c_int
        _REF_LIB = L
    return _REF_LIB


class using_reference:
    """``with oracle.using_reference(): oracle.grad_SE_cpp(...)`` runs the per-function wrappers of this module on
    the compiled reference sources instead of the restatement (the R-level sequencing -- OracleFit -- exists only in
    the restatement: the reference keeps it in R)."""

    def __enter__(self):
        global _LIB
        lib()
        self._saved = _LIB
        _LIB = ref_lib()
        return self

    def __exit__(self, *exc):
        global _LIB
        _LIB = self._saved


def _f(a, ndim=None):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if ndim == 2 and a.ndim == 1:
        a = np.asfortranarray(a.reshape(-1, 1))
    return a


def _p(a):
    return a.ctypes.data_as(_dp)


def _d(x):
    return C.c_double(float(x))


# --------------------------------------------------------------------------- kernel builds
def _kernmat(fn, X1, X2, Z1, Z2, parameters):
    X1, X2, Z1, Z2 = _f(X1, 2), _f(X2, 2), _f(Z1, 2), _f(Z2, 2)
    par = _f(parameters).ravel()
    n1, n2, p, Bz = X1.shape[0], X2.shape[0], X2.shape[1], Z1.shape[1]
    full = np.empty((n1, n2), order="F")
    elements = np.empty((n1, n2, Bz + 1), order="F")
    fn(_p(X1), _p(X2), _p(Z1), _p(Z2), n1, n2, p, Bz, _p(par), _p(full), _p(elements))
    return {"full": full, "elements": elements}


def _kernmat_sym(fn, X, Z, parameters):
    X, Z = _f(X, 2), _f(Z, 2)
    par = _f(parameters).ravel()
    n, p, Bz = X.shape[0], X.shape[1], Z.shape[1]
    full = np.empty((n, n), order="F")
    elements = np.empty((n, n, Bz + 1), order="F")
    fn(_p(X), _p(Z), n, p, Bz, _p(par), _p(full), _p(elements))
    return {"full": full, "elements": elements}


def kernmat_SE_cpp(X1, X2, Z1, Z2, parameters):
    return _kernmat(lib().ace_oracle_kernmat_SE, X1, X2, Z1, Z2, parameters)


def kernmat_SE_symmetric_cpp(X, Z, parameters):
    return _kernmat_sym(lib().ace_oracle_kernmat_SE_sym, X, Z, parameters)


def kernmat_Matern32_cpp(X1, X2, Z1, Z2, parameters):
    return _kernmat(lib().ace_oracle_kernmat_Matern32, X1, X2, Z1, Z2, parameters)


def kernmat_Matern32_symmetric_cpp(X, Z, parameters):
    return _kernmat_sym(lib().ace_oracle_kernmat_Matern32_sym, X, Z, parameters)


# --------------------------------------------------------------------------- inverse
def invkernel_cpp(pdmat, sigma, chol: bool = False):
    K = _f(pdmat)
    n = K.shape[0]
    eig = np.empty(n)
    inv = np.empty((n, n), order="F")
    fn = lib().ace_oracle_invkernel_chol if chol else lib().ace_oracle_invkernel
    fn(_p(K), n, _d(sigma), _p(eig), _p(inv))
    return {"eigenval": eig, "inv": inv}


# --------------------------------------------------------------------------- gradients / stats
def _grad(fn, y, X, Z, Kfull, K, invKmatn, eigenval, parameters, stats, B, std_y):
    y, X = _f(y).ravel(), _f(X, 2)
    Kfull, K, invKmatn = _f(Kfull), _f(K), _f(invKmatn)
    eigenval, par = _f(eigenval).ravel(), _f(parameters).ravel()
    n, p = X.shape
    g = np.empty(par.size)
    st = np.zeros(2)
    fn(_p(y), _p(X), _p(Kfull), _p(K), _p(invKmatn), _p(eigenval), _p(par), n, p, int(B), _d(std_y),
       _p(st), _p(g))
    stats[:] = st  # in-place through the reference's `arma::vec& stats`
    return g


def grad_SE_cpp(y, X, Z, Kfull, K, invKmatn, eigenval, parameters, stats, B, std_y):
    return _grad(lib().ace_oracle_grad_SE, y, X, Z, Kfull, K, invKmatn, eigenval, parameters, stats, B, std_y)


def grad_Matern_cpp(y, X, Z, Kfull, K, invKmatn, eigenval, parameters, stats, B, std_y):
    return _grad(lib().ace_oracle_grad_Matern, y, X, Z, Kfull, K, invKmatn, eigenval, parameters, stats, B,
                 std_y)


def stats_cpp(y, Kmat, invKmatn, eigenval, mu, std_y=1.0):
    y, Kmat, invKmatn, eigenval = _f(y).ravel(), _f(Kmat), _f(invKmatn), _f(eigenval).ravel()
    st = np.zeros(2)
    lib().ace_oracle_stats(_p(y), _p(Kmat), _p(invKmatn), _p(eigenval), _d(mu), _d(std_y), y.size, _p(st))
    return st


def mu_solution_cpp(y, invKmat):
    y, invKmat = _f(y).ravel(), _f(invKmat)
    return float(lib().ace_oracle_mu_solution(_p(y), _p(invKmat), y.size))


def norm_clip_cpp(flag, grads, max_length):
    assert grads.dtype == np.float64 and grads.flags.c_contiguous
    lib().ace_oracle_norm_clip(int(bool(flag)), _p(grads), grads.size, _d(max_length))


def Nadam_cpp(iter, learn_rate, beta1, beta2, eps, m, v, grad, para):
    g = _f(grad).ravel()
    return bool(lib().ace_oracle_Nadam(_d(iter), _d(learn_rate), _d(beta1), _d(beta2), _d(eps), _p(m), _p(v),
                                       _p(g), _p(para), para.size))


def Adam_cpp(iter, learn_rate, beta1, beta2, eps, m, v, grad, para):
    g = _f(grad).ravel()
    return bool(lib().ace_oracle_Adam(_d(iter), _d(learn_rate), _d(beta1), _d(beta2), _d(eps), _p(m), _p(v),
                                      _p(g), _p(para), para.size))


def Nesterov_cpp(learn_rate, momentum, nu, grad, para):
    g = _f(grad).ravel()
    return bool(lib().ace_oracle_Nesterov(_d(learn_rate), _d(momentum), _p(nu), _p(g), _p(para), para.size))


# --------------------------------------------------------------------------- posterior
def pred_cpp(y_X, sigma, mu, invK_XX, K_xX, K_xx, mean_y, std_y):
    y_X, invK_XX, K_xX, K_xx = _f(y_X).ravel(), _f(invK_XX), _f(K_xX), _f(K_xx)
    nx, nX = K_xX.shape
    m, ci, var = np.empty(nx), np.empty((nx, 2), order="F"), np.empty(nx)
    lib().ace_oracle_pred(_p(y_X), _d(sigma), _d(mu), _p(invK_XX), _p(K_xX), _p(K_xx), _d(mean_y), _d(std_y),
                          nx, nX, _p(m), _p(ci), _p(var))
    return {"map": m, "ci": ci, "var": var}


def pred_marginal_cpp(y_X, Z_x, sigma, mu, invK_XX, K_xX, K_xx, mean_y, std_y, std_Z, calculate_ate):
    y_X, Z_x, invK_XX = _f(y_X).ravel(), _f(Z_x).ravel(), _f(invK_XX)
    K_xX, K_xx = _f(K_xX), _f(K_xx)
    nx, nX, B = K_xX.shape
    m, ci, var = np.empty(nx), np.empty((nx, 2), order="F"), np.empty(nx)
    avg = np.zeros(12)
    lib().ace_oracle_pred_marginal(_p(y_X), _p(Z_x), _d(sigma), _d(mu), _p(invK_XX), _p(K_xX), _p(K_xx),
                                   _d(mean_y), _d(std_y), _d(std_Z), int(bool(calculate_ate)), nx, nX, B,
                                   _p(m), _p(ci), _p(var), _p(avg))
    out = {"map": m, "ci": ci, "var": var}
    if calculate_ate:
        for k, name in enumerate(("ate", "att", "atu")):
            out[name] = {"map": avg[4 * k], "ci": avg[4 * k + 1:4 * k + 3].copy(), "var": avg[4 * k + 3]}
    return out


# --------------------------------------------------------------------------- ncs basis
def ncs_basis(x, knots):
    x, knots = _f(x).ravel(), _f(knots).ravel()
    K = lib().ace_oracle_ncs_ncol(_p(knots), knots.size)
    out = np.empty((x.size, K), order="F")
    lib().ace_oracle_ncs_basis(_p(x), x.size, _p(knots), knots.size, _p(out))
    return out


def ncs_basis_deriv(x, knots):
    x, knots = _f(x).ravel(), _f(knots).ravel()
    K = lib().ace_oracle_ncs_ncol(_p(knots), knots.size)
    out = np.empty((x.size, K), order="F")
    lib().ace_oracle_ncs_basis_deriv(_p(x), x.size, _p(knots), knots.size, _p(out))
    return out


def normalize_train(y, X, Z):
    """In place on y, X, Z (like the reference, src/utilities_cpp.cpp:13-104); returns the moments matrix."""
    n, px = X.shape
    pz = Z.shape[1]
    assert y.flags.f_contiguous and X.flags.f_contiguous and Z.flags.f_contiguous
    mo = np.zeros((1 + px + pz, 3), order="F")
    lib().ace_oracle_normalize_train(_p(y), _p(X), _p(Z), n, px, pz, _p(mo))
    return mo


def normalize_test(X, Z, moments):
    n, px = X.shape
    pz = Z.shape[1]
    assert X.flags.f_contiguous and Z.flags.f_contiguous
    lib().ace_oracle_normalize_test(_p(X), _p(Z), n, px, pz, _p(_f(moments)))


# --------------------------------------------------------------------------- R-level sequencing
KERNELS = {"SE": 0, "Matern32": 1}
OPTIMIZERS = {"Nadam": 0, "Adam": 1, "GD": 2, "NAG": 2}


class OracleFit:
    """State of one Kernel R6 object + optimiser (R/kernel_SE_R6.R, R/optimizer_classes.R)."""

    def __init__(self, y, X, Z, parameters, kernel="SE", optimizer="Nadam", lr=0.01, beta1=0.9, beta2=0.999,
                 momentum=0.0, norm_clip=None, clip_at=1.0, std_y=1.0, use_chol=False):
        self.y, self.X, self.Z = _f(y).ravel(), _f(X, 2), _f(Z, 2)
        self.n, self.p = self.X.shape
        self.Bz = self.Z.shape[1]
        self.B = self.Bz + 1
        self.P = 2 + self.B + self.B * self.p
        self.par = np.array(parameters, dtype=np.float64).ravel().copy()
        assert self.par.size == self.P
        self.kernel, self.optimizer = KERNELS[kernel], OPTIMIZERS[optimizer]
        if optimizer == "GD":
            momentum = 0.0  # R/utilities.R:16
        if norm_clip is None:
            norm_clip = optimizer in ("Adam", "Nadam")  # R/main_ace.R:143
        self.lr, self.b1, self.b2, self.mom = lr, beta1, beta2, momentum
        self.clip, self.clip_at, self.std_y, self.use_chol = bool(norm_clip), clip_at, std_y, bool(use_chol)
        self.m, self.v = np.zeros(self.P), np.zeros(self.P)
        n = self.n
        self.Kfull = np.empty((n, n), order="F")
        self.Kcube = np.empty((n, n, self.B), order="F")
        self.invK = np.empty((n, n), order="F")
        self.eig = np.empty(n)
        self.grad = np.zeros(self.P)
        self.tsec = np.zeros(4)

    def para_update(self, it):
        st = np.zeros(2)
        ok = lib().ace_oracle_para_update(
            int(it), _p(self.y), _p(self.X), _p(self.Z), self.n, self.p, self.Bz, self.kernel, self.optimizer,
            int(self.use_chol), _d(self.lr), _d(self.b1), _d(self.b2), _d(self.mom), int(self.clip),
            _d(self.clip_at), _d(self.std_y), _p(self.par), _p(self.m), _p(self.v), _p(self.Kfull),
            _p(self.Kcube), _p(self.invK), _p(self.eig), _p(self.grad), _p(st), _p(self.tsec))
        if not ok:
            raise FloatingPointError("Some gradients are not finite, NaN, or NA.")
        return st

    def train(self, maxiter=1000, tol=1e-4):
        stats = np.zeros((2, maxiter + 2), order="F")
        it = lib().ace_oracle_train(
            _p(self.y), _p(self.X), _p(self.Z), self.n, self.p, self.Bz, self.kernel, self.optimizer,
            int(self.use_chol), int(maxiter), _d(tol), _d(self.lr), _d(self.b1), _d(self.b2), _d(self.mom),
            int(self.clip), _d(self.clip_at), _d(self.std_y), _p(self.par), _p(self.m), _p(self.v),
            _p(self.invK), _p(stats))
        if it < 0:
            raise FloatingPointError(f"Some gradients are not finite at iteration {-it}")
        return it, stats[:, 2:it + 2].copy()
