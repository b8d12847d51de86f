"""Second, independent restatement of the reference hot path in vectorised NumPy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Purpose: cross-check
``ace_oracle.cpp`` (two restatements written separately must agree to ~1e-13) and,
with ``dtype=np.longdouble``, give an extended-precision evaluation of the same
formulas at small n so that both the C++ oracle and the CUDA path can be scored
against something more accurate than either.  PARITY UNPINNED.

Formulas are written from the mathematical description in SURVEY.md section 8a
(file:line citations there and repeated below), not transliterated from the C++.
"""
from __future__ import annotations

import numpy as np


def _weights(par, B, p, dtype):
    """w[d, b] = exp(-theta[1 + b + B*(d+1)])  -- the BUILD indexing (quirk Q1;
    src/kernel_SE_cpp.cpp:33,89; src/kernel_Matern_cpp.cpp:72,210)."""
    idx = 1 + np.arange(B)[None, :] + B * (np.arange(p)[:, None] + 1)
    return np.exp(-np.asarray(par, dtype=dtype)[idx])


def _dist(X1, X2, w):
    """D[b] = sum_d (X1[:,d,None]-X2[None,:,d])^2 * w[d,b] -> (B, n1, n2)."""
    d2 = (X1[:, None, :] - X2[None, :, :]) ** 2  # n1 n2 p
    return np.einsum("ijd,db->bij", d2, w)


def kernmat(kind, X1, X2, Z1, Z2, par, dtype=np.float64):
    """kernmat_SE_cpp / kernmat_Matern32_cpp (src/kernel_SE_cpp.cpp:9-64,
    src/kernel_Matern_cpp.cpp:52-93); the symmetric variants (:67-134, :190-240)
    give the same values with X1=X2, Z1=Z2.  Returns (full, elements[n1,n2,B])."""
    X1, X2 = np.asarray(X1, dtype=dtype), np.asarray(X2, dtype=dtype)
    Z1 = np.asarray(Z1, dtype=dtype).reshape(X1.shape[0], -1)
    Z2 = np.asarray(Z2, dtype=dtype).reshape(X2.shape[0], -1)
    par = np.asarray(par, dtype=dtype).ravel()
    p, B = X1.shape[1], Z1.shape[1] + 1
    D = _dist(X1, X2, _weights(par, B, p, dtype))
    lam = par[2:2 + B]
    if kind == "SE":
        base = np.exp(lam[:, None, None] - D)
    elif kind == "Matern32":
        r = np.sqrt(D)
        s3 = np.sqrt(dtype(3.0))
        base = (1 + s3 * r) * np.exp(lam[:, None, None] - s3 * r)
    else:
        raise ValueError(kind)
    zz = np.ones((B,) + D.shape[1:], dtype=dtype)
    zz[1:] = Z1.T[:, :, None] * Z2.T[:, None, :]
    el = base * zz  # exact zeros wherever a basis value is zero
    return el.sum(axis=0), np.moveaxis(el, 0, 2)


def spd_inverse(A, dtype=np.float64):
    """(inverse, log det) of an SPD matrix.  float64: eigendecomposition exactly as
    invkernel_cpp (src/kernel_SE_cpp.cpp:137-157).  longdouble: hand Cholesky."""
    if dtype == np.float64:
        lam, V = np.linalg.eigh(A)
        Vs = V / np.sqrt(lam)[None, :]
        return Vs @ Vs.T, lam
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        s = A[j, j] - np.dot(L[j, :j], L[j, :j])
        L[j, j] = np.sqrt(s)
        if j + 1 < n:
            L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    Li = np.zeros_like(A)
    for j in range(n):  # forward substitution column by column
        e = np.zeros(n, dtype=A.dtype)
        e[j] = 1
        x = np.zeros(n, dtype=A.dtype)
        for i in range(j, n):
            x[i] = (e[i] - np.dot(L[i, j:i], x[j:i])) / L[i, i]
        Li[:, j] = x
    return Li.T @ Li, np.diag(L) ** 2


def invkernel(K, sigma, dtype=np.float64):
    K = np.asarray(K, dtype=dtype)
    A = K + np.exp(dtype(sigma)) * np.eye(K.shape[0], dtype=dtype)
    inv, lam = spd_inverse(A, dtype)
    return lam, inv


def grad(kind, y, X, Kfull, Kel, invK, eigenval, par, B, std_y, dtype=np.float64):
    """grad_SE_cpp (src/kernel_SE_cpp.cpp:161-243) / grad_Matern_cpp
    (src/kernel_Matern_cpp.cpp:340-377,420-467).  Returns (gradients, stats)."""
    y = np.asarray(y, dtype=dtype).ravel()
    X = np.asarray(X, dtype=dtype)
    par = np.asarray(par, dtype=dtype).ravel()
    Kfull, Kel, invK = (np.asarray(a, dtype=dtype) for a in (Kfull, Kel, invK))
    n, p = X.shape
    P = 2 + B + B * p
    g = np.zeros(P, dtype=dtype)
    ybar = y - par[1]
    alpha = invK @ ybar
    W = invK - np.outer(alpha, alpha)
    g[0] = -0.5 * np.trace(W) * np.exp(par[0])
    for b in range(B):
        g[2 + b] = -0.5 * np.sum(W * Kel[:, :, b].T)
    d2 = (X[:, None, :] - X[None, :, :]) ** 2
    Lg = par[2 + B:].reshape(p, B)  # Lg[d, b] = theta[2+B + b + B*d]  (GRADIENT indexing)
    if kind == "SE":
        for d in range(p):
            for b in range(B):
                g[2 + B + b + B * d] = -0.5 * np.sum(W * (Kel[:, :, b] * d2[:, :, d]).T) * np.exp(-Lg[d, b])
        g[1] = np.sum(invK @ ybar)
    else:
        Dg = np.einsum("ijd,db->ijb", d2, np.exp(-Lg))
        Tb = Kel / (1 + np.sqrt(3 * Dg))
        for d in range(p):
            for b in range(B):
                g[2 + B + b + B * d] = -0.25 * 9 * np.sum(W * (Tb[:, :, b] * d2[:, :, d]).T) * np.exp(-Lg[d, b])
        g[1] = 0
    rmse = std_y * np.sqrt(np.sum((ybar - Kfull @ alpha) ** 2)) / np.sqrt(dtype(n))
    ev = -0.5 * (n * np.log(2 * dtype(np.pi)) + np.sum(np.log(eigenval)) + np.dot(y, alpha))  # y, not ybar (Q3)
    return g, np.array([rmse, ev], dtype=dtype)


def stats(y, Kmat, invK, eigenval, mu, std_y=1.0, dtype=np.float64):
    """stats_cpp (src/stats_cpp.cpp:9-32)."""
    y = np.asarray(y, dtype=dtype).ravel()
    ybar = y - mu
    alpha = invK @ ybar
    n = y.size
    rmse = std_y * np.sqrt(np.sum((ybar - Kmat @ alpha) ** 2)) / np.sqrt(dtype(n))
    ev = -0.5 * (n * np.log(2 * dtype(np.pi)) + np.sum(np.log(eigenval)) + np.dot(y, alpha))
    return np.array([rmse, ev], dtype=dtype)


def mu_solution(y, invK):
    """mu_solution_cpp (src/utilities_cpp.cpp:6-10)."""
    return 0.5 * np.sum(invK @ np.asarray(y).ravel()) / np.sum(invK)


def norm_clip(flag, g, max_length):
    """norm_clip_cpp (src/utilities_cpp.cpp:121-129): rescale to UNIT norm (Q5)."""
    if flag:
        L2 = np.sqrt(np.sum(g * g))
        if L2 > max_length and np.isfinite(L2) and L2 != 0:
            return g / L2
    return g


def nadam(it, lr, b1, b2, eps, m, v, g, par):
    """Nadam_cpp (src/optimizer_cpp.cpp:23-42); returns new (m, v, par, finite)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g ** 2
    par = par + lr * ((b1 * m + (1 - b1) * g) / (1 - b1 ** it)) / (np.sqrt(v / (1 - b2 ** it)) + eps)
    return m, v, par, bool(np.all(np.isfinite(g)))


def adam(it, lr, b1, b2, eps, m, v, g, par):
    """Adam_cpp (src/optimizer_cpp.cpp:45-63)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g ** 2
    par = par + lr * (m / (1 - b1 ** it)) / (np.sqrt(v / (1 - b2 ** it)) + eps)
    return m, v, par, bool(np.all(np.isfinite(g)))


def nesterov(lr, mom, nu, g, par):
    """Nesterov_cpp (src/optimizer_cpp.cpp:8-20)."""
    nu = mom * nu + lr * g
    return nu, par + nu, bool(np.all(np.isfinite(g)))


def pred(y_X, sigma, mu, invK, K_xX, K_xx, mean_y, std_y):
    """pred_cpp (src/pred_cpp.cpp:8-34)."""
    T = K_xX @ invK
    m = mean_y + std_y * (T @ (np.asarray(y_X).ravel() - mu) + mu)
    C = K_xx - T @ K_xX.T
    sd = std_y * np.sqrt(np.abs(np.diag(C) + np.exp(sigma)))
    return {"map": m, "ci": np.stack([m - 1.96 * sd, m + 1.96 * sd], axis=1), "var": sd ** 2}


def pred_marginal(y_X, Z_x, sigma, mu, invK, K_xX, K_xx, mean_y, std_y, std_Z, calculate_ate):
    """pred_marginal_cpp (src/pred_cpp.cpp:37-126)."""
    B = K_xX.shape[2]
    if B > 1:
        KmxX, Kmxx = K_xX[:, :, 1:].sum(axis=2), K_xx[:, :, 1:].sum(axis=2)
    else:
        KmxX, Kmxx = K_xX[:, :, 0], K_xx[:, :, 0]
    T = KmxX @ invK
    m = std_y * (T @ (np.asarray(y_X).ravel() - mu)) / std_Z
    Cm = Kmxx - T @ KmxX.T
    sd = std_y * np.sqrt(np.abs(np.diag(Cm))) / std_Z
    out = {"map": m, "ci": np.stack([m - 1.96 * sd, m + 1.96 * sd], axis=1), "var": sd ** 2}
    if calculate_ate:
        nx = m.size
        Z_x = np.asarray(Z_x, dtype=float).ravel()

        def pack(val, s):
            return {"map": val, "ci": np.array([val - 1.96 * s, val + 1.96 * s]), "var": s ** 2}

        ate = m.mean()
        out["ate"] = pack(ate, std_y * np.sqrt(Cm.sum()) / nx)
        nt = int(Z_x.sum())
        att = m @ Z_x / nt
        out["att"] = pack(att, std_y * np.sqrt(Z_x @ Cm @ Z_x) / nt)
        nu = nx - nt
        u = (Z_x == 0).astype(float)
        out["atu"] = pack((ate * nx - att * nt) / nu, std_y * np.sqrt(u @ Cm @ u) / nu)
    return out


def ncs_basis(x, knots, deriv=False):
    """ncs_basis / ncs_basis_deriv (src/ncs_basis_cpp.cpp:5-99): K columns for K unique knots."""
    x = np.asarray(x, dtype=np.float64).ravel()
    k = np.unique(np.asarray(knots, dtype=np.float64))
    K = k.size
    if deriv:
        tp = lambda c: 3 * (x > c) * (x - c) ** 2  # noqa: E731
    else:
        tp = lambda c: (x > c) * (x - c) ** 3  # noqa: E731
    last = tp(k[K - 1])
    d = np.stack([(tp(k[i]) - last) / (k[K - 1] - k[i]) for i in range(K - 1)], axis=1)
    out = np.empty((x.size, K))
    out[:, 0] = 1.0 if deriv else x
    for i in range(K - 2):
        out[:, 1 + i] = d[:, i] - d[:, K - 2]
    out[:, K - 1] = -d[:, K - 2]
    return out


def para_update(kind, it, y, X, Z, par, m, v, optimizer="Nadam", lr=0.01, b1=0.9, b2=0.999, mom=0.0,
                clip=True, clip_at=1.0, std_y=1.0, dtype=np.float64):
    """Kernel$para_update + Optim$update (R/kernel_SE_R6.R:40-62, R/optimizer_classes.R:54-63).
    Returns (par, m, v, stats, clipped_gradients, invK)."""
    par = np.asarray(par, dtype=dtype).copy()
    B = np.asarray(Z).reshape(len(y), -1).shape[1] + 1
    Kfull, Kel = kernmat(kind, X, X, Z, Z, par, dtype)
    lam, invK = invkernel(Kfull, par[0], dtype)
    if it == 1:
        par[1] = mu_solution(np.asarray(y, dtype=dtype), invK)
    g, st = grad(kind, y, X, Kfull, Kel, invK, lam, par, B, std_y, dtype)
    g = norm_clip(clip, g, clip_at)
    if optimizer == "Nadam":
        m, v, par, ok = nadam(it, lr, b1, b2, 1e-8, m, v, g, par)
    elif optimizer == "Adam":
        m, v, par, ok = adam(it, lr, b1, b2, 1e-8, m, v, g, par)
    else:
        m, par, ok = nesterov(lr, mom, m, g, par)
    if not ok:
        raise FloatingPointError("gradients not finite")
    par[1] = mu_solution(np.asarray(y, dtype=dtype), invK)  # stale inverse (Q6)
    return par, m, v, st, g, invK
