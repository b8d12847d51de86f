"""Synthetic inputs of the shapes BASELINE.json names: identical bytes are fed to
the CUDA path and to the CPU oracle.  NumPy `default_rng(20260000 + config)`; X ~ U(-1,1) (the range
normalize_train leaves, src/utilities_cpp.cpp:72-95); z ~ N(0,1) median-centred and max-abs scaled (or
Bernoulli(0.3) for the binary variant); y = m(x) + sum_l g_l(x) b_l(z) + eps, standardised with the n-1 sd.
theta_0 follows R/parameters.R:4-19.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .basis import set_basis

CONFIGS = {
    # name: n, p, kernel, basis, n_knots, bspline order m (None = n/a)
    "C1": dict(n=300, p=2, kernel="SE", basis="cubic", n_knots=2),        # README example shape
    "C2": dict(n=4096, p=10, kernel="SE", basis="ncs", n_knots=5),
    "C3": dict(n=16384, p=20, kernel="Matern32", basis="B", n_knots=8),   # headline metric
    "C3w": dict(n=16384, p=20, kernel="Matern32", basis="B", n_knots=8, m=1),  # B-spline as wired (Q9): degree 0
    "C4": dict(n=65536, p=8, kernel="SE", basis="cubic", n_knots=1),
    "C5": dict(n=8192, p=10, kernel="SE", basis="linear", n_knots=1),
}
SEEDS = {"C1": 1, "C2": 2, "C3": 3, "C3w": 3, "C4": 4, "C5": 5}


@dataclass
class Problem:
    name: str
    y: np.ndarray        # n, standardised
    X: np.ndarray        # n x p, Fortran
    z: np.ndarray        # n, normalised treatment
    Z: np.ndarray        # n x Bz basis matrix (Basis$B)
    basis: object
    kernel: str
    parameters: np.ndarray  # theta_0, P
    mean_y: float
    std_y: float

    @property
    def n(self):
        return self.X.shape[0]

    @property
    def p(self):
        return self.X.shape[1]

    @property
    def B(self):
        return self.Z.shape[1] + 1


def initial_parameters(p, B, y, X, z, init_length_scale=20.0):
    """R/parameters.R:1-23.  init.sigma is always overwritten by the OLS residual variance (quirk Q7):
    log( y'(I - QQ')y / (n-1) ) with Q an orthonormal basis of [X z 1]."""
    n = y.size
    A = np.column_stack([X, z, np.ones(n)])
    Q, R = np.linalg.qr(A)
    rank = int(np.sum(np.abs(np.diag(R)) > 1e-7 * np.abs(np.diag(R)).max()))  # qr.default tol = 1e-07
    Q = Q[:, :rank]
    r = y - Q @ (Q.T @ y)
    sigma0 = np.log(float(y @ r) / (n - 1))
    return np.concatenate([[sigma0, 0.0], -np.log(np.ones(B)), np.log(np.full(B * p, init_length_scale))])


def make_problem(name="C3", n=None, p=None, seed_offset=0, binary_z=False, kernel=None, basis=None, n_knots=None):
    cfg = dict(CONFIGS[name])
    if n is not None:
        cfg["n"] = int(n)
    if p is not None:
        cfg["p"] = int(p)
    if kernel is not None:
        cfg["kernel"] = kernel
    if basis is not None:
        cfg["basis"] = basis
    if n_knots is not None:
        cfg["n_knots"] = n_knots
    n, p = cfg["n"], cfg["p"]
    rng = np.random.default_rng(20260000 + SEEDS[name] + 1000 * seed_offset)
    X = np.asfortranarray(rng.uniform(-1.0, 1.0, size=(n, p)))
    if binary_z:
        z = (rng.random(n) < 0.3).astype(np.float64)
    else:
        z = rng.standard_normal(n)
        z = z - np.median(z)
        z = z / np.max(np.abs(z))
    bobj = set_basis("binary" if binary_z else cfg["basis"], True)
    if cfg["basis"] == "B" and not binary_z:
        bobj.trainbasis(z, cfg["n_knots"], m=cfg.get("m", 4))
    else:
        bobj.trainbasis(z, cfg["n_knots"])
    Z = np.asfortranarray(bobj.B)
    # smooth random nuisance and effect functions of a few columns of X
    k = min(p, 3)
    W = rng.normal(0, 1.5, size=(k, 4))
    ph = rng.uniform(0, 2 * np.pi, size=4)
    feats = np.cos(X[:, :k] @ W + ph)
    y = feats @ rng.normal(0, 1.0, size=4)
    for l in range(Z.shape[1]):
        y = y + (feats @ rng.normal(0, 0.6, size=4) + 0.5) * Z[:, l]
    y = y + rng.normal(0, 0.1, size=n)
    mean_y = float(y.mean())
    y = y - mean_y
    std_y = float(np.sqrt(np.sum((y - y.mean()) ** 2) / (n - 1)))
    y = y / std_y
    par = initial_parameters(p, Z.shape[1] + 1, y, X, z)
    return Problem(name, y, X, z, Z, bobj, cfg["kernel"], par, mean_y, std_y)


def mid_trajectory_parameters(prob, rng_seed=7):
    """A non-degenerate theta: theta_0 with moderate length-scales and spread-out scales, so that
    K + e^sigma I is well conditioned (used by parity tests alongside theta_0)."""
    rng = np.random.default_rng(rng_seed)
    p, B = prob.p, prob.B
    par = prob.parameters.copy()
    par[0] = np.log(0.2)
    par[1] = 0.05
    par[2:2 + B] = rng.normal(-0.5, 0.4, size=B)
    par[2 + B:] = np.log(20.0) + rng.normal(-1.5, 0.7, size=B * p)
    return par
