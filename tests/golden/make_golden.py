"""Generates tests/golden/hotpath_golden.npz.

The reference (R package, needs R + Rcpp + RcppArmadillo) cannot run in the build container and ships no
golden vectors of its own, so these fixtures come from the CPU oracle
(oracle/ace_oracle.cpp, the literal restatement) and are cross-checked here against the independent
NumPy restatement (oracle/np_oracle.py) before being written.  PARITY UNPINNED against the real package.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402


def case(kind, n, p, Bz, seed):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.uniform(-1, 1, (n, p)))
    Z = rng.uniform(-1, 1, (n, Bz))
    Z[rng.random((n, Bz)) < 0.2] = 0.0
    Z = np.asfortranarray(Z)
    y = rng.standard_normal(n)
    B = Bz + 1
    par = np.concatenate([[np.log(0.3), 0.1], rng.normal(0, 0.3, B), np.log(20) + rng.normal(-1.0, 0.5, B * p)])
    sym = oracle.kernmat_SE_symmetric_cpp if kind == "SE" else oracle.kernmat_Matern32_symmetric_cpp
    grad = oracle.grad_SE_cpp if kind == "SE" else oracle.grad_Matern_cpp
    ks = sym(X, Z, par)
    iv = oracle.invkernel_cpp(ks["full"], par[0])
    st = np.zeros(2)
    g = grad(y, X, Z, ks["full"], ks["elements"], iv["inv"], iv["eigenval"], par, st, B, 1.5)
    # cross-check with the second restatement
    f2, el2 = npo.kernmat(kind, X, X, Z, Z, par)
    lam2, inv2 = npo.invkernel(f2, par[0])
    g2, st2 = npo.grad(kind, y, X, f2, el2, inv2, lam2, par, B, 1.5)
    assert np.abs(f2 - ks["full"]).max() < 1e-13
    assert np.abs(g2 - g).max() <= 1e-10 * np.abs(g).max()
    assert abs(st2[1] - st[1]) <= 1e-12 * abs(st[1])
    mu = oracle.mu_solution_cpp(y, iv["inv"])
    nx = 23
    X2 = np.asfortranarray(rng.uniform(-1, 1, (nx, p)))
    Z2 = np.asfortranarray(rng.uniform(-1, 1, (nx, Bz)))
    cross = oracle.kernmat_SE_cpp if kind == "SE" else oracle.kernmat_Matern32_cpp
    kx = cross(X2, X, Z2, Z, par)
    kxx = sym(X2, Z2, par)
    pr = oracle.pred_cpp(y, par[0], par[1], iv["inv"], kx["full"], kxx["full"], 0.3, 1.5)
    return {"X": X, "Z": Z, "y": y, "par": par, "K": ks["full"], "logdet": np.sum(np.log(iv["eigenval"])),
            "grad": g, "stats": st, "mu": mu, "X2": X2, "Z2": Z2, "pred_map": pr["map"], "pred_var": pr["var"]}


def main():
    out = {}
    for name, args in {"se": ("SE", 96, 3, 4, 11), "matern": ("Matern32", 80, 5, 2, 12)}.items():
        for k, v in case(*args).items():
            out[f"{name}_{k}"] = v
    # optimiser / clip known answers
    P = 9
    rng = np.random.default_rng(5)
    g = rng.standard_normal(P) * 3
    gc = g.copy()
    oracle.norm_clip_cpp(True, gc, 1.0)
    m, v, par = np.zeros(P), np.zeros(P), rng.standard_normal(P)
    par0 = par.copy()
    oracle.Nadam_cpp(3, 0.01, 0.9, 0.999, 1e-8, m, v, gc, par)
    out.update({"opt_g": g, "opt_gclip": gc, "opt_par0": par0, "opt_par1": par, "opt_m": m, "opt_v": v})
    z = rng.uniform(-1, 1, 40)
    kn = np.array([-0.4, 0.1, 0.6, -1.0, 1.0])
    out.update({"ncs_z": z, "ncs_knots": kn, "ncs_B": oracle.ncs_basis(z, kn), "ncs_dB": oracle.ncs_basis_deriv(z, kn)})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
