"""Pins the CPU oracle (oracle/ace_oracle.cpp), the host-side preprocessing of the product and the golden fixtures
against the REFERENCE'S OWN sources: oracle/_ref/libace_ref.so is /root/reference/src/*.cpp compiled unmodified
against a stand-in RcppArmadillo header (oracle/miniarma/).  In the build container the library is (re)built from
the reference; elsewhere the prebuilt library is used if it travelled, and the committed reference-generated
fixture tests/golden/ref_golden.npz is checked in any case.

Observed agreement restatement vs compiled reference: kernel builds, inverse, eigenvalues, posterior, optimisers,
clip, spline bases: bit for bit; quantities that pass through a long sum (gradient traces, mu, marginal means):
<= 1e-13 relative (Armadillo-style two-accumulator sums vs. the restatement's single running sum)."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")
have_ref = oracle.reference_available()
needs_ref = pytest.mark.skipif(not have_ref, reason="oracle/_ref/libace_ref.so not available on this machine")


def _problem(n, p, Bz, seed, binary=False):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.uniform(-1, 1, (n, p)))
    if binary:
        Z = np.asfortranarray((rng.random((n, Bz)) < 0.3).astype(float))
    else:
        Z = rng.uniform(-1, 1, (n, Bz))
        Z[rng.random((n, Bz)) < 0.15] = 0.0
        Z = np.asfortranarray(Z)
    y = rng.standard_normal(n)
    B = Bz + 1
    par = np.concatenate([[np.log(0.3), 0.1], rng.normal(0, 0.3, B), np.log(20) + rng.normal(-1.0, 0.5, B * p)])
    return y, X, Z, par


def _both(fn, *a, **k):
    r1 = fn(*a, **k)
    with oracle.using_reference():
        r2 = fn(*a, **k)
    return r1, r2


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@needs_ref
def test_reference_library_is_the_reference():
    assert oracle.ref_lib().ace_oracle_is_reference() == 1
    if os.path.isdir(oracle.REFERENCE_SRC):  # built from the sources where they lie, never from a copy
        mk = open(os.path.join(os.path.dirname(oracle.__file__), "Makefile")).read()
        assert "REF ?= /root/reference/src" in mk


@needs_ref
@pytest.mark.parametrize("kind", ["SE", "Matern32"])
@pytest.mark.parametrize("n,p,Bz,binary", [(150, 3, 4, False), (97, 10, 7, False), (64, 20, 11, False), (120, 2, 1, True)])
def test_port_equals_reference(kind, n, p, Bz, binary):
    y, X, Z, par = _problem(n, p, Bz, 7 * n + p, binary)
    _, X2, Z2, _ = _problem(31, p, Bz, 5 * n + p, binary)
    B = Bz + 1
    sym = getattr(oracle, f"kernmat_{kind}_symmetric_cpp")
    rect = getattr(oracle, f"kernmat_{kind}_cpp")
    a, b = _both(sym, X, Z, par)
    assert np.array_equal(a["full"], b["full"]) and np.array_equal(a["elements"], b["elements"])
    c, d = _both(rect, X2, X, Z2, Z, par)
    assert np.array_equal(c["full"], d["full"]) and np.array_equal(c["elements"], d["elements"])
    i1, i2 = _both(oracle.invkernel_cpp, a["full"], par[0])
    assert np.array_equal(i1["inv"], i2["inv"]) and np.array_equal(i1["eigenval"], i2["eigenval"])
    gfn = oracle.grad_SE_cpp if kind == "SE" else oracle.grad_Matern_cpp
    s1, s2 = np.zeros(2), np.zeros(2)
    g1 = gfn(y, X, Z, a["full"], a["elements"], i1["inv"], i1["eigenval"], par, s1, B, 1.7)
    with oracle.using_reference():
        g2 = gfn(y, X, Z, a["full"], a["elements"], i1["inv"], i1["eigenval"], par, s2, B, 1.7)
    assert _rel(g1, g2) <= 1e-13 and _rel(s1, s2) <= 1e-14
    if kind == "Matern32":
        assert g2[1] == 0.0                       # mu gradient forced to 0 (src/kernel_Matern_cpp.cpp:458)
    t1, t2 = _both(oracle.stats_cpp, y, a["full"], i1["inv"], i1["eigenval"], par[1], 1.3)
    assert _rel(t1, t2) <= 1e-14
    m1, m2 = _both(oracle.mu_solution_cpp, y, i1["inv"])
    assert abs(m1 - m2) <= 1e-13 * abs(m2)
    kxx = sym(X2, Z2, par)
    p1, p2 = _both(oracle.pred_cpp, y, par[0], par[1], i1["inv"], c["full"], kxx["full"], 0.4, 1.3)
    for k in ("map", "ci", "var"):
        assert _rel(p1[k], p2[k]) <= 1e-14
    zx = Z2[:, 0].copy()
    q1, q2 = _both(oracle.pred_marginal_cpp, y, zx, par[0], par[1], i1["inv"], c["elements"], kxx["elements"], 0.4, 1.3,
                   0.9, binary)
    for k in ("map", "ci", "var"):
        assert _rel(q1[k], q2[k]) <= 1e-13
    if binary:
        for k in ("ate", "att", "atu"):
            assert abs(q1[k]["map"] - q2[k]["map"]) <= 1e-13 * max(abs(q2[k]["map"]), 1e-3)
            assert abs(q1[k]["var"] - q2[k]["var"]) <= 1e-12 * abs(q2[k]["var"])


@needs_ref
def test_optimisers_clip_and_bases_equal_reference_bitwise():
    rng = np.random.default_rng(3)
    P = 40
    g = rng.standard_normal(P)
    for nm in ("Nadam_cpp", "Adam_cpp"):
        for it in (1, 2, 50):
            m1, v1, p1 = rng.normal(0, 0.1, P), rng.uniform(0, 0.2, P), rng.standard_normal(P)
            m2, v2, p2 = m1.copy(), v1.copy(), p1.copy()
            ok1 = getattr(oracle, nm)(it, 0.01, 0.9, 0.999, 1e-8, m1, v1, g, p1)
            with oracle.using_reference():
                ok2 = getattr(oracle, nm)(it, 0.01, 0.9, 0.999, 1e-8, m2, v2, g, p2)
            assert ok1 == ok2 and np.array_equal(m1, m2) and np.array_equal(v1, v2) and np.array_equal(p1, p2)
    gn = g.copy()
    gn[3] = np.nan
    m, v, p_ = np.zeros(P), np.zeros(P), np.zeros(P)
    with oracle.using_reference():
        assert oracle.Nadam_cpp(1, 0.01, 0.9, 0.999, 1e-8, m, v, gn, p_) is False   # the reference's is_finite flag
    nu1, p1 = rng.normal(0, 0.1, P), rng.standard_normal(P)
    nu2, p2 = nu1.copy(), p1.copy()
    oracle.Nesterov_cpp(0.01, 0.5, nu1, g, p1)
    with oracle.using_reference():
        oracle.Nesterov_cpp(0.01, 0.5, nu2, g, p2)
    assert np.array_equal(nu1, nu2) and np.array_equal(p1, p2)
    for scale in (3.0, 0.01):
        g1, g2 = scale * g, scale * g
        oracle.norm_clip_cpp(True, g1, 1.0)
        with oracle.using_reference():
            oracle.norm_clip_cpp(True, g2, 1.0)
        assert _rel(g1, g2) <= 1e-15   # the 2-norm's summation order is the BLAS library's business (dnrm2)
    x = rng.uniform(-1, 1, 700)
    for kn in ([-1, 1], [-0.4, 0.1, 0.6, -1.0, 1.0], [0.5, -0.5, 0.0, -1, 1, 0.0]):
        kn = np.array(kn, dtype=float)
        a, b = _both(oracle.ncs_basis, x, kn)
        assert np.array_equal(a, b)
        a, b = _both(oracle.ncs_basis_deriv, x, kn)
        assert np.array_equal(a, b)


@needs_ref
def test_product_host_code_equals_reference_bitwise():
    """ncs_basis / normalize_* of the PRODUCT library (host code, csrc/host_utils.inl) against the compiled reference:
    the north star's "bit-exact on basis"."""
    from additivecausalexpansion_b200 import api

    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, 900)
    for kn in ([-1, 1], [-0.3, 0.2, -1, 1], [0.5, -0.5, 0.0, -1, 1, 0.0], [0.1, -1, 1]):
        kn = np.array(kn, dtype=float)
        with oracle.using_reference():
            rb, rd = oracle.ncs_basis(x, kn), oracle.ncs_basis_deriv(x, kn)
        assert np.array_equal(api.ncs_basis(x, kn), rb)
        assert np.array_equal(api.ncs_basis_deriv(x, kn), rd)
    n = 300
    y = rng.normal(3, 2, n)
    X = np.asfortranarray(np.column_stack([rng.normal(1, 3, n), (rng.random(n) < 0.4) * 2.0 + 1.0, rng.uniform(-5, 2, n)]))
    for Z in (np.asfortranarray(rng.normal(-1, 0.7, (n, 1))), np.asfortranarray((rng.random((n, 1)) < 0.3) * 1.0)):
        y1, X1, Z1 = y.copy(), X.copy(order="F"), Z.copy(order="F")
        y2, X2, Z2 = y.copy(), X.copy(order="F"), Z.copy(order="F")
        mo1 = api.normalize_train(y1, X1, Z1)
        with oracle.using_reference():
            mo2 = oracle.normalize_train(y2, X2, Z2)
        assert np.array_equal(mo1, mo2) and np.array_equal(y1, y2) and np.array_equal(X1, X2) and np.array_equal(Z1, Z2)
        Xt1 = np.asfortranarray(rng.normal(0, 2, (17, 3)))
        Zt1 = np.asfortranarray(rng.normal(0, 1, (17, 1)))
        Xt2, Zt2 = Xt1.copy(order="F"), Zt1.copy(order="F")
        api.normalize_test(Xt1, Zt1, mo1)
        with oracle.using_reference():
            oracle.normalize_test(Xt2, Zt2, mo2)
        assert np.array_equal(Xt1, Xt2) and np.array_equal(Zt1, Zt2)


def test_port_and_host_code_reproduce_the_reference_generated_fixture():
    """tests/golden/ref_golden.npz was written by the compiled reference (tests/golden/make_ref_golden.py); this check
    needs neither /root/reference nor oracle/_ref."""
    from additivecausalexpansion_b200 import api

    g = np.load(GOLD)
    for name, kind in (("se_a", "SE"), ("se_bin", "SE"), ("mat_a", "Matern32"), ("mat_c3", "Matern32"), ("se_c2", "SE")):
        G = {k.split("__", 1)[1]: g[k] for k in g.files if k.startswith(name + "__")}
        y, X, Z, X2, Z2, par = G["y"], G["X"], G["Z"], G["X2"], G["Z2"], G["par"]
        B = Z.shape[1] + 1
        sym = getattr(oracle, f"kernmat_{kind}_symmetric_cpp")
        rect = getattr(oracle, f"kernmat_{kind}_cpp")
        ks = sym(X, Z, par)
        assert np.array_equal(ks["full"], G["K"]) and np.array_equal(ks["elements"], G["cube"])
        iv = oracle.invkernel_cpp(ks["full"], par[0])
        assert _rel(iv["inv"], G["inv"]) <= 1e-12     # LAPACK threading may differ between machines
        assert abs(np.sum(np.log(iv["eigenval"])) - G["logdet"][0]) <= 1e-12 * abs(G["logdet"][0])
        st = np.zeros(2)
        gfn = oracle.grad_SE_cpp if kind == "SE" else oracle.grad_Matern_cpp
        gr = gfn(y, X, Z, G["K"], G["cube"], G["inv"], iv["eigenval"], par, st, B, 1.7)
        assert _rel(gr, G["grad"]) <= 1e-12 and _rel(st, G["stats"]) <= 1e-12
        kx = rect(X2, X, Z2, Z, par)
        assert np.array_equal(kx["full"], G["KxX"]) and np.array_equal(kx["elements"], G["KxX_cube"])
        pr = oracle.pred_cpp(y, par[0], par[1], G["inv"], G["KxX"], G["Kxx"], 0.4, 1.3)
        assert _rel(pr["map"], G["pred_map"]) <= 1e-13 and _rel(pr["var"], G["pred_var"]) <= 1e-12
    for nm in ("Nadam_cpp", "Adam_cpp"):
        m, v, par, gg = (a.copy() for a in g[f"opt__{nm}_in"])
        getattr(oracle, nm)(7, 0.01, 0.9, 0.999, 1e-8, m, v, gg, par)
        assert np.array_equal(np.stack([m, v, par]), g[f"opt__{nm}_out"])
        m, v, par, gg = (a.copy() for a in g[f"opt__{nm}_in"])
        getattr(api, nm)(7, 0.01, 0.9, 0.999, 1e-8, m, v, gg, par)   # product host code
        assert np.array_equal(np.stack([m, v, par]), g[f"opt__{nm}_out"])
    gc = g["opt__clip_in"].copy()
    api.norm_clip_cpp(True, gc, 1.0)
    assert np.array_equal(gc, g["opt__clip_out"])
    assert np.array_equal(api.ncs_basis(g["ncs__z"], g["ncs__knots"]), g["ncs__B"])
    assert np.array_equal(api.ncs_basis_deriv(g["ncs__z"], g["ncs__knots"]), g["ncs__dB"])
    y, X, Z = g["norm__in_y"].copy(), g["norm__in_X"].copy(order="F"), g["norm__in_Z"].copy(order="F")
    mo = api.normalize_train(y, X, Z)
    assert np.array_equal(mo, g["norm__moments"]) and np.array_equal(y, g["norm__y"])
    assert np.array_equal(X, g["norm__X"]) and np.array_equal(Z, g["norm__Z"])
    Xt, Zt = g["norm__in_Xt"].copy(order="F"), g["norm__in_Zt"].copy(order="F")
    api.normalize_test(Xt, Zt, mo)
    assert np.array_equal(Xt, g["norm__Xt"]) and np.array_equal(Zt, g["norm__Zt"])
