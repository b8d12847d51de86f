"""Host-side logic of the product against the oracle: O(n) preprocessing routines (bit-exact basis and knot
indices), the O(P) optimiser / clip entry points, bases, initial parameters."""
import numpy as np

import oracle
from additivecausalexpansion_b200 import api, basis, synth
from additivecausalexpansion_b200 import kernel as K


def test_ncs_basis_bit_exact_vs_oracle():
    rng = np.random.default_rng(0)
    for n, kn in [(1, [-1, 1]), (50, [-0.3, 0.2, -1, 1]), (1000, [0.5, -0.5, 0.0, -1, 1, 0.0]),
                  (333, [0.1, -1, 1])]:
        z = rng.uniform(-1, 1, n)
        kn = np.array(kn, dtype=float)
        assert np.array_equal(api.ncs_basis(z, kn), oracle.ncs_basis(z, kn))
        assert np.array_equal(api.ncs_basis_deriv(z, kn), oracle.ncs_basis_deriv(z, kn))
    # golden fixture
    g = np.load(__file__.replace("test_host.py", "golden/hotpath_golden.npz"))
    assert np.array_equal(api.ncs_basis(g["ncs_z"], g["ncs_knots"]), g["ncs_B"])
    assert np.array_equal(api.ncs_basis_deriv(g["ncs_z"], g["ncs_knots"]), g["ncs_dB"])


def test_knots_and_interval_indices_bit_exact():
    rng = np.random.default_rng(1)
    z = rng.standard_normal(501)
    z = (z - np.median(z)) / np.max(np.abs(z - np.median(z)))
    for nk in (1, 2, 5, 8):
        probs = np.arange(1, nk + 1) / (nk + 1)
        q = basis.quantile7(z, probs)
        # R type 7 written out independently: sorted x, h = (n-1)p
        xs = np.sort(z)
        h = (z.size - 1) * probs
        lo = np.floor(h).astype(int)
        ref = np.where(h > lo, (1 - (h - lo)) * xs[lo] + (h - lo) * xs[np.minimum(lo + 1, z.size - 1)], xs[lo])
        assert np.array_equal(q, ref)
        knots = np.concatenate([[-1.0], q, [1.0]])
        idx = basis.knot_interval_index(z, knots)
        ref_idx = np.array([max(0, min(len(knots) - 2, int(np.sum(knots <= v)) - 1)) for v in z])
        assert np.array_equal(idx, ref_idx)


def test_bspline_partition_of_unity_and_derivative():
    x = np.linspace(-1, 1, 201)
    kn = [-0.6, -0.1, 0.3, 0.7]
    for deg in (0, 1, 3):
        Bm = basis.bspline_design(x, kn, degree=deg)
        assert Bm.shape == (201, len(kn) + deg)
        # with the dropped intercept function added back the basis sums to one
        full = basis.bspline_design(x, kn, degree=deg)
        first = 1.0 - full.sum(axis=1)
        assert np.all(first > -1e-12) and np.all(first < 1 + 1e-12)
        if deg >= 1:
            dB = basis.bspline_design(x, kn, degree=deg, deriv=1)
            h = 1e-6
            xm = np.clip(x[5:-5], -1 + 2 * h, 1 - 2 * h)
            fd = (basis.bspline_design(xm + h, kn, degree=deg) - basis.bspline_design(xm - h, kn, degree=deg)) / (2 * h)
            ok = np.ones(xm.size, bool)
            for k in kn:  # skip points within h of a knot (kinks at degree 1)
                ok &= np.abs(xm - k) > 10 * h
            assert np.abs(fd[ok] - basis.bspline_design(xm, kn, degree=deg, deriv=1)[ok]).max() < 1e-5
            assert dB.shape == Bm.shape


def test_clip_and_optimisers_match_oracle_and_golden():
    g = np.load(__file__.replace("test_host.py", "golden/hotpath_golden.npz"))
    gc = g["opt_g"].copy()
    api.norm_clip_cpp(True, gc, 1.0)
    assert np.array_equal(gc, g["opt_gclip"])
    P = gc.size
    m, v, par = np.zeros(P), np.zeros(P), g["opt_par0"].copy()
    assert api.Nadam_cpp(3, 0.01, 0.9, 0.999, 1e-8, m, v, gc, par)
    assert np.array_equal(par, g["opt_par1"]) and np.array_equal(m, g["opt_m"]) and np.array_equal(v, g["opt_v"])
    rng = np.random.default_rng(2)
    for name, fa, fo in (("adam", api.Adam_cpp, oracle.Adam_cpp), ("nadam", api.Nadam_cpp, oracle.Nadam_cpp)):
        grad = rng.standard_normal(P)
        a = [np.full(P, 0.1), np.full(P, 0.2), rng.standard_normal(P)]
        b = [x.copy() for x in a]
        assert fa(7, 0.02, 0.8, 0.99, 1e-8, a[0], a[1], grad, a[2]) == fo(7, 0.02, 0.8, 0.99, 1e-8, b[0], b[1], grad, b[2])
        for x, yv in zip(a, b):
            assert np.array_equal(x, yv), name
    nu_a, nu_b, pa, pb = np.full(P, 0.3), np.full(P, 0.3), np.ones(P), np.ones(P)
    grad = rng.standard_normal(P)
    api.Nesterov_cpp(0.01, 0.5, nu_a, grad, pa)
    oracle.Nesterov_cpp(0.01, 0.5, nu_b, grad, pb)
    assert np.array_equal(nu_a, nu_b) and np.array_equal(pa, pb)
    bad = grad.copy()
    bad[3] = np.nan
    assert api.Nadam_cpp(1, 0.01, 0.9, 0.999, 1e-8, np.zeros(P), np.zeros(P), bad, np.zeros(P)) is False
    # clip leaves a short vector alone and never touches a non-finite one
    short = np.full(P, 0.01)
    api.norm_clip_cpp(True, short, 1.0)
    assert np.array_equal(short, np.full(P, 0.01))
    api.norm_clip_cpp(True, bad, 1.0)
    assert np.isnan(bad[3]) and bad[0] == grad[0]


def test_optimizer_classes_clip_then_step_and_stop():
    class _K:
        parameters = np.zeros(5)

    opt = K.set_optimizer("Nadam", _K, 0.01, 0.0, 0.9, 0.999, True, 1.0)
    par, grad = np.ones(5), np.array([3.0, 4.0, 0, 0, 0])
    opt.update(1, par, grad)
    assert abs(np.linalg.norm(grad) - 1.0) < 1e-15  # clipped in place to UNIT norm (quirk Q5)
    import pytest

    with pytest.raises(FloatingPointError):
        opt.update(2, par, np.array([np.nan, 0, 0, 0, 0]))
    assert isinstance(K.set_optimizer("GD", _K, 0.01, 0.7, 0.9, 0.999, False, 1.0), K.optNesterov)
    assert K.set_optimizer("GD", _K, 0.01, 0.7, 0.9, 0.999, False, 1.0).momentum == 0.0


def test_normalize_train_test_roundtrip():
    rng = np.random.default_rng(3)
    n = 200
    y = np.asfortranarray(rng.normal(3.0, 2.0, n))
    X = np.asfortranarray(np.column_stack([rng.normal(1, 3, n), (rng.random(n) < 0.5) * 2.0 + 1.0, rng.uniform(5, 9, n)]))
    Z = np.asfortranarray(rng.normal(0, 1, (n, 1)))
    X0, Z0, y0 = X.copy(), Z.copy(), y.copy()
    mom = api.normalize_train(y, X, Z)
    assert abs(y.mean()) < 1e-12 and abs(y.std(ddof=1) - 1) < 1e-12
    assert abs(mom[0, 0] - y0.mean()) <= 4e-16 * abs(y0.mean())  # Armadillo sums in two interleaved accumulators
    for c in (0, 2):
        assert abs(np.median(X[:, c])) < 1e-12 and abs(np.max(np.abs(X[:, c])) - 1) < 1e-12
    assert set(np.unique(X[:, 1])) == {0.0, 1.0}  # binary column relocated to {0, 1}
    assert mom[2, 2] == 1.0
    assert abs(np.median(Z[:, 0])) < 1e-12 and abs(np.max(np.abs(Z[:, 0])) - 1) < 1e-12
    # normalize_test applies rows i+1 of the moments (for non-binary columns this reproduces the training transform)
    X2, Z2 = np.asfortranarray(X0.copy()), np.asfortranarray(Z0.copy())
    api.normalize_test(X2, Z2, mom)
    assert np.abs(X2[:, 0] - X[:, 0]).max() < 1e-12 and np.abs(X2[:, 2] - X[:, 2]).max() < 1e-12
    assert np.abs(Z2 - Z).max() < 1e-12


def test_initial_parameters_layout_and_sigma():
    prob = synth.make_problem("C1")
    p, B = prob.p, prob.B
    par = prob.parameters
    assert par.size == 2 + B + B * p
    assert par[1] == 0 and np.all(par[2:2 + B] == 0) and np.allclose(par[2 + B:], np.log(20.0))
    A = np.column_stack([prob.X, prob.z, np.ones(prob.n)])
    res = prob.y - A @ np.linalg.lstsq(A, prob.y, rcond=None)[0]
    assert abs(par[0] - np.log(res @ res / (prob.n - 1))) < 1e-10
    assert prob.Z.shape == (300, 4)  # "cubic" -> ns_spline, n.knots = 2 -> 4 knots -> 4 columns (quirk Q8)
    assert synth.make_problem("C3", n=256).Z.shape[1] == 11 and synth.make_problem("C3w", n=256).Z.shape[1] == 8
