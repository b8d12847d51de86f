"""Every exact-shape instantiation of the pair kernels (csrc/pair_grad3.cu, csrc/pair_kernmat2.cu: B = 2..16 additive
terms x ceil(p / 8) = 1..4 x kernel kind, symmetric and rectangular) against the CPU oracle, through the C ABI.
The BASELINE shapes only exercise a handful of the 256 compiled kernels; this sweep launches each of them once on a
small problem.  Tolerances as in test_gpu_path.py (K entries <= 1e-12 absolute, exact zeros; gradients, log-evidence
<= 1e-9 relative)."""
import numpy as np
import pytest

import oracle
from additivecausalexpansion_b200 import api

from test_gpu_path import KINDS, RTOL, _problem

pytestmark = pytest.mark.gpu

P_OF_TILECOUNT = {1: 5, 2: 12, 3: 17, 4: 31}   # one p per NT = ceil(p / 8), none of them a multiple of 4


@pytest.mark.parametrize("kind", ["SE", "Matern32"])
@pytest.mark.parametrize("nt", [1, 2, 3, 4])
def test_all_term_counts(kind, nt):
    p = P_OF_TILECOUNT[nt]
    n, n2 = 97, 70
    for Bz in range(1, 16):          # B = 2 .. 16
        B = Bz + 1
        y, X, Z, par = _problem(n, p, Bz, seed=100 * nt + Bz)
        _, X2, Z2, _ = _problem(n2, p, Bz, seed=7)
        # symmetric build (with the cube) and rectangular build
        g = KINDS[kind][1](X, Z, par)
        o = KINDS[kind][4](X, Z, par)
        assert np.abs(g["full"] - o["full"]).max() <= 1e-12, (kind, p, B)
        assert np.abs(g["elements"] - o["elements"]).max() <= 1e-12, (kind, p, B)
        assert np.array_equal(g["elements"] == 0.0, o["elements"] == 0.0), (kind, p, B)
        assert np.array_equal(g["full"], g["full"].T), (kind, p, B)
        gr = KINDS[kind][0](X2, X, Z2, Z, par)
        orr = KINDS[kind][3](X2, X, Z2, Z, par)
        assert np.abs(gr["full"] - orr["full"]).max() <= 1e-12, (kind, p, B)
        assert np.array_equal(gr["elements"] == 0.0, orr["elements"] == 0.0), (kind, p, B)
        # gradient pass
        iv = oracle.invkernel_cpp(o["full"], par[0])
        st_o, st_g = np.zeros(2), np.zeros(2)
        go = KINDS[kind][5](y, X, Z, o["full"], o["elements"], iv["inv"], iv["eigenval"], par, st_o, B, 1.7)
        gg = KINDS[kind][2](y, X, Z, None, None, iv["inv"], iv["eigenval"], par, st_g, B, 1.7)
        assert np.abs(gg - go).max() <= RTOL * np.abs(go).max(), (kind, p, B)
        assert abs(st_g[1] - st_o[1]) <= RTOL * abs(st_o[1]), (kind, p, B)
        assert abs(st_g[0] - st_o[0]) <= 1e-8 * abs(st_o[0]), (kind, p, B)
