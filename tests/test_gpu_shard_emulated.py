"""The multi-GPU sharded fit, played by ONE process on one GPU (ace_fit_shard_emulate): every rank's share of the
kernel build, of the split triangular inverse, of the U U^T tiles and of the gradient tiles is computed in turn,
without NCCL.  Numerically this is the N-GPU path, so its parity with the unsharded fit (itself checked against the
oracle in test_gpu_path.py) is testable on the single-GPU tier.  tests/test_gpu_shard.py repeats it with real ranks.

Tolerances: log-evidence <= 1e-12 relative, gradients <= 1e-9 of the max-norm, parameters <= 1e-10 absolute,
inverse entries <= 1e-9 of the max entry."""
import os

import numpy as np
import pytest

from additivecausalexpansion_b200.fit import AceFit

pytestmark = pytest.mark.gpu


def _problem(n, p, Bz, seed):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.uniform(-1, 1, (n, p)))
    Z = rng.uniform(-1, 1, (n, Bz))
    Z[rng.random((n, Bz)) < 0.1] = 0.0
    Z = np.asfortranarray(Z)
    y = rng.standard_normal(n)
    B = Bz + 1
    par = np.concatenate([[np.log(0.3), 0.1], rng.normal(0, 0.3, B), np.log(20) + rng.normal(-1.0, 0.5, B * p)])
    return y, X, Z, par


def _env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    for k, v in kw.items():
        os.environ[k] = str(v)
    return old


def _restore(old):
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


# (n, p, Bz, kernel, world, h_min, panel-cyclic potrf): nb = ceil(n/128) blocks; levels with child size h >= h_min, h % (2 world) == 0 split
CASES = [
    (2048, 5, 3, "SE", 2, 4, 1),         # nb 16: levels 4 and 8 split, slices of 1 and 2 blocks
    (2048, 5, 3, "SE", 2, 4, 0),         # the same with the redundant Cholesky
    (2500, 20, 2, "Matern32", 2, 4, 1),  # nb 20, ragged: lone children, ragged top node (16 + 4), 5 panels
    (4000, 8, 4, "Matern32", 4, 8, 1),   # nb 32 (n_pad 4096): world 4, levels 8 and 16
    (3072, 3, 1, "SE", 3, 6, 0),         # nb 24, world 3: no level has h % 6 == 0 -> redundant inverse, split U U^T
    (1536, 4, 2, "SE", 2, 1 << 20, 1),   # no level split at all: U U^T / gradient tile ownership, trmv, panel potrf
]


@pytest.mark.parametrize("incr", [1, 0])
@pytest.mark.parametrize("n,p,Bz,kernel,world,hmin,spotrf", CASES)
def test_emulated_sharded_fit_tracks_unsharded(n, p, Bz, kernel, world, hmin, spotrf, incr):
    """incr 1: the inverse grows behind the Cholesky panels (column panels per rank, one gather at the end);
    incr 0: split merge tree after the factorisation."""
    if incr == 0 and spotrf == 0 and hmin > 64:
        pytest.skip("same path as incr=1 for this case")
    y, X, Z, par = _problem(n, p, Bz, 11 + n)
    old = _env(ACE_SHARD_HMIN=hmin, ACE_SHARD_DENSE=1, ACE_SHARD_POTRF=spotrf, ACE_SHARD_INCR=incr)
    try:
        with AceFit(y, X, Z, par, kernel=kernel, use_graph=False) as s, \
                AceFit(y, X, Z, par, kernel=kernel, use_graph=False) as g:
            s.shard_emulate(world)
            for it in range(1, 4):
                st_s, gn_s = s.para_update(it)
                st_g, gn_g = g.para_update(it)
                assert abs(st_s[1] - st_g[1]) <= 1e-12 * abs(st_g[1])
                assert abs(st_s[0] - st_g[0]) <= 1e-9 * abs(st_g[0])
                gs, gg = s.gradients, g.gradients
                assert np.abs(gs - gg).max() <= 1e-9 * np.abs(gg).max()
                assert np.abs(s.parameters - g.parameters).max() <= 1e-10
                s.parameters = g.parameters  # re-synchronise: compare each step on identical inputs
            Ks, Kg = s.invKmatn, g.invKmatn
            assert np.abs(Ks - Kg).max() <= 1e-9 * np.abs(Kg).max()
            assert np.abs(Ks - Ks.T).max() == 0.0
            ts, tg = s.get_train_stats(), g.get_train_stats()
            assert abs(ts[1] - tg[1]) <= 1e-12 * abs(tg[1]) and abs(ts[0] - tg[0]) <= 1e-9 * abs(tg[0])
    finally:
        _restore(old)


def test_emulated_redundant_dense_path():
    """ACE_SHARD_DENSE=0: build blocks + every world-th gradient tile only, dense phases as on one GPU."""
    y, X, Z, par = _problem(1024, 6, 2, 5)
    old = _env(ACE_SHARD_DENSE=0)
    try:
        with AceFit(y, X, Z, par, kernel="SE", use_graph=False) as s, AceFit(y, X, Z, par, kernel="SE", use_graph=False) as g:
            s.shard_emulate(2)
            for it in range(1, 3):
                st_s, _ = s.para_update(it)
                st_g, _ = g.para_update(it)
                assert abs(st_s[1] - st_g[1]) <= 1e-12 * abs(st_g[1])
                assert np.abs(s.gradients - g.gradients).max() <= 1e-9 * np.abs(g.gradients).max()
                assert np.abs(s.parameters - g.parameters).max() <= 1e-10
    finally:
        _restore(old)
