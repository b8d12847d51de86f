"""Host-side plan of the multi-GPU sharded kernel build (ace_shard_plan, no GPU needed): every column block
has exactly one owner, the two blocks of a rank balance the lower-trapezoid work, bad shapes are refused."""
import ctypes as C

import pytest

from additivecausalexpansion_b200 import _lib


def _plan(n, world, rank):
    b = (C.c_int * 2)()
    w = C.c_int(0)
    st = _lib.lib().ace_shard_plan(n, world, rank, b, C.byref(w))
    return st, list(b), w.value


@pytest.mark.parametrize("n,world", [(16384, 2), (16384, 4), (16384, 8), (65536, 8), (8192, 8), (1000, 2)])
def test_blocks_cover_and_balance(n, world):
    n_pad = (n + 127) // 128 * 128
    owned, work = {}, []
    for r in range(world):
        st, blocks, w = _plan(n, world, r)
        assert st == 0 and w * 2 * world == n_pad and w % 64 == 0
        for b in blocks:
            assert b not in owned
            owned[b] = r
        # lower-trapezoid pairs of a block starting at column c0: (n_pad - c0) * w
        work.append(sum((n_pad - b * w) * w for b in blocks))
    assert sorted(owned) == list(range(2 * world))
    assert max(work) == min(work)  # perfectly balanced by the r / 2W-1-r pairing


def test_refuses_indivisible_shapes():
    st, _, _ = _plan(300, 2, 0)  # 384 / (2*2*64) is not an integer
    assert st == -4
    st, _, _ = _plan(16384, 3, 0)
    assert st == -4
    st, _, _ = _plan(16384, 2, 2)
    assert st == -1
