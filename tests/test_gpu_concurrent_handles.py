"""Two fit handles on ONE device stepping concurrently from two host threads.  The exact-shape pair kernels keep
lambda_b and the length-scale weights of the launch in flight in a per-module __constant__ table; launches of
different handles that use it are ordered per device (csrc/ace_b200.cu: with_const_chain).  Both fits have the same
shape and kernel (so they share the kernels and the table) but different data and parameters: if the ordering failed,
one handle's gradient pass would read the other's weights.  Each must reproduce its own sequential trajectory bit for
bit."""
import threading

import numpy as np
import pytest

from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit

pytestmark = pytest.mark.gpu


def _run(prob, par, iters, out, key, barrier=None):
    with AceFit(prob.y, prob.X, prob.Z, par, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
        if barrier is not None:
            barrier.wait()
        ev = []
        for it in range(1, iters + 1):
            st, _ = f.para_update(it)
            ev.append(st[1])
        out[key] = (np.array(ev), np.array(f.parameters), np.array(f.gradients))


@pytest.mark.parametrize("cfg", ["C3", "C2"])
def test_two_handles_one_device(cfg):
    n, iters = 1536, 12
    pa = synth.make_problem(cfg, n=n)
    pb = synth.make_problem(cfg, n=n, seed_offset=777)
    par_a = np.array(pa.parameters)
    par_b = np.array(pb.parameters) + 0.3 * np.sin(np.arange(pb.parameters.size))  # different weights for sure
    seq, con = {}, {}
    _run(pa, par_a, iters, seq, "a")
    _run(pb, par_b, iters, seq, "b")
    bar = threading.Barrier(2)
    ta = threading.Thread(target=_run, args=(pa, par_a, iters, con, "a", bar))
    tb = threading.Thread(target=_run, args=(pb, par_b, iters, con, "b", bar))
    ta.start(); tb.start(); ta.join(); tb.join()
    for k in ("a", "b"):
        for i in range(3):
            assert np.array_equal(seq[k][i], con[k][i]), (cfg, k, i)
    assert not np.array_equal(seq["a"][0], seq["b"][0])
