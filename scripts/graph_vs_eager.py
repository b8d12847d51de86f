"""Eager multi-stream launch vs. CUDA-graph replay of one iteration, per BASELINE shape (ms per iteration)."""
import sys
sys.path.insert(0, '/root/repo')
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit
for cfg, n in (("C2", None), ("C5", None), ("C1", None), ("C3", 4096), ("C3", None)):
    prob = synth.make_problem(cfg, n=n)
    res = {}
    for graph in (False, True):
        with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=graph) as f:
            for it in range(1, 5):
                f.para_update(it)
            f.timer_start()
            K = 20 if prob.n <= 8192 else 5
            for it in range(5, 5 + K):
                st, _ = f.para_update(it)
            res[graph] = (f.timer_stop() / K, st[1])
    print(cfg, prob.n, "eager %.3f ms  graph %.3f ms" % (res[False][0], res[True][0]), "evidence equal:", res[False][1] == res[True][1], flush=True)
