"""Single GPU: time of the Cholesky phase of the sharded schedule played by one process (all panels owned, no
NCCL) next to the regular single-GPU look-ahead potrf -- both are complete factorizations of the same matrix."""
import sys
sys.path.insert(0, '/root/repo')
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
prob = synth.make_problem(cfg)
for world in (1, 2):
    with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
        if world > 1:
            f.shard_emulate(world)
        for it in range(1, 4):
            f.para_update(it)
        print("emulated world", world, f.last_timing_ms)
