"""Times the kernel-build and gradient phases of a C3 iteration with the library named by ACE_B200_LIB (tuning
experiments: csrc/Makefile EXTRA / OUT)."""
import os, sys
sys.path.insert(0, '/root/repo')
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit
prob = synth.make_problem("C3", n=16384)
with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
    b, g = [], []
    for it in range(1, 5):
        st, gn = f.para_update(it)
        b.append(f.last_timing_ms["build"]); g.append(f.last_timing_ms["grad"])
print(os.environ.get("ACE_B200_LIB", "default").split("/")[-1], "build", round(min(b), 3), "grad", round(min(g), 3), "evidence", st[1], flush=True)
