import os, sys, json
sys.path.insert(0, '/root/repo')
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit
prob = synth.make_problem("C3")
for bt, gpc in ((3, 4), (3, 2), (2, 2), (2, 3), (2, 4), (3, 1), (2, 1)):
    os.environ["ACE_GRAD_BT"] = str(bt); os.environ["ACE_GRAD_GPC"] = str(gpc)
    with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
        for it in range(1, 4):
            st, gn = f.para_update(it)
        print(bt, gpc, f.last_timing_ms["grad"], st, flush=True)
