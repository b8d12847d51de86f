"""Small end-to-end pass of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from additivecausalexpansion_b200 import api, synth
from additivecausalexpansion_b200.fit import AceFit
for kind in ("SE", "Matern32"):
    prob = synth.make_problem("C1", kernel=kind)
    with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=kind, std_y=prob.std_y, use_graph=False) as f:
        for it in (1, 2):
            st, gn = f.para_update(it)
        ts = f.get_train_stats()
        rng = np.random.default_rng(0)
        X2 = np.asfortranarray(rng.uniform(-1, 1, (70, prob.p))); z2 = rng.uniform(-1, 1, 70); tb = prob.basis.testbasis(z2)
        pr = f.predict(X2, tb["B"], 0.0, 1.0)
        pm = f.predict_marginal(X2, tb["B"], tb["dB"], 0.0, 1.0, 1.0, True)
        print(kind, st, ts, pr["map"][:2], pm["map"][:2])
A = np.eye(300) * 2 + 0.01
r = api.dbg_spd_inverse(A)
print("inv ok", np.abs(r["inv"] @ A - np.eye(300)).max())
k = api.kernmat_SE_cpp(prob.X[:50], prob.X, prob.Z[:50], prob.Z, prob.parameters)
print("rect ok", k["full"].shape)
