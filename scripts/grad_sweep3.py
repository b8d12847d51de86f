"""Gradient-pass generations per BASELINE shape: grad2_kernel (16 warps) vs grad3_kernel (exact-shape, lock-step math).
Prints the device time of the gemv + gradient + finalize phase, and the evidence / gradient agreement between the two."""
import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit
out = {}
shapes = (("C3", 8192), ("C2", 4096), ("C5", 8192), ("C4", 8192), ("C1", 300))
if len(sys.argv) > 1 and sys.argv[1] == "full":
    shapes = (("C3", 16384),) + shapes
for cfg, n in shapes:
    prob = synth.make_problem(cfg, n=n)
    ref = None
    for impl, cw in ((2, 0), (3, 0), (3, 1)):
        os.environ["ACE_GRAD_IMPL"] = str(impl); os.environ["ACE_GRAD2_WARPS"] = "16"; os.environ["ACE_GRAD3_CW"] = str(cw)
        with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
            ts = []
            for it in range(1, 5):
                st, gn = f.para_update(it)
                ts.append(f.last_timing_ms["grad"])
            g = np.array(f.gradients)
        if ref is None:
            ref = (g, st)
        dg = float(np.abs(g - ref[0]).max() / np.abs(ref[0]).max())
        out[f"{cfg}_n{n}_impl{impl}_cw{cw}"] = {"grad_ms": min(ts), "evidence": st[1], "rmse": st[0], "gnorm": gn, "dgrad_rel_vs_impl2": dg}
        print(cfg, n, "impl", impl, "cw", cw, "grad ms", round(min(ts), 3), "evidence", st[1], "rmse", st[0], "dgrad", dg, flush=True)
json.dump(out, open('/root/repo/gpurun_out/r02_grad_sweep3.json', 'w'), indent=1)
