"""Gradient-pass variants per BASELINE shape: grad_kernel (b-groups) vs grad2_kernel with 8 / 16 warps per CTA.
Prints the device time of the gemv + gradient + finalize phase and the evidence / gradient norm (must agree)."""
import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit
out = {}
for cfg, n in (("C3", 8192), ("C2", 4096), ("C5", 8192), ("C4", 8192)):
    prob = synth.make_problem(cfg, n=n)
    for impl, nw in ((1, 8), (2, 8), (2, 16)):
        os.environ["ACE_GRAD_IMPL"] = str(impl); os.environ["ACE_GRAD2_WARPS"] = str(nw)
        with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
            ts = []
            for it in range(1, 5):
                st, gn = f.para_update(it)
                ts.append(f.last_timing_ms["grad"])
            g = f.gradients
        out[f"{cfg}_n{n}_impl{impl}_w{nw}"] = {"grad_ms": min(ts), "evidence": st[1], "gnorm": gn, "g0": float(g[0])}
        print(cfg, n, "impl", impl, "warps", nw, "grad ms", round(min(ts), 3), "evidence", st[1], "gnorm", gn, flush=True)
json.dump(out, open('/root/repo/gpurun_out/r02_grad_sweep2.json', 'w'), indent=1)
