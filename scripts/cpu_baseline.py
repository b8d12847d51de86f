"""CPU baseline table for BASELINE.md: the oracle port (literal dsyevd path and the CPU-favourable
dpotrf/dpotri variant) timed on the box's host cores, per-stage split, at reduced n."""
import json, os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import oracle
from additivecausalexpansion_b200 import synth
cores = os.cpu_count()
oracle.lib(cores)
out = {"cores": cores, "threads": oracle.threads(), "rows": []}
for cfg, sizes in (("C2", (1024, 2048)), ("C3", (1024, 2048)), ("C5", (2048,)), ("C1", (300,))):
    for n in sizes:
        prob = synth.make_problem(cfg, n=n)
        for chol in (False, True):
            of = oracle.OracleFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_chol=chol)
            ts, st = [], []
            for it in range(1, 3):
                t0 = time.perf_counter(); of.para_update(it); ts.append(time.perf_counter() - t0); st.append(of.tsec.copy())
            row = {"config": cfg, "n": n, "p": prob.p, "B": prob.B, "kernel": prob.kernel, "variant": "dpotrf" if chol else "dsyevd",
                   "sec_per_iter": min(ts), "stages": dict(zip(("build", "inverse", "gradient", "rest"), np.min(np.array(st), axis=0).tolist()))}
            out["rows"].append(row); print(row, flush=True)
json.dump(out, open('/root/repo/gpurun_out/cpu_baseline.json', 'w'), indent=1)
