"""Quick device-time probe: dense phases (potrf / trtri / UU^T) and one full iteration per config."""
import json, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from additivecausalexpansion_b200 import api, synth
from additivecausalexpansion_b200.fit import AceFit

out = {"dense": {}, "iter": {}}
sizes = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [4096, 8192, 16384]
for n in sizes:
    ms = api.bench_dense(n, 2)
    fl = n ** 3 / 3
    out["dense"][n] = {"ms": ms.tolist(), "tflops": [fl / (m * 1e-3) * 1e-12 for m in ms]}
    print(n, out["dense"][n], flush=True)
cfgs = sys.argv[2].split(',') if len(sys.argv) > 2 else ["C2", "C3"]
for name in cfgs:
    t0 = time.time()
    prob = synth.make_problem(name)
    tgen = time.time() - t0
    for graph in (False, True):
        with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=graph) as f:
            res = []
            for it in range(1, 5):
                t1 = time.time()
                st, gn = f.para_update(it)
                wall = time.time() - t1
                res.append({"it": it, "stats": st.tolist(), "gnorm": gn, "wall_ms": wall * 1e3, "dev": f.last_timing_ms})
            out["iter"][f"{name}_graph{int(graph)}"] = res
            print(name, graph, json.dumps(res[-1]), flush=True)
    print(name, "gen s", tgen)
json.dump(out, open('/root/repo/gpurun_out/perf_probe.json', 'w'), indent=1)
