"""Measures the BASELINE.json configs other than the headline one (they are parity-test cases, not bench lines):
M2 = FP64 TFLOP/s inside potrf at n in {16384, 32768, 65536}; C2 iterations/s; C4 = one iteration + posterior
on 4096 points at n = 65536; C5 = iterations/s of an n = 8192 restart.  Writes gpurun_out/config_runs.json."""
import json, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from additivecausalexpansion_b200 import api, synth
from additivecausalexpansion_b200.fit import AceFit

out = {}
which = sys.argv[1].split(',') if len(sys.argv) > 1 else ["M2", "C2", "C5", "C4"]
if "M2" in which:
    out["M2_potrf"] = {}
    for n, reps in ((16384, 2), (32768, 2), (65536, 1)):
        ms = api.bench_dense(n, reps)
        fl = n ** 3 / 3
        out["M2_potrf"][n] = {"potrf_ms": ms[0], "trtri_ms": ms[1], "uut_ms": ms[2],
                              "potrf_tflops": fl / ms[0] * 1e-9, "trtri_tflops": fl / ms[1] * 1e-9, "uut_tflops": fl / ms[2] * 1e-9,
                              "potrf_frac_of_37.07": fl / ms[0] * 1e-9 / 37.07}
        print(n, out["M2_potrf"][n], flush=True)
for name in ("C2", "C5"):
    if name not in which: continue
    prob = synth.make_problem(name)
    with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
        for it in range(1, 4): f.para_update(it)
        f.timer_start()
        K = 20
        for it in range(4, 4 + K): st, gn = f.para_update(it)
        ms = f.timer_stop() / K
        out[name] = {"n": prob.n, "p": prob.p, "B": prob.B, "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "phases": f.last_timing_ms, "stats": st.tolist()}
        print(name, out[name], flush=True)
if "C4" in which:
    t0 = time.time(); prob = synth.make_problem("C4"); tg = time.time() - t0
    with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, use_graph=False) as f:
        t0 = time.time(); st, gn = f.para_update(1); w1 = time.time() - t0
        ph = f.last_timing_ms
        rng = np.random.default_rng(1)
        nx = 4096
        X2 = np.asfortranarray(rng.uniform(-1, 1, (nx, prob.p))); z2 = rng.uniform(-1, 1, nx)
        tb = prob.basis.testbasis(z2)
        t0 = time.time(); pr = f.predict(X2, tb["B"], prob.mean_y, prob.std_y); wp = time.time() - t0
        out["C4"] = {"n": prob.n, "gen_s": tg, "iter_wall_s": w1, "phases_ms": ph, "stats": st.tolist(), "predict_wall_s": wp,
                     "potrf_tflops": prob.n ** 3 / 3 / ph["potrf"] * 1e-9, "map_head": pr["map"][:3].tolist(), "var_head": pr["var"][:3].tolist(),
                     "finite": bool(np.all(np.isfinite(pr["map"])) and np.all(np.isfinite(pr["var"])))}
        print("C4", out["C4"], flush=True)
json.dump(out, open('/root/repo/gpurun_out/config_runs.json', 'w'), indent=1, default=float)
