"""Config-5 pattern on ONE GPU: k independent fits (n = 8192) stepped concurrently from k host threads vs. one after the
other.  A single fit of this size is bound by the serial panel chain of its Cholesky (profiles/r02): a second fit's
GEMMs fill the SMs the chain leaves idle."""
import sys, threading, time
sys.path.insert(0, '/root/repo')
import numpy as np
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit

cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
iters = 20
for k in (1, 2, 3, 4):
    probs = [synth.make_problem(cfg, seed_offset=i) for i in range(k)]
    fits = [AceFit(p.y, p.X, p.Z, p.parameters, kernel=p.kernel, std_y=p.std_y, use_graph=False) for p in probs]
    for f in fits:
        for it in range(1, 4):
            f.para_update(it)
    bar = threading.Barrier(k + 1)
    ev = [None] * k
    def work(i):
        bar.wait()
        for it in range(4, 4 + iters):
            st, _ = fits[i].para_update(it)
        ev[i] = st[1]
    th = [threading.Thread(target=work, args=(i,)) for i in range(k)]
    for t in th: t.start()
    bar.wait(); t0 = time.perf_counter()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"{cfg}: {k} concurrent fits: {k * iters / dt:8.2f} it/s aggregate ({dt / iters * 1e3:7.2f} ms per round of {k})", ev[0], flush=True)
    for f in fits: f.close()
