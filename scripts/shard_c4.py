"""torchrun --nproc-per-node G scripts/shard_c4.py : BASELINE config 4 (n = 65536, SE, cubic basis) as ONE fit sharded
over G GPUs.  Iteration 1 is checked against the single-GPU statistics recorded in profiles/r01/config_runs.json
(same seeded problem), iteration 2 is timed (iteration 1 also pays NCCL's lazy connection set-up)."""
import json, os, sys, time
sys.path.insert(0, '/root/repo')
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
import torch, torch.distributed as dist
from additivecausalexpansion_b200 import synth
from additivecausalexpansion_b200.fit import AceFit

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
prob = synth.make_problem(cfg)
ref = None
try:
    ref = json.load(open('/root/repo/profiles/r01/config_runs.json'))[cfg]["stats"]
except Exception:
    pass
with AceFit(prob.y, prob.X, prob.Z, prob.parameters, kernel=prob.kernel, std_y=prob.std_y, device=lr, use_graph=False) as f:
    f.shard(dist)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time(); st1, _ = f.para_update(1); torch.cuda.synchronize(); w1 = time.time() - t0
    ph1 = f.last_timing_ms
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time(); st2, _ = f.para_update(2); torch.cuda.synchronize(); w2 = time.time() - t0
    ph2 = f.last_timing_ms
    out = {"config": cfg, "n": prob.n, "world": world, "iter1_wall_s": w1, "iter2_wall_s": w2, "phases_iter1_ms": ph1,
           "phases_iter2_ms": ph2, "stats_iter1": st1.tolist(), "stats_iter2": st2.tolist(), "single_gpu_stats_iter1": ref}
    if ref is not None:
        out["evid_rel_vs_single_gpu"] = abs(st1[1] - ref[1]) / abs(ref[1])
        out["rmse_rel_vs_single_gpu"] = abs(st1[0] - ref[0]) / abs(ref[0])
    n = prob.n
    out["dense_tflops_per_gpu_equiv"] = n ** 3 / ((ph2["potrf"] + ph2["trtri"] + ph2["uut"]) * 1e-3) * 1e-12
if rank == 0:
    print(json.dumps(out, default=float))
    json.dump(out, open(f"/root/repo/gpurun_out/shard_{cfg}_w{world}.json", "w"), indent=1, default=float)
dist.barrier()
dist.destroy_process_group()
