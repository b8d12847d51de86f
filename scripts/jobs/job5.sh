set -x
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_shard_emulated.py -x -q > gpurun_out/r02_j5_shard.log 2>&1; echo "shard rc=$?" | tee -a gpurun_out/r02_j5_shard.log
tail -12 gpurun_out/r02_j5_shard.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_j5_bench_2gpu.json 2> gpurun_out/r02_j5_bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02_j5_bench_2gpu.err; cat gpurun_out/r02_j5_bench_2gpu.json | cut -c1-1500
timeout 600 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v3.log 2>&1; head -20 gpurun_out/r02_diag_timeline_v3.log
