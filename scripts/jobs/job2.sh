set -x
timeout 600 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r02_j2_dense.log 2>&1; echo "dense rc=$?" | tee -a gpurun_out/r02_j2_dense.log
tail -15 gpurun_out/r02_j2_dense.log
timeout 900 python -m pytest tests/test_gpu_shard_emulated.py tests/test_gpu_path.py -x -q > gpurun_out/r02_j2_emul.log 2>&1; echo "emul rc=$?" | tee -a gpurun_out/r02_j2_emul.log
tail -15 gpurun_out/r02_j2_emul.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_j2_bench.json 2> gpurun_out/r02_j2_bench.err; echo "bench rc=$?"
ACE_DIAG_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_j2_bench_nofuse.json 2> gpurun_out/r02_j2_bench_nofuse.err; echo "bench rc=$?"
ACE_POTRF_TRACE=1 timeout 300 python scripts/dense_only.py 16384 > gpurun_out/r02_j2_potrf_trace_16384.log 2>&1
timeout 300 python scripts/dense_only.py 4096 > gpurun_out/r02_j2_dense_4096.log 2>&1
tail -3 gpurun_out/r02_j2_bench.err
