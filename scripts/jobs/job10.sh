set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_j10_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_j10_pytest.log
tail -15 gpurun_out/r02_j10_pytest.log
timeout 300 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_final.log 2>&1; tail -4 gpurun_out/r02_diag_timeline_final.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_j10_bench.json 2> gpurun_out/r02_j10_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_j10_bench.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"grad_kernel|kernmat_kernel" -s 4 -c 2 -o gpurun_out/r02_pair_kernels python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log
