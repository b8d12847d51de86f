set -x
timeout 300 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r02_j17_dense.log 2>&1; echo "dense rc=$?" | tee -a gpurun_out/r02_j17_dense.log
tail -4 gpurun_out/r02_j17_dense.log
timeout 300 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v9.log 2>&1; tail -8 gpurun_out/r02_diag_timeline_v9.log
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"grad2_kernel|kernmat_kernel" -s 4 -c 2 -o gpurun_out/r02_pair_kernels_v2 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_pair2.log 2>&1
tail -3 gpurun_out/ncu_pair2.log
