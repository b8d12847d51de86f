timeout 300 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r02_j27_dense.log 2>&1; echo "dense rc=$?"; tail -3 gpurun_out/r02_j27_dense.log
timeout 120 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v15.log 2>&1; cat gpurun_out/r02_diag_timeline_v15.log | head -22
timeout 300 python scripts/dense_only.py 4096
timeout 300 ncu --set full --clock-control none --import-source on -k regex:diag_block_kernel -s 10 -c 1 -o gpurun_out/r02_diag_v15 python scripts/dense_only.py 4096 > gpurun_out/ncu_diag.log 2>&1; tail -2 gpurun_out/ncu_diag.log
