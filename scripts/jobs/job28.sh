timeout 300 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r02_j31_dense.log 2>&1; echo "dense rc=$?"; tail -3 gpurun_out/r02_j28_dense.log
timeout 120 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v19.log 2>&1; cat gpurun_out/r02_diag_timeline_v16.log | head -22
timeout 300 python scripts/dense_only.py 4096
