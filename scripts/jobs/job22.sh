set -x
timeout 300 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r02_j22_dense.log 2>&1; echo "dense rc=$?" | tee -a gpurun_out/r02_j22_dense.log
tail -6 gpurun_out/r02_j22_dense.log
timeout 120 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v10.log 2>&1; cat gpurun_out/r02_diag_timeline_v10.log
timeout 300 python scripts/dense_only.py 4096
