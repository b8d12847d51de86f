timeout 300 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r02_j26_dense.log 2>&1; echo "dense rc=$?"; tail -3 gpurun_out/r02_j25_dense.log
timeout 120 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v14_chain4.log 2>&1; cat gpurun_out/r02_diag_timeline_v13_nc16.log | head -24
timeout 300 python scripts/dense_only.py 4096
ACE_DIAG_CLUSTER=8 timeout 300 python scripts/dense_only.py 4096
timeout 300 python scripts/dense_only.py 16384
ACE_DIAG_CLUSTER=8 timeout 300 python scripts/dense_only.py 16384
