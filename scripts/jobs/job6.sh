set -x
timeout 300 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r02_j16_dense.log 2>&1; echo "dense rc=$?" | tee -a gpurun_out/r02_j16_dense.log
tail -12 gpurun_out/r02_j16_dense.log
timeout 300 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v8.log 2>&1; cat gpurun_out/r02_diag_timeline_v8.log
ACE_POTRF_TRACE=1 timeout 300 python scripts/dense_only.py 16384 > gpurun_out/r02_j16_potrf_trace_16384.log 2>&1; tail -8 gpurun_out/r02_j16_potrf_trace_16384.log
timeout 300 python scripts/dense_only.py 4096
