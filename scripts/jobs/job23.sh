timeout 120 python scripts/diag_timeline.py 512 > gpurun_out/r02_diag_timeline_v11.log 2>&1; tail -20 gpurun_out/r02_diag_timeline_v11.log
