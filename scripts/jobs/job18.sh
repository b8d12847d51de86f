set -x
timeout 600 python scripts/grad_sweep3.py full > gpurun_out/r02_grad_sweep3.log 2>&1; cat gpurun_out/r02_grad_sweep3.log
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/r02_j18_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_j18_pytest.log
