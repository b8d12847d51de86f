set -x
timeout 600 python scripts/grad_sweep3.py full > gpurun_out/r02_grad_sweep3b.log 2>&1; cat gpurun_out/r02_grad_sweep3b.log
timeout 900 python -m pytest tests/test_gpu_path.py -x -q > gpurun_out/r02_j20_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_j20_pytest.log
