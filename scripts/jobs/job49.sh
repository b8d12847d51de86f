timeout 900 python -m pytest tests/test_gpu_exact_shapes.py -x -q 2>&1 | tail -15
