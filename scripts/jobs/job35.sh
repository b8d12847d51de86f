timeout 600 python -m pytest tests/test_gpu_shard_emulated.py -x -q 2>&1 | tail -3
