ACE_POTRF_TRACE=1 timeout 300 python scripts/dense_only.py 4096 2>&1 | tail -14
