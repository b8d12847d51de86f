timeout 1200 python scripts/config_runs.py M2,C4 2>&1 | tail -6
