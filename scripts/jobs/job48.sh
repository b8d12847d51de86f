timeout 600 python scripts/graph_vs_eager.py 2>&1 | tail -6
