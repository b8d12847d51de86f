import sys
sys.path.insert(0, '/root/repo')
from additivecausalexpansion_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
print(api.bench_dense(n, 1))
