"""One big lower-triangular SYRK launch of dgemm_nt_kernel (M = N = 8192, K = 4096) for an ncu --set full capture."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from additivecausalexpansion_b200 import api
M, K = 8192, 4096
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((M, K)))
C = np.zeros((M, M), order="F")
out = api.dbg_gemm_nt(A, A, C, alpha=1.0, beta=0.0, lower_only=True)
i = rng.integers(0, M, 50); j = rng.integers(0, M, 50); lo = np.maximum(i, j); hi = np.minimum(i, j)
ref = np.einsum('ik,ik->i', A[lo], A[hi])
print("max err", np.abs(out[lo, hi] - ref).max())
