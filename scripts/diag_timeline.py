"""Phase timeline of the fused diagonal-block kernel (SM clock stamps, csrc/diag_block.cuh) for one 512 x 512 block."""
import ctypes as C
import sys

sys.path.insert(0, '/root/repo')
import numpy as np

from additivecausalexpansion_b200 import api
from additivecausalexpansion_b200._lib import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
A = (Q * np.logspace(0, 3, n)) @ Q.T
A = 0.5 * (A + A.T)
for rep in range(3):
    api.dbg_diag_block(A)
buf = (C.c_longlong * 320)()
lib().ace_dbg_diag_block_timeline(buf, 320)
t = np.array(buf[:], dtype=np.int64).reshape(40, 8)
nt = 4 * ((n + 127) // 128)
GHZ = 1.965
t0 = t[0, 0]
print("k : B1 wait | own column solve | B2 wait | own trailing | F: diag update | factor   (us)")
for k in range(nt):
    T, F = t[k, :5], t[k, 5:8]
    us = lambda a, b: (b - a) / GHZ / 1e3 if a and b else float('nan')
    print(f"{k:2d}: {us(T[0], T[1]):7.2f} {us(T[1], T[2]):7.2f} {us(T[2], T[3]):7.2f} {us(T[3], T[4]):7.2f} | "
          f"{us(F[0], F[1]):7.2f} {us(F[1], F[2]):7.2f}   at {us(t0, T[0]):8.2f}")
lv = t[nt]
end = None
print("end of kernel (output layout written):")
flat = np.concatenate([t[nt], t[nt + 1], t[nt + 2]])
prev = t[nt - 1, 3]
for i, v in enumerate(flat):
    if v == 0:
        break
    print(f"  stamp {i}: +{(v - prev) / GHZ / 1e3:7.2f} us   at {(v - t0) / GHZ / 1e3:8.2f}")
    prev = v


print("first trailing-update task of each step (one warp): start -> staged | products | stores (us)")
for k in range(nt - 1):
    r = t[20 + k]
    if r[0]:
        print(f"{k:2d}: {(r[1]-r[0])/GHZ/1e3:6.2f} {(r[2]-r[1])/GHZ/1e3:6.2f} {(r[3]-r[2])/GHZ/1e3:6.2f}")
