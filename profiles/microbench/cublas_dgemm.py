# cuBLAS DGEMM calibrator (OFF the product path): achievable FP64 GEMM rate on this B200.
import json, torch
torch.backends.cuda.matmul.allow_tf32 = False
out = {}
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[str(n)] = {"ms": best, "tflops": 2 * n**3 / best * 1e-9}
    del a, b, c
print(json.dumps({"cublas_dgemm": out}))
