"""Measure the box's fp64 matmul peak (cuBLAS DGEMM through torch) - the FP64 roofline denominator."""
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
for _ in range(3): a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40): a @ b
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / 40
K = torch.randn(n, n, dtype=torch.float64, device='cuda'); K = K @ K.T + n * torch.eye(n, dtype=torch.float64, device='cuda')
torch.linalg.cholesky(K); torch.cuda.synchronize()
e0.record(); L = torch.linalg.cholesky(K); e1.record(); torch.cuda.synchronize()
tchol = e0.elapsed_time(e1)
e0.record(); Ki = torch.cholesky_inverse(L); e1.record(); torch.cuda.synchronize()
tinv = e0.elapsed_time(e1)
print(json.dumps({"dgemm_tflops_burst": 2 * n**3 / best * 1e-9, "dgemm_tflops_sustained": 2 * n**3 / sus * 1e-9,
                  "torch_cholesky_8192_ms": tchol, "torch_cholesky_inverse_8192_ms": tinv}))
