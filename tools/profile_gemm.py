"""Times (CUDA events) or exposes to ncu single launches of the DMMA GEMM in the shapes the dense path uses at N=8192:
   lauum  S = M^T M       (A m-major, B n-major, lower tiles, K >= i0)   N^3/3 FLOP
   syrk   C -= A A^T      (A,B k-major, lower tiles, full K)            n x n x k
   trsm   L21 = A21 M11^T (k-major/k-major, K <= j0)                    top level of the recursion
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200 import _lib as B
ap = argparse.ArgumentParser(); ap.add_argument('--n', type=int, default=8192); ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--which', default='lauum,syrk,trsm,full')
a = ap.parse_args()
L = B.lib(); n = a.n
M = torch.tril(torch.randn(n, n, dtype=torch.float64, device='cuda')) * 0.01
C = torch.zeros(n, n, dtype=torch.float64, device='cuda')
def run(name, fn, flops):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(f'{name:8s} {ms:8.3f} ms  {flops / ms * 1e-9:7.2f} TFLOP/s')
st = B.stream_ptr()
h = n // 2
cfg = {
 'lauum': (lambda: L.ffgp_gemm_f64(0, 0, B.ptr(M), n, 0, B.ptr(M), n, 0, B.ptr(C), n, 0, n, n, n, 1.0, 0.0, 1, 4, 1, st), n**3 / 3),
 'syrk': (lambda: L.ffgp_gemm_f64(1, 1, B.ptr(M), n, 0, B.ptr(M), n, 0, B.ptr(C), n, 0, h, h, h, -1.0, 1.0, 1, 0, 1, st), h**3),
 'trsm': (lambda: L.ffgp_gemm_f64(1, 1, B.ptr(M), n, 0, B.ptr(M), n, 0, B.ptr(C), n, 0, h, h, h, 1.0, 0.0, 0, 2, 1, st), h**3),
 'full': (lambda: L.ffgp_gemm_f64(1, 0, B.ptr(M), n, 0, B.ptr(M), n, 0, B.ptr(C), n, 0, h, h, h, 1.0, 0.0, 0, 0, 1, st), 2 * h**3),
}
for w in a.which.split(','):
    run(w, *cfg[w])
