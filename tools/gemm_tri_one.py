"""One shape of tools/gemm_tri_bench.py, for ncu:  python tools/gemm_tri_one.py <akm> <bkm> <kmode> <M> <N> <K> [batch [lower [beta]]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200 import _lib as B
akm, bkm, kmode, M, N, K = [int(v) for v in sys.argv[1:7]]
batch = int(sys.argv[7]) if len(sys.argv) > 7 else 4 * 148
lower = int(sys.argv[8]) if len(sys.argv) > 8 else 0
beta = float(sys.argv[9]) if len(sys.argv) > 9 else 0.0
L = B.lib(); st = B.stream_ptr()
g = torch.Generator(device='cuda').manual_seed(3)
A = torch.randn(batch, *((M, K) if akm else (K, M)), generator=g, dtype=torch.float64, device='cuda')
Bm = torch.randn(batch, *((N, K) if bkm else (K, N)), generator=g, dtype=torch.float64, device='cuda')
C = torch.zeros(batch, M, N, dtype=torch.float64, device='cuda')
for _ in range(6):
    L.ffgp_gemm_f64(akm, bkm, B.ptr(A), A.shape[2], A.shape[1] * A.shape[2], B.ptr(Bm), Bm.shape[2], Bm.shape[1] * Bm.shape[2],
                    B.ptr(C), N, M * N, M, N, K, 1.0, beta, lower, kmode, batch, st)
torch.cuda.synchronize()
