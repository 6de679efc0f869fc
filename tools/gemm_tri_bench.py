"""Per-tile time of the TMA + DMMA GEMM on the small-K shapes of the batched factorisation (N = 512: 128- and 256-deep
tiles), dense vs triangular K blocks.  batch = 4 x 148 problems of ONE 128 x 128 tile (or 2 x 2 tiles) each, so a launch
is exactly 4 (16) rounds on 148 SMs.  Run with FFGP_PERSIST=2 (persistent grid) and FFGP_PERSIST=0 (one CTA per tile).
  dense 128-cube at the DMMA issue peak (37.1 TFLOP/s / 148 SMs): 16.7 us"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200 import _lib as B
L = B.lib(); st = B.stream_ptr()
g = torch.Generator(device='cuda').manual_seed(3)
def run(akm, bkm, kmode, M, N, K, batch, lower=0, beta=0.0):
    A = torch.randn(batch, *((M, K) if akm else (K, M)), generator=g, dtype=torch.float64, device='cuda')
    Bm = torch.randn(batch, *((N, K) if bkm else (K, N)), generator=g, dtype=torch.float64, device='cuda')
    C = torch.zeros(batch, M, N, dtype=torch.float64, device='cuda')
    f = lambda: L.ffgp_gemm_f64(akm, bkm, B.ptr(A), A.shape[2], A.shape[1] * A.shape[2], B.ptr(Bm), Bm.shape[2],
                                Bm.shape[1] * Bm.shape[2], B.ptr(C), N, M * N, M, N, K, 1.0, beta, lower, kmode, batch, st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
nb = 4 * 148
print('persist', os.environ.get('FFGP_PERSIST', '1'), 'n64', os.environ.get('FFGP_GEMM_N64', '1'))
for name, akm, bkm, kmode in (('dense  A[i][p] B[j][p]', 1, 1, 0), ('dense  A[i][p] B[p][j]', 1, 0, 0), ('dense  A[p][i] B[p][j]', 0, 0, 0),
                              ('K_LE_COL (trsm  L21 = A21 M11^T)', 1, 1, 2), ('K_LE_ROW (T = M22 L21)', 1, 0, 1),
                              ('K_GE_COL (M21 = -T M11)', 1, 0, 3), ('K_GE_ROW (S = M^T M)', 0, 0, 4)):
    for (M, N, K) in ((128, 128, 128), (256, 256, 256)):
        us = run(akm, bkm, kmode, M, N, K, nb)
        tiles = (M // 128) * (N // 128)
        rounds = nb * tiles / 148
        print(f'{name:36s} {M}x{N}x{K}: {us:8.1f} us/launch  {us / rounds:6.2f} us per tile-round')
us = run(1, 1, 0, 256, 256, 256, nb, lower=1, beta=1.0)
print(f'syrk lower 256 (beta=1)                                 : {us:8.1f} us/launch  {us / (nb * 3 / 148):6.2f} us per tile-round')
us = run(0, 0, 4, 512, 512, 512, 148, lower=1)
print(f'S = M^T M 512 lower (10 tiles, K_GE_ROW), batch 148       : {us:8.1f} us/launch  (20 K-blocks of 128 per problem, ideal 20 x 16.7 = 334; tri half: 250)')
