"""Stress of the TMA GEMM stage release (persistent grid, beta != 0: C is read at the start of every tile, which backs up
the load/store queue while the ring is full).  Reports tiles whose result differs from torch and, for those, which 8x8
fragments of which warp are wrong.  History: with the stage released right after ISSUING the last fragment loads, ~1 tile
in 500 came out wrong in one warp's last-loaded A / B fragments (see gemm_tma.cuh, "Stage release").
   FFGP_PERSIST=2 python tools/diag_stage_release.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200 import _lib as B
L = B.lib(); st = B.stream_ptr()
g = torch.Generator(device='cuda').manual_seed(3)
batch, M, N, K = 200, 256, 256, 256
A = torch.randn(batch, K, M, generator=g, dtype=torch.float64, device='cuda')
Bm = torch.randn(batch, N, K, generator=g, dtype=torch.float64, device='cuda')
C0 = torch.randn(batch, M, N, generator=g, dtype=torch.float64, device='cuda')
for rep in range(8):
    C = C0.clone()
    rc = L.ffgp_gemm_f64(0, 1, B.ptr(A), M, K * M, B.ptr(Bm), K, N * K, B.ptr(C), N, M * N, M, N, K, 0.75, 1.0, 0, 0, batch, st)
    AB = 0.75 * (A.transpose(1, 2) @ Bm.transpose(1, 2))
    ref = AB + C0
    err = (C - ref).abs().reshape(batch, 2, 128, 2, 128).amax(dim=(2, 4))          # [batch, ti, tj]
    bad = (err > 1e-9).nonzero()
    print('rep', rep, 'bad tiles', bad.shape[0], 'of', batch * 4)
    for z, ti, tj in bad[:12].tolist():
        sl = (z, slice(ti * 128, ti * 128 + 128), slice(tj * 128, tj * 128 + 128))
        d = C[sl] - AB[sl]          # what was added instead of C0
        c0 = C0[sl]
        frac_zero = float((d.abs() < 1e-9).double().mean()); frac_one = float(((d - c0).abs() < 1e-9).double().mean()); frac_two = float(((d - 2 * c0).abs() < 1e-9).double().mean())
        rows_bad = ((d - c0).abs() > 1e-9).any(1).nonzero().reshape(-1)
        cols_bad = ((d - c0).abs() > 1e-9).any(0).nonzero().reshape(-1)
        bm = ((d - c0).abs() > 1e-9)
        fr = bm.reshape(16, 8, 16, 8).any(3).any(1)                       # [row frag 0..15][col frag 0..15]
        print('   bad 8x8 fragments (row frag, col frag):', fr.nonzero().tolist()[:40])
        e = (d - c0)[bm]
        print('   err sample', e[:4].tolist(), ' c0 there', c0[bm][:4].tolist(), ' AB there', AB[sl][bm][:4].tolist())
        print(f'  w={z * 4 + (tj * 2 + ti)} z={z} ti={ti} tj={tj}: added 0*C {frac_zero:.3f}  1*C {frac_one:.3f}  2*C {frac_two:.3f}; bad rows {rows_bad[:6].tolist()}..({rows_bad.numel()}) bad cols {cols_bad[:6].tolist()}..({cols_bad.numel()})')
