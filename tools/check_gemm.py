"""Checks ffgp_gemm_f64 (every operand layout x K-range mode x lower_only x batch) against torch.matmul on the GPU.
Run by tests/test_gpu_gemm.py in sub-processes with FFGP_PERSIST=0 (one CTA per tile) and FFGP_PERSIST=2 (persistent
grid forced for every launch).  Prints 'max rel err <v>' per case and exits non-zero on a mismatch."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200 import _lib as B

L = B.lib()
st = B.stream_ptr()
g = torch.Generator(device='cuda').manual_seed(11)
bad = 0
# kmode -> (a_kmaj, b_kmaj) the mode is defined for (ffgp.h) and which operand is triangular
KMODES = {0: None, 1: 'A_le_row', 2: 'B_le_col', 3: 'B_ge_col', 4: 'A_ge_row'}
cases = []
for akm, bkm, kmode in itertools.product((1, 0), (1, 0), range(5)):
    if kmode == 1 and not akm: continue      # A[i][p] lower
    if kmode == 4 and akm: continue          # A[p][i] lower
    if kmode == 2 and not bkm: continue      # B[j][p] lower
    if kmode == 3 and bkm: continue          # B[p][j] lower
    for lower_only, batch, (M, N, K), beta in (
            (0, 1, (256, 384, 256), 0.0), (0, 300, (128, 128, 128), 0.0), (0, 160, (256, 256, 256), 0.0),
            (1, 170, (384, 384, 384), 1.0), (0, 9, (512, 512, 512), -0.5), (0, 200, (256, 256, 256), 1.0),
            (1, 170, (384, 384, 384), 0.0), (1, 400, (128, 128, 128), 1.0), (0, 1, (1024, 1024, 512), 0.5),
            # 64-wide / 192-wide outputs: the 128 x 64-tile variant (two CTAs per SM), also forced by FFGP_GEMM_N64=2
            (0, 320, (128, 64, 128), 0.0), (0, 200, (256, 192, 256), 1.0), (0, 310, (128, 128, 256), -0.5), (0, 150, (512, 64, 512), 0.0)):
        if kmode in (1, 4) and M > K: continue
        if kmode in (2, 3) and N > K: continue
        if kmode in (1, 2, 3, 4) and (M != K or N != K) and lower_only: continue
        cases.append((akm, bkm, kmode, lower_only, batch, M, N, K, beta))
SMALL = os.environ.get('FFGP_CHECK_SMALL') == '1'        # compute-sanitizer runs: same cases, batches cut to <= 12
for akm, bkm, kmode, lower_only, batch, M, N, K, beta in cases:
    if SMALL:
        batch = min(batch, 12)
        if M * N * K > 256 ** 3: continue
    A = torch.randn(batch, *((M, K) if akm else (K, M)), generator=g, dtype=torch.float64, device='cuda')
    Bm = torch.randn(batch, *((N, K) if bkm else (K, N)), generator=g, dtype=torch.float64, device='cuda')
    if kmode in (1, 4): A = torch.tril(A)
    if kmode in (2, 3): Bm = torch.tril(Bm)
    C0 = torch.randn(batch, M, N, generator=g, dtype=torch.float64, device='cuda')
    C = C0.clone()
    alpha = 0.75
    rc = L.ffgp_gemm_f64(akm, bkm, B.ptr(A), A.shape[2], A.shape[1] * A.shape[2], B.ptr(Bm), Bm.shape[2],
                         Bm.shape[1] * Bm.shape[2], B.ptr(C), N, M * N, M, N, K, alpha, beta, lower_only, kmode, batch, st)
    assert rc == 0, L.ffgp_last_error_string()
    Aop = A if akm else A.transpose(1, 2)
    Bop = Bm.transpose(1, 2) if bkm else Bm
    ref = alpha * (Aop @ Bop) + beta * C0
    if lower_only:
        err = (torch.tril(C) - torch.tril(ref)).abs().max() / ref.abs().max()
        keep = (torch.triu(C, 128) - torch.triu(C0, 128)).abs().max()      # tiles above the diagonal are never touched
        if float(keep) > 0: print('   upper tiles touched:', float(keep))
        err = max(float(err), float(keep))
    else:
        err = float((C - ref).abs().max() / ref.abs().max())
    ok = err < 1e-13
    bad += not ok
    print(f'akm={akm} bkm={bkm} kmode={kmode} lower={lower_only} batch={batch} {M}x{N}x{K} beta={beta}: max rel err {err:.2e}'
          + ('' if ok else '   <-- MISMATCH'))
print('cases', len(cases), 'bad', bad)
sys.exit(1 if bad else 0)
