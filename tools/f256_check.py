"""factor256_kernel (csrc/factor256.cuh) against torch.linalg and against the six-launch path it replaces:
   python tools/f256_check.py            (sub-processes with FFGP_F256 = 0, 3, 2; 3 = fused kernel for single problems too)
prints max errors of L, L^-1, log|A| per (n, batch) and the time per call (CUDA events)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, ROOT)
    import torch
    from fidelityfusion_b200 import ops
    torch.manual_seed(0)
    for n, batch in ((256, 1), (256, 5), (512, 3), (384, 2), (1024, 1), (300, 2), (512, 296), (2048, 1)):
        X = torch.randn(batch, n, n + 8, dtype=torch.float64, device='cuda')
        A = X @ X.transpose(1, 2) / n + 0.5 * torch.eye(n, dtype=torch.float64, device='cuda')
        L, Mi, ld = ops.potrf_trtri(A)
        Lr = torch.linalg.cholesky(A)
        eL = float((L - Lr).abs().max() / Lr.abs().max())
        eM = float((Mi @ Lr - torch.eye(n, dtype=torch.float64, device='cuda')).abs().max())
        eD = float((ld - 2 * Lr.diagonal(dim1=1, dim2=2).log().sum(1)).abs().max())      # log|A|
        for _ in range(2): ops.potrf_trtri(A, want_L=False, want_inv=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.potrf_trtri(A, want_L=False, want_inv=False)
        e1.record(); torch.cuda.synchronize()
        print(f'F256={os.environ.get("FFGP_F256")} n={n:5d} batch={batch:4d}  |L-Lref|={eL:.2e}  |M L-I|={eM:.2e}  |logdet|={eD:.2e}  '
              f'{e0.elapsed_time(e1) / 5 * 1e3:9.1f} us/call', flush=True)
        assert eL < 1e-11 and eM < 1e-10 and eD < 1e-9
    # non-PD input: the failing column is reported through info like the base kernel does
    A = torch.eye(256, dtype=torch.float64, device='cuda'); A[200, 200] = -1.0
    try:
        ops.potrf_trtri(A); print('non-PD NOT detected'); sys.exit(1)
    except torch.linalg.LinAlgError as e:
        print('non-PD detected:', str(e)[:80])
    sys.exit(0)
for mode in ('0', '3', '2'):
    r = subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, FFGP_F256=mode), capture_output=True, text=True)
    print(r.stdout, r.stderr[-2000:] if r.returncode else '')
    if r.returncode: sys.exit(r.returncode)
