"""Phase timing of one dense evaluation (CUDA events, warm): runs the C ABI with FFGP_DEBUG_STOP_AFTER = 1, 2, 3, 0 in
sub-processes and prints cumulative and per-phase times.   python tools/phase_times.py [--n 8192 --d 16 --batch 1]"""
import argparse, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=8192); ap.add_argument('--d', type=int, default=16)
ap.add_argument('--batch', type=int, default=1); ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--child', type=int, default=-1)
a = ap.parse_args()
if a.child >= 0:
    sys.path.insert(0, ROOT)
    import torch
    from fidelityfusion_b200 import _lib as B
    L = B.lib()
    g = torch.Generator().manual_seed(0)
    bt = a.batch
    x = torch.randn(bt, a.n, a.d, generator=g, dtype=torch.float64).cuda()
    y = torch.randn(bt, a.n, 1, generator=g, dtype=torch.float64).cuda()
    il = torch.full((a.d,), 0.3 if a.n > 1024 else 1.0, dtype=torch.float64).cuda(); amp = torch.ones(1, dtype=torch.float64).cuda()
    dg = torch.full((a.n,), 0.37, dtype=torch.float64).cuda()
    nll = torch.empty(bt, dtype=torch.float64).cuda(); alpha = torch.empty(bt, a.n, 1, dtype=torch.float64).cuda()
    gil = torch.empty(bt, a.d, dtype=torch.float64).cuda(); gamp = torch.empty(bt, dtype=torch.float64).cuda()
    gd = torch.empty(bt, a.n, dtype=torch.float64).cuda(); info = torch.zeros(bt, dtype=torch.int32).cuda()
    wsb = L.ffgp_dense_workspace_bytes(a.n, a.d, 1, 0, bt); ws = torch.empty(wsb, dtype=torch.uint8).cuda()
    def run():
        rc = L.ffgp_dense_nll_f64(B.ptr(x), B.ptr(y), B.ptr(il), B.ptr(amp), B.ptr(dg), None, a.n, a.d, 1, bt, 0, 1, 1, B.ptr(ws), wsb,
                                  B.ptr(nll), None, B.ptr(alpha), B.ptr(gil), B.ptr(gamp), B.ptr(gd), None, B.ptr(info), B.stream_ptr())
        assert rc == 0, L.ffgp_last_error_string()
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps): run()
    e1.record(); torch.cuda.synchronize()
    print(e0.elapsed_time(e1) / a.reps)
    sys.exit(0)
names = {1: 'kernel matrix + potrf', 2: '+ trtri', 3: '+ solves + S=M^T M', 0: '+ gradient contraction (full)'}
prev = 0.0
for stop in (1, 2, 3, 0):
    env = dict(os.environ, FFGP_DEBUG_STOP_AFTER=str(stop))
    out = subprocess.run([sys.executable, __file__, '--child', str(stop), '--n', str(a.n), '--d', str(a.d), '--batch', str(a.batch),
                          '--reps', str(a.reps)], env=env, capture_output=True, text=True)
    try:
        t = float(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(out.stdout, out.stderr); raise
    print(f'{names[stop]:34s} cumulative {t:8.3f} ms   phase {t - prev:8.3f} ms')
    prev = t
