"""Run-to-run determinism stress of the batched path (persistent TMA GEMM grid): every evaluation must reproduce the
first one bit for bit.   python tools/persist_stress.py [--iters 40]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200.batched import batched_cigp_eval
ap = argparse.ArgumentParser(); ap.add_argument('--iters', type=int, default=40)
a = ap.parse_args()
bad = 0
for (B, n, d, ns) in ((1024, 512, 8, 64), (600, 384, 5, 16), (2000, 256, 8, 32), (64, 1024, 16, 64), (333, 640, 3, 0)):
    g = torch.Generator().manual_seed(B + n)
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64)
    w = torch.randn(B, d, 1, generator=g, dtype=torch.float64)
    y = torch.sin(3 * x @ w) + 0.05 * torch.randn(B, n, 1, generator=g, dtype=torch.float64)
    ls = torch.exp(torch.rand(B, d, generator=g, dtype=torch.float64) * 2 - 1)
    sv = torch.ones(B, dtype=torch.float64); lb = torch.rand(B, generator=g, dtype=torch.float64) * 3
    xs = torch.rand(B, ns, d, generator=g, dtype=torch.float64) if ns else None
    x, y, ls, sv, lb = (t.cuda() for t in (x, y, ls, sv, lb))
    xs = xs.cuda() if ns else None
    first, ndiff, worst = None, 0, 0.0
    for it in range(a.iters):
        r = batched_cigp_eval(x, y, ls, sv, lb, xs)
        cat = torch.cat([r['nll'].reshape(B, -1), r['g_length_scales'], r['g_log_beta'].reshape(B, -1)] +
                        ([r['mean'].reshape(B, -1), r['var']] if ns else []), 1)
        if first is None: first = cat.clone()
        else:
            dd = float((cat - first).abs().max())
            if dd > 0: ndiff += 1; worst = max(worst, dd)
    print(f'B={B} n={n} d={d} ns={ns}: {ndiff} of {a.iters - 1} repeats differ, worst abs diff {worst:.3e}')
    bad += ndiff
sys.exit(1 if bad else 0)
