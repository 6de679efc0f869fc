"""Per-problem difference of the batched C5 evaluation between the persistent and the one-CTA-per-tile GEMM grid
(sub-processes with FFGP_PERSIST=1/0), and run-to-run determinism of each."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, ROOT)
    import torch
    from fidelityfusion_b200.batched import batched_cigp_eval
    g = torch.Generator().manual_seed(5)
    B, n, d, ns = 1024, 512, 8, 64
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64)
    w = torch.randn(B, d, 1, generator=g, dtype=torch.float64)
    y = torch.sin(3 * x @ w) + 0.05 * torch.randn(B, n, 1, generator=g, dtype=torch.float64)
    ls = torch.exp(torch.rand(B, d, generator=g, dtype=torch.float64) * 2 - 1)
    sv = torch.ones(B, dtype=torch.float64); lb = torch.rand(B, generator=g, dtype=torch.float64) * 3
    xs = torch.rand(B, ns, d, generator=g, dtype=torch.float64)
    x, y, ls, sv, lb, xs = (t.cuda() for t in (x, y, ls, sv, lb, xs))
    outs = []
    for _ in range(3):
        r = batched_cigp_eval(x, y, ls, sv, lb, xs)
        outs.append(torch.cat([r['nll'].reshape(B, -1), r['g_length_scales'], r['mean'].reshape(B, -1), r['var']], 1).cpu())
    print('run-to-run max abs diff', float((outs[0] - outs[1]).abs().max()), float((outs[0] - outs[2]).abs().max()))
    torch.save(outs[0], sys.argv[2])
    sys.exit(0)
import torch
for mode in ('0', '1'):
    subprocess.run([sys.executable, __file__, 'child', f'/tmp/persist_{mode}.pt'], env=dict(os.environ, FFGP_PERSIST=mode), check=True)
a, b = torch.load('/tmp/persist_0.pt'), torch.load('/tmp/persist_1.pt')
rel = ((a - b).abs() / a.abs().clamp_min(1e-300))
print('persist 0 vs 1: max rel diff', float(rel.max()), 'problems differing', int((rel.max(1).values > 0).sum()), 'of', a.shape[0])
print('worst problems', rel.max(1).values.topk(5))
