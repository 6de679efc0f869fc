"""Stress of the batched path's run-to-run reproducibility: many back-to-back sweeps of one batch, every sweep's packed result
rows compared ON THE DEVICE with the first sweep's (one flag read back per sweep group); mismatching sweeps are reported with
the problem index, the column of the packed row and the size of the difference.
  python tools/c5_stress.py [batch] [sweeps] [check=0|1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200.batched import batched_cigp_eval, result_layout
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 500
check = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
n, d, ns = 512, 8, 64
g = torch.Generator().manual_seed(5)
x = torch.rand(B, n, d, generator=g, dtype=torch.float64); w = torch.randn(B, d, 1, generator=g, dtype=torch.float64)
y = torch.sin(3 * sum(x[..., k:k + 1] * w[:, k:k + 1, :] for k in range(d))) + 0.05 * torch.randn(B, n, 1, generator=g, dtype=torch.float64)
ls = torch.exp(torch.rand(B, d, generator=g, dtype=torch.float64) * 2 - 1); sv = torch.ones(B, dtype=torch.float64)
lb = torch.rand(B, generator=g, dtype=torch.float64) * 3; xs = torch.rand(B, ns, d, generator=g, dtype=torch.float64)
x, y, ls, sv, lb, xs = (t.cuda() for t in (x, y, ls, sv, lb, xs))
ref = batched_cigp_eval(x, y, ls, sv, lb, xs, check=check)['_packed'].clone()
keys, widths = result_layout(d, 1, ns, True, check)
bad = 0
for s in range(S):
    cur = batched_cigp_eval(x, y, ls, sv, lb, xs, check=check)['_packed']
    ne = cur != ref
    if bool(ne.any()):
        bad += 1
        idx = ne.nonzero()
        probs = sorted(set(int(i) for i in idx[:, 0].tolist()))
        cols = sorted(set(int(i) for i in idx[:, 1].tolist()))
        diff = (cur - ref).abs()
        print(f'sweep {s}: {idx.shape[0]} entries differ; problems {probs[:8]}{"..." if len(probs) > 8 else ""} ({len(probs)}), '
              f'columns {cols[:12]}{"..." if len(cols) > 12 else ""} ({len(cols)}), max abs {float(diff.max()):.3e}, '
              f'max rel {float((diff / ref.abs().clamp_min(1e-300)).max()):.3e}', flush=True)
print(f'batch {B}, {S} sweeps (check={check}, FFGP_SPLIT={os.environ.get("FFGP_SPLIT", "default")}): {bad} sweeps differ from the first; '
      f'row layout {list(zip(keys, widths))}')
