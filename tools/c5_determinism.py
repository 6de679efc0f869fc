"""Run-to-run bit reproducibility of the batched path (BASELINE config 5): several sweeps in one process compared bit for bit,
and a digest of every output for comparison across processes.   python tools/c5_determinism.py [batch]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200.batched import batched_cigp_eval
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n, d, ns = 512, 8, 64
g = torch.Generator().manual_seed(5)
x = torch.rand(B, n, d, generator=g, dtype=torch.float64); w = torch.randn(B, d, 1, generator=g, dtype=torch.float64)
y = torch.sin(3 * sum(x[..., k:k + 1] * w[:, k:k + 1, :] for k in range(d))) + 0.05 * torch.randn(B, n, 1, generator=g, dtype=torch.float64)
ls = torch.exp(torch.rand(B, d, generator=g, dtype=torch.float64) * 2 - 1); sv = torch.ones(B, dtype=torch.float64)
lb = torch.rand(B, generator=g, dtype=torch.float64) * 3; xs = torch.rand(B, ns, d, generator=g, dtype=torch.float64)
x, y, ls, sv, lb, xs = (t.cuda() for t in (x, y, ls, sv, lb, xs))
ref = None
for rep in range(4):
    r = batched_cigp_eval(x, y, ls, sv, lb, xs)
    cur = {k: v.detach().cpu().clone() for k, v in r.items() if torch.is_tensor(v)}
    if ref is None: ref = cur
    else:
        for k in ref:
            if not torch.equal(ref[k], cur[k]):
                diff = (ref[k].double() - cur[k].double()).abs()
                print(f'sweep {rep}: {k} differs from sweep 0 in {int((diff > 0).sum())} entries, max abs {float(diff.max()):.3e}')
h = hashlib.sha256()
for k in sorted(ref): h.update(ref[k].numpy().tobytes())
print('digest', h.hexdigest()[:16], 'nll sum (cpu, fixed order)', float(ref['nll'].double().sum()), 'max |nll|', float(ref['nll'].abs().max()))
