"""Does any result depend on workspace memory the library did not write in the same call?
The caller-owned workspace is poisoned (zeros / NaN / large finite noise / the previous call's contents of a DIFFERENT problem
size) before an evaluation; every output must be bit-identical across the poisons.
  python tools/ws_poison.py [batch] [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200 import ops, _lib as B
from fidelityfusion_b200.batched import batched_cigp_eval
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
d, ns = 8, 64
g = torch.Generator().manual_seed(5)
x = torch.rand(Bn, n, d, generator=g, dtype=torch.float64); w = torch.randn(Bn, d, 1, generator=g, dtype=torch.float64)
y = torch.sin(3 * sum(x[..., k:k + 1] * w[:, k:k + 1, :] for k in range(d))) + 0.05 * torch.randn(Bn, n, 1, generator=g, dtype=torch.float64)
ls = torch.exp(torch.rand(Bn, d, generator=g, dtype=torch.float64) * 2 - 1); sv = torch.ones(Bn, dtype=torch.float64)
lb = torch.rand(Bn, generator=g, dtype=torch.float64) * 3; xs = torch.rand(Bn, ns, d, generator=g, dtype=torch.float64)
x, y, ls, sv, lb, xs = (t.cuda() for t in (x, y, ls, sv, lb, xs))
L = B.lib()
wsb = L.ffgp_dense_workspace_bytes(n, d, 1, ns, Bn)
ws = ops._ws_cache.get(wsb, x.device)                      # the buffer batched_cigp_eval will be handed
wsd = ws[: (ws.numel() // 8) * 8].view(torch.float64)
def poison(kind):
    if kind == 'zeros': wsd.zero_()
    elif kind == 'nan': wsd.fill_(float('nan'))
    elif kind == 'noise': wsd.copy_(torch.randn(wsd.numel(), device='cuda', dtype=torch.float64) * 1e3)
    elif kind == 'inf': wsd.fill_(float('inf'))
    torch.cuda.synchronize()
ref = None
for want_grad in (True, False):
    ref = None
    for kind in ('zeros', 'nan', 'noise', 'inf', 'zeros'):
        poison(kind)
        assert ops._ws_cache.get(wsb, x.device).data_ptr() == ws.data_ptr()
        out = batched_cigp_eval(x, y, ls, sv, lb, xs, want_grad=want_grad, check=False)['_packed'].clone()
        if ref is None:
            ref = out
            continue
        same = torch.equal(out, ref) or bool(((out == ref) | (out.isnan() & ref.isnan())).all())
        if same:
            print(f'want_grad={want_grad} poison {kind:6s}: identical')
        else:
            ne = ~((out == ref) | (out.isnan() & ref.isnan()))
            idx = ne.nonzero()
            diff = (out - ref).abs()
            print(f'want_grad={want_grad} poison {kind:6s}: {idx.shape[0]} entries differ, problems {sorted(set(idx[:, 0].tolist()))[:10]}, '
                  f'columns {sorted(set(idx[:, 1].tolist()))[:16]}, nan in out {int(out.isnan().sum())}, max abs {float(diff[~diff.isnan()].max()) if (~diff.isnan()).any() else float("nan"):.3e}')
print('workspace bytes', wsb, 'batch', Bn, 'n', n)
