"""One chunk of the batched config (C5: B x N=512, d=8, N*=64, NLL + grad + predict) for an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv \
      python tools/launches_c5.py [--batch 1024] [--evals 1]
Without ncu it prints the CUDA-event time of one warm evaluation."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200.batched import batched_cigp_eval

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--n', type=int, default=512)
ap.add_argument('--d', type=int, default=8)
ap.add_argument('--ns', type=int, default=64)
ap.add_argument('--evals', type=int, default=1)
ap.add_argument('--warmup', type=int, default=1)
a = ap.parse_args()
g = torch.Generator().manual_seed(5)
B, n, d, ns = a.batch, a.n, a.d, a.ns
x = torch.rand(B, n, d, generator=g, dtype=torch.float64)
w = torch.randn(B, d, 1, generator=g, dtype=torch.float64)
y = torch.sin(3 * sum(x[..., k:k + 1] * w[:, k:k + 1, :] for k in range(d))) + 0.05 * torch.randn(B, n, 1, generator=g, dtype=torch.float64)
ls = torch.exp(torch.rand(B, d, generator=g, dtype=torch.float64) * 2 - 1)
sv = torch.ones(B, dtype=torch.float64)
lb = torch.rand(B, generator=g, dtype=torch.float64) * 3
xs = torch.rand(B, ns, d, generator=g, dtype=torch.float64)
x, y, ls, sv, lb, xs = (t.cuda() for t in (x, y, ls, sv, lb, xs))
for _ in range(a.warmup):
    batched_cigp_eval(x, y, ls, sv, lb, xs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.evals):
    r = batched_cigp_eval(x, y, ls, sv, lb, xs)
e1.record()
torch.cuda.synchronize()
import hashlib
print('ms per eval', e0.elapsed_time(e1) / a.evals, 'nll checksum', float(r['nll'].sum()), 'input checksum', float(x.sum() + y.sum() + xs.sum()),
      'y sha', hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:12])   # y comes from a CPU matmul (MKL: run-to-run alignment-dependent rounding)
