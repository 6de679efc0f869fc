"""One (or a few) C2 NLL+grad evaluations through the public module API - the command profiled with ncu.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_c2.py --evals 2
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ap = argparse.ArgumentParser()
ap.add_argument('--evals', type=int, default=2)
ap.add_argument('--n', type=int, default=8192)
ap.add_argument('--d', type=int, default=16)
ap.add_argument('--batch', type=int, default=0, help='>0: batched C5-style problems of size n instead of one problem')
a = ap.parse_args()
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
from fidelityfusion_b200.batched import batched_cigp_eval
g = torch.Generator().manual_seed(0)
if a.batch == 0:
    x = torch.randn(a.n, a.d, generator=g); y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn(a.n, 1, generator=g)
    x, y = x.cuda(), y.cuda()
    m = cigp(ARDKernel(a.d), 1.0).cuda()
    for _ in range(a.evals):
        m.zero_grad(); loss = -m.negative_log_likelihood(x, y); loss.backward()
    torch.cuda.synchronize(); print('nll', loss.item())
else:
    x = torch.rand(a.batch, a.n, a.d, generator=g); y = torch.randn(a.batch, a.n, 1, generator=g)
    ls = torch.ones(a.batch, a.d); sv = torch.ones(a.batch); lb = torch.ones(a.batch)
    for _ in range(a.evals):
        out = batched_cigp_eval(x.cuda(), y.cuda(), ls.cuda(), sv.cuda(), lb.cuda())
    torch.cuda.synchronize(); print('nll sum', out['nll'].sum().item())
