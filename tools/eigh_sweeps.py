"""Sweeps and time of the one-sided Jacobi eigensolver on the C4 mode matrices (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200 import tensorly_compat as tl, _lib
g = torch.Generator().manual_seed(4)
for n in (16, 32, 64, 128):
    x = torch.randn(n, 5, generator=g).cuda()
    K = torch.exp(-0.5 * torch.cdist(x, x) ** 2)
    tl._eigh_launch(K); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tl._eigh_launch(K); e1.record(); torch.cuda.synchronize()
    sw = _lib.lib().ffgp_debug_last_eigh_sweeps()
    us = e0.elapsed_time(e1) * 1e3
    steps = sw * (n - 1)
    print(f'n={n:4d}: {sw} sweeps, {us:8.1f} us, {us / max(steps, 1):.3f} us/step')
