"""One warm call of the streaming mode product per mode on a 1024 x 32 x 32 x 16 tensor - the command profiled with
ncu --set full (tools/prof13.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200 import tensorly_compat as tl
g = torch.Generator().manual_seed(0)
T = torch.randn(1024, 32, 32, 16, generator=g).cuda()
for mode in (1, 3):
    U = torch.randn(T.shape[mode], T.shape[mode], generator=g).cuda()
    for _ in range(3):
        out = tl._mode_dot_raw(T, U, mode, False)
    G = tl._mode_gram_raw(T, out, mode)
torch.cuda.synchronize()
print('ok', float(out.sum()), float(G.sum()))
