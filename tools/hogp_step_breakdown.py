"""HOGP (C4 shape 128 x 32 x 32 x 16) loss+grad step: eager vs CUDA-graph replay, and the per-mode eigensolver times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200 import tensorly_compat as tl
from fidelityfusion_b200.MFGP_ver2023May import HOGP
from fidelityfusion_b200.training import GraphedTrainer
g = torch.Generator().manual_seed(4)
x = torch.rand(128, 5, generator=g).cuda(); Y = torch.randn(128, 32, 32, 16, generator=g).cuda()
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for n in (16, 32, 128):
    xx = torch.rand(n, 3, generator=g).cuda(); K = torch.exp(-0.5 * torch.cdist(xx, xx) ** 2)
    print(f'eigh n={n}: {timed(lambda: tl.eigh(K)):.3f} ms')
h = HOGP({'fidelity_shapes': [torch.Size([32, 32, 16])]}).double().cuda()
def step():
    h.zero_grad(set_to_none=True); h.compute_loss(x, Y).backward()
print(f'eager loss+grad step: {timed(step):.3f} ms')
h2 = HOGP({'fidelity_shapes': [torch.Size([32, 32, 16])]}).double().cuda()
tr = GraphedTrainer(lambda: h2.compute_loss(x, Y), h2.parameters(), lr=1e-3, history=64)
print(f'graph replay epoch (loss+grad+Adam): {timed(lambda: tr.run(1, check=False)):.3f} ms')
tr.check()
