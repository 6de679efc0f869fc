import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200.MFGP_ver2023May import HOGP
g = torch.Generator().manual_seed(4)
x = torch.rand(128, 5, generator=g).cuda(); Y = torch.randn(128, 32, 32, 16, generator=g).cuda()
h = HOGP({'fidelity_shapes': [torch.Size([32, 32, 16])]}).double().cuda()
for _ in range(3):
    h.zero_grad(set_to_none=True); h.compute_loss(x, Y).backward()
torch.cuda.synchronize()
