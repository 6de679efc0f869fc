"""ms per C2 evaluation (N=8192, d=16 NLL+grad through the module API), CUDA events, for quick A/B of env knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator().manual_seed(0)
x = torch.randn(n, 16, generator=g); y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn(n, 1, generator=g)
x, y = x.cuda(), y.cuda()
m = cigp(ARDKernel(16), 1.0).cuda()
def ev():
    m.zero_grad(); loss = -m.negative_log_likelihood(x, y); loss.backward(); return loss
for _ in range(3): l = ev()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): l = ev()
e1.record(); torch.cuda.synchronize()
print(os.environ.get('FFGP_SYRK_RESERVE', '0'), 'ms/eval', e0.elapsed_time(e1) / 10, 'nll', float(l))
