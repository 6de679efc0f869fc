"""Launch timeline of ONE warm C2 evaluation (FFGP_TRACE=1): completion time of every level-3 launch per stream.
  FFGP_TRACE=1 python tools/timeline_c2.py > gpurun_out/timeline.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200 import _lib
from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator().manual_seed(0)
x = torch.randn(n, 16, generator=g); y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn(n, 1, generator=g)
x, y = x.cuda(), y.cuda()
m = cigp(ARDKernel(16), 1.0).cuda()
L = _lib.lib()
for i in range(3):
    m.zero_grad(); loss = -m.negative_log_likelihood(x, y); loss.backward()
    torch.cuda.synchronize()
    if i < 2:
        os.environ['FFGP_TRACE_QUIET'] = '1'
        import contextlib, io
        # discard the warm-up timelines
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
        L.ffgp_trace_dump()
        os.dup2(saved, 1); os.close(devnull)
    else:
        sys.stdout.flush()
        L.ffgp_trace_dump()
