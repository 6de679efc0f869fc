"""Golden vector for HOGP.compute_loss with a TENSOR-valued y_var (`A = A + y_var`, MFGP_ver2023May/base_gp/hogp.py:176),
produced by the UNMODIFIED reference class -> tests/golden/hogp2023_yvar.npz.
TEST INFRASTRUCTURE.   python oracle/gen_golden_hogp_yvar.py"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
import _ref_stubs  # noqa: E402

_ref_stubs.install()
import torch  # noqa: E402

warnings.filterwarnings('ignore')
torch.set_default_dtype(torch.float64)
with contextlib.redirect_stdout(io.StringIO()):
    import MFGP_ver2023May as G23

g = torch.Generator().manual_seed(176)
N, d, shape = 12, 3, (6, 5, 3)
x = torch.rand(N, d, generator=g)
grid = [torch.linspace(0, 1, s) for s in shape]
Y = torch.sin(3 * x.sum(1)).reshape(N, 1, 1, 1) * torch.cos(2 * grid[0]).reshape(1, -1, 1, 1) \
    * (1 + grid[1]).reshape(1, 1, -1, 1) * torch.exp(-grid[2]).reshape(1, 1, 1, -1) + 0.05 * torch.randn(N, *shape, generator=g)
xs = torch.rand(4, d, generator=g)
out = {}
for tag, yv in (('full', 0.02 + 0.1 * torch.rand(N, *shape, generator=g)),          # one value per element of A
                ('bcast', 0.02 + 0.1 * torch.rand(N, 1, 1, 1, generator=g))):        # per-sample variance, broadcast
    h = G23.HOGP({'fidelity_shapes': [torch.Size(shape)]}).double()
    with torch.no_grad():
        h.noise_box.value.fill_(2.5)
        for i, k in enumerate(h.kernel_list):
            k.length_scale.fill_(-1.2 + 0.25 * i)
            k.scale.fill_(0.1 * (i + 1))
    Yp = Y.clone().requires_grad_(True)
    yvp = yv.clone().requires_grad_(True)
    loss = h.compute_loss(x, Yp, y_var=yvp)
    loss.backward()
    u, v = h.forward(xs)
    out.update({f'{tag}_y_var': yv.numpy(), f'{tag}_loss': loss.detach().numpy(), f'{tag}_gY': Yp.grad.numpy(),
                f'{tag}_g_y_var': yvp.grad.numpy(), f'{tag}_A': h.A.detach().numpy(), f'{tag}_g': h.g.detach().numpy(),
                f'{tag}_u': u.detach().numpy(), f'{tag}_var': v.detach().numpy(),
                f'{tag}_g_noise': h.noise_box.value.grad.numpy()})
    for i, k in enumerate(h.kernel_list):
        out[f'{tag}_g_ls{i}'] = k.length_scale.grad.numpy()
        out[f'{tag}_g_sc{i}'] = k.scale.grad.numpy()
    print(tag, float(loss), [float(k.length_scale.grad) for k in h.kernel_list])
out.update(x=x.numpy(), Y=Y.numpy(), xs=xs.numpy())
np.savez_compressed(os.path.join(HERE, '..', 'tests', 'golden', 'hogp2023_yvar.npz'), **out)
print('wrote hogp2023_yvar.npz')
