"""Golden vectors for FIDES + Kernel_res from the UNMODIFIED reference
(MFGP_ver2023May/base_gp/fides.py, kernel/MCMC_res_kernel.py), under the stub modules of oracle/_ref_stubs.py.
TEST INFRASTRUCTURE; run once in the build container:  python oracle/gen_golden_fides.py -> tests/golden/fides2023.npz"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'fides2023.npz')
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
import _ref_stubs  # noqa: E402

_ref_stubs.install()
import torch  # noqa: E402

warnings.filterwarnings('ignore')
torch.set_default_dtype(torch.float64)
with contextlib.redirect_stdout(io.StringIO()):
    from MFGP_ver2023May.base_gp.fides import FIDES
    from MFGP_ver2023May.kernel.MCMC_res_kernel import Kernel_res

g = torch.Generator().manual_seed(12)
x = torch.rand(40, 3, generator=g) * 2
y = torch.stack([torch.sin(2 * x.sum(1)), x[:, 0] * x[:, 1] - x[:, 2]], 1) + 0.05 * torch.randn(40, 2, generator=g)
xs = torch.rand(9, 3, generator=g) * 2
out = {'x': x, 'y': y, 'xs': xs}
# (a) the kernel alone, exp format with non-trivial parameters, rectangular
k = Kernel_res(True, [0.7, 1.1, 0.9], 1.4, 0.6).double()
with torch.no_grad():
    k.b.fill_(0.8)
fid = (0.0, 1.0, 0.5, 2.0)
K = k(x, xs, *fid)
out.update(k_exp=K.detach(), k_fid=np.array(fid), k_raw=np.array([float(v) for v in k.length_scale] + [float(k.scale), float(k.length_scale_z), float(k.b)]))
# (b) the model as the reference builds it (create_kernel passes the config dict positionally => linear format, defaults)
m = FIDES({}).double()
m.set_fidelity(0.0, 1.0, 0.0, 2.0)
out['cfg_is_exp'] = np.array(m.kernel.noise_exp_format is True)
with torch.no_grad():
    m.kernel.length_scale.fill_(0.9); m.kernel.scale.fill_(1.3); m.kernel.length_scale_z.fill_(0.7); m.kernel.b.fill_(0.6)
    m.noise_box.value.fill_(0.4)
loss = m.compute_loss(x, y)
loss.backward()
with torch.no_grad():
    u, v = m.forward(xs)
out.update(loss=loss.detach(), u=u, var=v, fid=np.array([0.0, 1.0, 0.0, 2.0]),
           g_noise=m.noise_box.value.grad, g_length_scale=m.kernel.length_scale.grad, g_scale=m.kernel.scale.grad,
           g_length_scale_z=m.kernel.length_scale_z.grad, g_b=m.kernel.b.grad,
           rng_after=torch.rand(3))          # the global-RNG side effect of the last kernel evaluation (App. A-13)
np.savez_compressed(OUT, **{kk: (vv.detach().cpu().numpy() if isinstance(vv, torch.Tensor) else np.asarray(vv)) for kk, vv in out.items()})
print('wrote', OUT, {kk: np.asarray(vv.detach() if isinstance(vv, torch.Tensor) else vv).shape for kk, vv in out.items()})
