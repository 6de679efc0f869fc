"""Golden vectors for the discrete multi-fidelity acquisition functions, produced by the UNMODIFIED reference class
(MF_BayesianOptimization/Discrete/DMF_acq.py:15-166: UCB_MF, EI_MF with its scipy.stats.norm host round trip and float32
rounding, PI_MF in its log-density form) -> tests/golden/acq.npz.   TEST INFRASTRUCTURE.   python oracle/gen_golden_acq.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
import importlib.util  # noqa: E402

# the package __init__ of MF_BayesianOptimization/Discrete has stale imports (`from v1.CFKG import ...`): load the
# UNMODIFIED module file itself
_spec = importlib.util.spec_from_file_location('DMF_acq_ref', os.path.join(REF, 'MF_BayesianOptimization', 'Discrete', 'DMF_acq.py'))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
DiscreteAcquisitionFunction = _mod.DiscreteAcquisitionFunction

g = torch.Generator().manual_seed(2024)
m = 257
mean = torch.randn(m, 1, generator=g) * 1.5
var = torch.rand(m, 1, generator=g) * 2.0
var[:6, 0] = torch.tensor([0.0, 1e-20, 1e-18, 1e-12, 4.0, 1e-30])      # around the clamp(std, min=1e-9) of :97 / :121
f_best, xdim = 0.3, 7
out = {'mean': mean.numpy(), 'var': var.numpy(), 'f_best': np.array(f_best), 'x_dimension': np.array(xdim)}
for kind in ('UCB', 'EI', 'PI'):
    mu = mean.clone().requires_grad_(True)
    v = var.clone().requires_grad_(True)
    acq = DiscreteAcquisitionFunction(lambda x, s: mu, lambda x, s: v, fidelity_num=2, x_dimension=xdim, f_best=torch.tensor(f_best))
    score = getattr(acq, kind + '_MF')(None, 0)
    score.sum().backward()
    out[kind + '_score'] = score.detach().numpy()
    out[kind + '_dmean'] = mu.grad.numpy()
    out[kind + '_dvar'] = v.grad.numpy()
    print(kind, score.dtype, tuple(score.shape), float(score.sum()))
np.savez_compressed(os.path.join(HERE, '..', 'tests', 'golden', 'acq.npz'), **out)
print('wrote acq.npz')
