"""Golden vectors for the single-fidelity acquisition classes, produced by the UNMODIFIED reference module
(Bayesian_optimization/acq.py:118-294: UCB, EI and PI with their scipy.stats.norm host round trip and float32 cdf / pdf
tensors, PF) -> tests/golden/acq_sf.npz.   TEST INFRASTRUCTURE.   python oracle/gen_golden_acq_sf.py"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
_spec = importlib.util.spec_from_file_location('acq_sf_ref', os.path.join(REF, 'Bayesian_optimization', 'acq.py'))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)

g = torch.Generator().manual_seed(2025)
m = 301
mean = torch.randn(m, 1, generator=g) * 1.5
var = torch.rand(m, 1, generator=g) * 2.0
var[:6, 0] = torch.tensor([0.0, 1e-20, 1e-18, 1e-12, 4.0, 1e-30])      # around the clamp(std, min=1e-9) of :175 / :224
f_best, kappa, xi = 0.3, 2.5, 0.02
out = {'mean': mean.numpy(), 'var': var.numpy(), 'f_best': np.array(f_best), 'kappa': np.array(kappa), 'xi': np.array(xi)}
for kind in ('UCB', 'EI'):
    mu = mean.clone().requires_grad_(True)
    v = var.clone().requires_grad_(True)
    if kind == 'UCB':
        score = _mod.UCB(lambda X: mu, lambda X: v, kappa=kappa).forward(None)
    else:
        score = _mod.EI(lambda X: mu, lambda X: v, xi=xi).forward(None, f_best)
    score.sum().backward()
    out[kind + '_score'] = score.detach().numpy()
    out[kind + '_dmean'] = mu.grad.numpy()
    out[kind + '_dvar'] = v.grad.numpy()
    print(kind, score.dtype, tuple(score.shape), float(score[6:].sum()))
pi = _mod.PI(lambda X: mean, lambda X: var, sita=xi).forward(None, f_best)          # no autograd: Z.numpy() (:230)
out['PI_score'] = pi.numpy()
print('PI', pi.dtype, float(pi.sum()))
mu3 = torch.randn(m, 3, generator=g)
v3 = torch.rand(m, 3, generator=g) + 0.05
thr = [0.5, -0.2, 1.0]
pf = _mod.PF(lambda X: mu3, lambda X: v3, thr).forward(torch.zeros(m, 2))
out.update(pf_mean=mu3.numpy(), pf_var=v3.numpy(), pf_thresholds=np.array(thr), PF_score=np.asarray(pf))
print('PF', np.asarray(pf).dtype, float(np.asarray(pf).sum()))
np.savez_compressed(os.path.join(HERE, '..', 'tests', 'golden', 'acq_sf.npz'), **out)
print('wrote acq_sf.npz')
