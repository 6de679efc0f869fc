"""FIDES + Kernel_res (reference MFGP_ver2023May/base_gp/fides.py, kernel/MCMC_res_kernel.py).
CPU: the oracle restatement against golden vectors of the unmodified reference (oracle/gen_golden_fides.py).
GPU: the drop-in modules (fused kernel assembly / NLL / posterior through the C ABI) against the same vectors."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'fides2023.npz')
TOL = 1e-9      # north_star: 1e-9 relative in fp64


def gold():
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(GOLD).items()}


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64).cpu(), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def test_oracle_kernel_res_and_fides_match_the_reference():
    from oracle import ff_oracle as O
    g = gold()
    x, y, xs = g['x'], g['y'], g['xs']
    raw = g['k_raw']
    K = O.kernel_res(x, xs, raw[:3], raw[3], raw[4], raw[5], *[float(v) for v in g['k_fid']], exp_format=True)
    assert rel(K, g['k_exp']) < 1e-13
    assert not bool(g['cfg_is_exp'])               # create_kernel passes the config dict: linear format (quirk reproduced)
    p = {n: torch.tensor(v, dtype=torch.float64, requires_grad=True) for n, v in
         (('ls', 0.9), ('sc', 1.3), ('lz', 0.7), ('b', 0.6), ('noise', 0.4))}
    fid = [float(v) for v in g['fid']]
    kern = lambda a, b_: O.kernel_res(a, b_, p['ls'], p['sc'], p['lz'], p['b'], *fid)
    loss = O.CIGP_loss(kern(x, x), torch.exp(p['noise']), y)
    loss.backward()
    assert rel(loss.detach(), g['loss']) < 1e-13
    for name, key in (('ls', 'g_length_scale'), ('sc', 'g_scale'), ('lz', 'g_length_scale_z'), ('b', 'g_b'), ('noise', 'g_noise')):
        assert rel(p[name].grad, g[key]) < 1e-11, key
    with torch.no_grad():
        u, v = O.FIDES_predict(kern(x, x), kern(x, xs), kern(xs, xs).diag(), torch.exp(p['noise']), y)
    assert rel(u, g['u']) < 1e-12 and rel(v, g['var']) < 1e-12 and v.shape == g['var'].shape
    assert torch.equal(torch.rand(3, dtype=torch.float64), g['rng_after'])     # the kernel reseeded the global RNG (A-13)


@pytest.mark.gpu
def test_gpu_kernel_res_matches_the_reference():
    from fidelityfusion_b200.MFGP_ver2023May.kernel.MCMC_res_kernel import Kernel_res
    g = gold()
    k = Kernel_res(True, [0.7, 1.1, 0.9], 1.4, 0.6).double().cuda()
    with torch.no_grad():
        k.b.fill_(0.8)
    K = k(g['x'].cuda(), g['xs'].cuda(), *[float(v) for v in g['k_fid']])
    assert rel(K.detach(), g['k_exp']) < TOL


@pytest.mark.gpu
def test_gpu_fides_loss_gradients_and_posterior_match_the_reference():
    from fidelityfusion_b200.MFGP_ver2023May import FIDES
    g = gold()
    x, y, xs = g['x'].cuda(), g['y'].cuda(), g['xs'].cuda()
    m = FIDES({}).double().cuda()
    assert m.kernel.noise_exp_format is not True
    m.set_fidelity(*[float(v) for v in g['fid']])
    with torch.no_grad():
        m.kernel.length_scale.fill_(0.9); m.kernel.scale.fill_(1.3); m.kernel.length_scale_z.fill_(0.7); m.kernel.b.fill_(0.6)
        m.noise_box.value.fill_(0.4)
    assert m.forward(xs) is None                                    # not trained yet (fides.py:88-90)
    loss = m.compute_loss(x, y)
    loss.backward()
    assert rel(loss.detach(), g['loss']) < TOL
    for prm, key in ((m.kernel.length_scale, 'g_length_scale'), (m.kernel.scale, 'g_scale'),
                     (m.kernel.length_scale_z, 'g_length_scale_z'), (m.kernel.b, 'g_b'), (m.noise_box.value, 'g_noise')):
        assert rel(prm.grad, g[key]) < TOL, key
    u, v = m.forward(xs)
    assert rel(u, g['u']) < TOL and rel(v, g['var']) < TOL and tuple(v.shape) == tuple(g['var'].shape)
    assert torch.equal(torch.rand(3, dtype=torch.float64), g['rng_after'])
    # new fidelity bounds invalidate the resident factorisation
    m.set_fidelity(0.0, 1.0, 0.0, 1.0)
    u2, _ = m.forward(xs)
    assert not torch.equal(u2, u)
