"""GPU parity tests of the API surface around the hot path that round 1 left open (VERDICT r1 "small API holes"):
every Kinv_method of Gaussian_log_likelihood for D > 1 incl. gradients, GP_basic's own 'cholesky2', the posterior
differentiated w.r.t. hyper-parameters and training targets, and the kernels outside the north-star path
(Linear / Matern / RationalQuadratic / MaternKernel_scalarLengthScale).  Golden: tests/golden/kernels2.npz from the
UNMODIFIED reference (oracle/gen_golden_kernels2.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-9


def G(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64).to(DEV)


@pytest.mark.parametrize('meth', ['cholesky1', 'cholesky2', 'cholesky3', 'direct'])
def test_gaussian_log_likelihood_every_method_three_columns(meth):
    from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Gaussian_log_likelihood
    g = load_golden('kernels2')
    y = G(g['gll_y']).requires_grad_(True)
    S = G(g['gll_S']).requires_grad_(True)
    v = Gaussian_log_likelihood(y, S, meth)
    ref = g[f'gll_{meth}']
    assert tuple(v.shape) == tuple(ref.shape)
    assert rel_err(v.detach().cpu(), ref) < TOL
    ((v * G(g['gll_W'])).sum() if v.dim() == 2 else v).backward()
    assert rel_err(y.grad.cpu(), g[f'gll_{meth}_gy']) < TOL
    gS = 0.5 * (S.grad + S.grad.T)                      # the covariance is symmetric: only the symmetric part is defined
    assert rel_err(gS.cpu(), g[f'gll_{meth}_gS']) < TOL


def test_gaussian_log_likelihood_torch_distribution_variants():
    from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Gaussian_log_likelihood
    g = load_golden('kernels2')
    v = Gaussian_log_likelihood(G(g['gll_yn']), G(g['gll_S']), 'torch_distribution_MN1')
    assert rel_err(v.cpu(), g['gll_MN1']) < TOL
    v2 = Gaussian_log_likelihood(G(g['gll_yn']), G(g['gll_S']), 'torch_distribution_MN2')
    assert rel_err(v2.cpu(), g['gll_MN1']) < TOL
    with pytest.raises(ValueError):                     # event size != len(cov): the reference's MultivariateNormal raises too
        Gaussian_log_likelihood(G(g['gll_y']), G(g['gll_S']), 'torch_distribution_MN1')


def test_gp_basic_cholesky2_sums_the_matrix():
    from fidelityfusion_b200.GaussianProcess.gp_basic import GP_basic
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    g = load_golden('kernels2')
    k = ARDKernel(3, 1.1, 0.9)
    gp = GP_basic(k, 0.4).to(DEV)
    ll = gp.log_likelihood(G(g['gll_x']), G(g['gll_y']), 'cholesky2')
    assert rel_err(ll.detach().cpu().reshape(-1), g['gpb_c2'].reshape(-1)) < TOL
    ll.backward()
    assert rel_err(gp.noise_variance.grad.cpu(), g['gpb_c2_g_noise']) < TOL
    assert rel_err(k.length_scales.grad.cpu(), g['gpb_c2_g_ls']) < TOL


def test_posterior_gradient_wrt_hyper_parameters_and_targets():
    """cigp.forward with differentiable_posterior=True is the reference's autograd expression (cigp_v10.py:24-48):
    gradients of a functional of (mean, cov) w.r.t. length scales, signal variance, log_beta and y."""
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    g = load_golden('kernels2')
    m = cigp(ARDKernel(3, 0.8, 1.3), 0.7).to(DEV)
    m.differentiable_posterior = True
    y = G(g['post_y']).requires_grad_(True)
    mean, cov = m(G(g['gll_x']), y, G(g['post_xs']))
    assert rel_err(mean.detach().cpu(), g['post_mean']) < TOL and rel_err(cov.detach().cpu(), g['post_cov']) < TOL
    ((mean * G(g['post_wm'])).sum() + (cov * G(g['post_wc'])).sum()).backward()
    assert rel_err(y.grad.cpu(), g['post_g_y']) < TOL
    assert rel_err(m.kernel.length_scales.grad.cpu(), g['post_g_ls']) < TOL
    assert rel_err(m.kernel.signal_variance.grad.cpu(), g['post_g_sv']) < TOL
    assert rel_err(m.log_beta.grad.cpu(), g['post_g_lb']) < TOL
    # the default (fast) path gives the same posterior without tracking the hyper-parameters
    m2 = cigp(ARDKernel(3, 0.8, 1.3), 0.7).to(DEV)
    mean2, cov2 = m2(G(g['gll_x']), G(g['post_y']), G(g['post_xs']))
    assert rel_err(mean2.detach().cpu(), g['post_mean']) < TOL and rel_err(cov2.detach().cpu(), g['post_cov']) < TOL
    assert not mean2.requires_grad


@pytest.mark.parametrize('name,tol', [('linear', 1e-12), ('rq', 1e-12), ('matern_scalar', 1e-7), ('matern05', 1e-7),
                                      ('matern15', 1e-7), ('matern25', 1e-9)])
def test_kernels_outside_the_north_star_path(name, tol):
    """Same formulas on the device.  The Matern family takes sqrt of a squared distance that cancels to ~1e-16 on
    near-coincident points; torch.cdist's own rounding there is 1e-8 of the kernel value, hence the 1e-7 bar."""
    from fidelityfusion_b200.GaussianProcess import kernel as K
    g = load_golden('kernels2')
    ks = {'linear': lambda: K.LinearKernel(3, 0.8, 1.4), 'matern05': lambda: K.MaternKernel(3, 0.9, 1.2, nu=0.5),
          'matern15': lambda: K.MaternKernel(3, 0.9, 1.2, nu=1.5, rho=1.3), 'matern25': lambda: K.MaternKernel(3, 1.1, 0.7, nu=2.5),
          'rq': lambda: K.RationalQuadraticKernel(0.9, 1.3, 1.7), 'matern_scalar': lambda: K.MaternKernel_scalarLengthScale(1.2, 0.8, 2.5)}
    k = ks[name]()
    if name == 'linear':
        with torch.no_grad():
            k.center.copy_(torch.tensor([0.1, -0.2, 0.3]))
    k = k.to(DEV)
    out = k(G(g['x1']), G(g['x2']))
    assert rel_err(out.detach().cpu(), g['K_' + name]) < tol
    out.sum().backward()                                # autograd flows into the parameters
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in k.parameters() if p.requires_grad)
