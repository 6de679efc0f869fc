"""Golden vectors for the kernels outside the north-star path (Linear / Matern / RationalQuadratic /
MaternKernel_scalarLengthScale, GaussianProcess/kernel.py:23-63, 109-169, 275-347), for D > 1 with the non-default
Kinv_methods of gp_computation_pack.Gaussian_log_likelihood (:55-88) and for the hyper-parameter gradient of the
posterior (cigp_v10.py:24-48), all from the UNMODIFIED reference -> tests/golden/kernels2.npz.   TEST INFRASTRUCTURE."""
import contextlib, io, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
sys.path.insert(0, HERE); sys.path.insert(0, REF)
import _ref_stubs
_ref_stubs.install()
import torch
torch.set_default_dtype(torch.float64)
with contextlib.redirect_stdout(io.StringIO()):
    from GaussianProcess import kernel as gpk
    from GaussianProcess import gp_computation_pack as pack
    from GaussianProcess.gp_basic import GP_basic
    from GaussianProcess.cigp_v10 import cigp
g = torch.Generator().manual_seed(31)
x1 = torch.randn(40, 3, generator=g); x2 = torch.randn(33, 3, generator=g)
out = {'x1': x1, 'x2': x2}
ks = {'linear': gpk.LinearKernel(3, 0.8, 1.4), 'matern05': gpk.MaternKernel(3, 0.9, 1.2, nu=0.5), 'matern15': gpk.MaternKernel(3, 0.9, 1.2, nu=1.5, rho=1.3),
      'matern25': gpk.MaternKernel(3, 1.1, 0.7, nu=2.5), 'rq': gpk.RationalQuadraticKernel(0.9, 1.3, 1.7), 'matern_scalar': gpk.MaternKernel_scalarLengthScale(1.2, 0.8, 2.5)}
with torch.no_grad():
    ks['linear'].center.copy_(torch.tensor([0.1, -0.2, 0.3]))
for name, k in ks.items():
    with torch.no_grad():
        out['K_' + name] = k(x1, x2)
# D = 3 with every Kinv_method, value and gradients w.r.t. y and cov
n = 30
x = torch.randn(n, 3, generator=g)
with torch.no_grad():
    S = gpk.ARDKernel(3, 1.1, 0.9)(x, x) + 0.2 * torch.eye(n)
y3 = torch.randn(n, 3, generator=g)
W = torch.randn(3, 3, generator=g)
out.update(gll_x=x, gll_S=S, gll_y=y3, gll_W=W)
for meth in ('cholesky1', 'cholesky2', 'cholesky3', 'direct'):
    yy = y3.clone().requires_grad_(True); SS = S.clone().requires_grad_(True)
    v = pack.Gaussian_log_likelihood(yy, SS, meth)
    (v * W).sum().backward() if v.dim() == 2 else v.backward()
    out[f'gll_{meth}'] = v.detach(); out[f'gll_{meth}_gy'] = yy.grad; out[f'gll_{meth}_gS'] = 0.5 * (SS.grad + SS.grad.T)
yn = torch.randn(n, n, generator=g)
out['gll_yn'] = yn
out['gll_MN1'] = pack.Gaussian_log_likelihood(yn, S, 'torch_distribution_MN1')
k = gpk.ARDKernel(3, 1.1, 0.9)
gp = GP_basic(k, 0.4)
ll = gp.log_likelihood(x, y3, 'cholesky2'); ll.backward()
out.update(gpb_c2=ll.detach(), gpb_c2_g_noise=gp.noise_variance.grad, gpb_c2_g_ls=k.length_scales.grad)
# posterior differentiated w.r.t. the hyper-parameters and y (the reference's forward is a plain autograd expression)
m = cigp(gpk.ARDKernel(3, 0.8, 1.3), 0.7)
xs = torch.randn(7, 3, generator=g); y1 = torch.randn(n, 2, generator=g).requires_grad_(True)
wm = torch.randn(7, 2, generator=g); wc = torch.randn(7, 7, generator=g)
mean, cov = m(x, y1, xs)
((mean * wm).sum() + (cov * wc).sum()).backward()
out.update(post_xs=xs, post_y=y1.detach(), post_wm=wm, post_wc=wc, post_mean=mean.detach(), post_cov=cov.detach(), post_g_y=y1.grad,
           post_g_ls=m.kernel.length_scales.grad, post_g_sv=m.kernel.signal_variance.grad, post_g_lb=m.log_beta.grad)
np.savez_compressed(os.path.join(HERE, '..', 'tests', 'golden', 'kernels2.npz'),
                    **{k_: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k_, v in out.items()})
print('wrote kernels2.npz', sorted(out))
