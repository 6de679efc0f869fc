"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE.  Run once in the build container (the reference does not travel
to the GPU box):   python oracle/gen_golden.py
Inputs are deterministic (closed form or seeded torch.Generator, fp64) and are stored
next to the reference's outputs so nothing has to be regenerated at test time.
Harness convention of SURVEY.md section 8(c): default dtype fp64, `model.double()`.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '..', 'tests', 'golden')
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import _ref_stubs  # noqa: E402

_ref_stubs.install()
import torch  # noqa: E402

warnings.filterwarnings('ignore')
torch.set_default_dtype(torch.float64)
with contextlib.redirect_stdout(io.StringIO()):
    from GaussianProcess.cigp_v10 import cigp
    from GaussianProcess import kernel as gpk
    from GaussianProcess import gp_computation_pack as pack
    from GaussianProcess.gp_basic import GP_basic
    from GaussianProcess.hogp_simple import HOGP_simple as HOGP_simple_gp
    from FidelityFusion_Models.two_fidelity_models.hogp_simple import HOGP_simple as HOGP_simple_ffm
    import MFGP_ver2023May as G23
    from MFGP_ver2023May.kernel.SE_kernel import SE_kernel
    from MFGP_ver2023May.multiscale_coupling.matrix import Matrix_Mapping
    from MFGP_ver2023May.multiscale_coupling.Residual import Residual


def npy(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy().copy()
    return np.asarray(t)


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **{k: npy(v) for k, v in arrs.items()})
    print('wrote', name, {k: tuple(npy(v).shape) for k, v in arrs.items()})


def grads(model):
    return {n: p.grad for n, p in model.named_parameters() if p.grad is not None}


def kat_inputs():
    N, d, Ns = 8, 2, 3
    i = torch.arange(N * d, dtype=torch.float64).reshape(N, d)
    x = torch.sin(1 + i)
    y = torch.stack([torch.cos(x.sum(1)), x[:, 0] * x[:, 1]], 1)
    xs = torch.sin(0.5 + 0.7 * torch.arange(Ns * d, dtype=torch.float64).reshape(Ns, d))
    return x, y, xs


# ---------------------------------------------------------------- kernels
def gen_kernels():
    g = torch.Generator().manual_seed(11)
    out = {}
    for tag, n1, n2, d in (('small', 10, 7, 3), ('mm', 40, 33, 5)):   # cdist switches to the mm path above 25 rows
        x1 = torch.randn(n1, d, generator=g)
        x2 = torch.randn(n2, d, generator=g)
        ls = torch.rand(d, generator=g) + 0.5
        ls[0] = -ls[0]                      # abs() branch
        k = gpk.ARDKernel(d)
        with torch.no_grad():
            k.length_scales.copy_(ls)
            k.signal_variance.fill_(-1.7)
        out.update({f'{tag}_x1': x1, f'{tag}_x2': x2, f'{tag}_ls': ls, f'{tag}_ard': k(x1, x2), f'{tag}_ard_sym': k(x1, x1)})
        k2 = gpk.SquaredExponentialKernel(0.3, -0.2)
        out[f'{tag}_sqexp'] = k2(x1, x2)
        k3 = SE_kernel(True, [0.7 + 0.1 * i for i in range(d)], 1.3)
        out[f'{tag}_se_exp'] = k3(x1, x2)
        k4 = SE_kernel({'noise_exp_format': True, 'length_scale': 1., 'scale': 1.})   # what kernel_utils.create_kernel does
        out[f'{tag}_se_cfg_is_exp'] = np.array(k4.noise_exp_format is True)
        k5 = SE_kernel(False, 0.8, 2.0)
        out[f'{tag}_se_lin'] = k5(x1, x2)
    save('kernels', **out)


# ---------------------------------------------------------------- KATs
def gen_kats():
    x, y, xs = kat_inputs()
    m = cigp(gpk.ARDKernel(2), 1.0)
    ll = m.negative_log_likelihood(x, y)
    (-ll).backward()
    with torch.no_grad():
        mean, cov = m(x, y, xs)
    save('kat1_cigp_ard', x=x, y=y, xs=xs, ll=ll, g_length_scales=m.kernel.length_scales.grad,
         g_signal_variance=m.kernel.signal_variance.grad, g_log_beta=m.log_beta.grad, mean=mean, cov=cov)

    c = G23.CIGP(None).double()
    loss = c.compute_loss(x, y)
    loss.backward()
    u, v = c.forward(xs)
    save('kat2_CIGP2023', x=x, y=y, xs=xs, loss=loss, g_noise=c.noise_box.value.grad,
         g_length_scale=c.kernel.length_scale.grad, g_scale=c.kernel.scale.grad, u=u, var=v,
         exp_format=np.array(c.kernel.noise_exp_format is True))

    Y = torch.sin(0.37 * torch.arange(48, dtype=torch.float64).reshape(8, 3, 2))
    h = G23.HOGP({'fidelity_shapes': [torch.Size([3, 2])]}).double()
    loss = h.compute_loss(x, Y)
    loss.backward()
    gr = grads(h)
    u, v = h.forward(xs)
    save('kat3_HOGP2023', x=x, Y=Y, xs=xs, loss=loss, u=u, var=v, A=h.A, g=h.g,
         **{'g_' + k.replace('.', '_'): val for k, val in gr.items()})


# ---------------------------------------------------------------- C2-shaped dense NLL+grad
def gen_c2():
    for N in (512,):
        g = torch.Generator().manual_seed(0)
        x = torch.randn(N, 16, generator=g)
        y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn(N, 1, generator=g)
        out = {'x': x, 'y': y}
        cases = [('init', 1.0, 1.0, 1.0), ('ls_half', 0.5, 1.0, 1.0), ('ls_double', 2.0, 1.0, 1.0),
                 ('lb_m2', 1.0, 1.0, -2.0), ('lb_3', 1.0, 1.0, 3.0), ('sv_neg', 1.3, -0.6, 0.5)]
        for tag, ls, sv, lb in cases:
            m = cigp(gpk.ARDKernel(16, ls, sv), lb)
            yy = y.clone().requires_grad_(True)
            ll = m.negative_log_likelihood(x, yy)
            (-ll).backward()
            out.update({f'{tag}_params': np.array([ls, sv, lb]), f'{tag}_ll': ll,
                        f'{tag}_g_length_scales': m.kernel.length_scales.grad,
                        f'{tag}_g_signal_variance': m.kernel.signal_variance.grad,
                        f'{tag}_g_log_beta': m.log_beta.grad, f'{tag}_g_y': yy.grad})
        xs = torch.randn(24, 16, generator=g)
        m = cigp(gpk.ARDKernel(16, 2.0, 1.0), 1.0)
        with torch.no_grad():
            mean, cov = m(x, y, xs)
        out.update({'xs': xs, 'pred_mean': mean, 'pred_cov': cov})
        save(f'c2_n{N}', **out)


# ---------------------------------------------------------------- C3-shaped: cigp with y_var, many outputs, Tensor_linear
def gen_c3():
    g = torch.Generator().manual_seed(3)
    N, d, Dl, Dh = 48, 5, 16, 64
    x = torch.rand(N, d, generator=g)
    grid_l = torch.linspace(0, 1, Dl)
    grid_h = torch.linspace(0, 1, Dh)
    amp = 1 + x[:, :1]
    y_low = amp * torch.sin(2 * np.pi * grid_l[None, :] * (1 + x[:, 1:2])) + x[:, 2:3]
    y_high = 1.1 * amp * torch.sin(2 * np.pi * grid_h[None, :] * (1 + x[:, 1:2])) + x[:, 2:3] + 0.1 * torch.cos(3 * grid_h)[None, :]
    B = torch.randn(N, N, generator=g) * 0.05
    y_var = (B @ B.T).abs()
    tl = pack.Tensor_linear([Dl], [Dh])
    m = cigp(gpk.ARDKernel(d, 0.7, 1.2), 0.5)
    res = y_high - tl(y_low)
    ll = m.negative_log_likelihood(x, [res, y_var])
    (-ll).backward()
    xs = torch.rand(8, d, generator=g)
    with torch.no_grad():
        mean, cov = m(x, [res, y_var], xs)
    save('c3_small', x=x, y_low=y_low, y_high=y_high, y_var=y_var, tl_init=tl.vectors[0], res=res, ll=ll,
         g_length_scales=m.kernel.length_scales.grad, g_signal_variance=m.kernel.signal_variance.grad,
         g_log_beta=m.log_beta.grad, g_tl=tl.vectors[0].grad, xs=xs, mean=mean, cov=cov,
         params=np.array([0.7, 1.2, 0.5]))
    # Tensor_linear on a 2-mode tensor: only the LAST mode is applied (SURVEY A-8)
    tl2 = pack.Tensor_linear([4, 6], [8, 12])
    t = torch.randn(5, 4, 6, generator=g)
    save('tensor_linear_2mode', t=t, w0=tl2.vectors[0], w1=tl2.vectors[1], out=tl2(t))


# ---------------------------------------------------------------- functional pack + GP_basic
def gen_pack():
    g = torch.Generator().manual_seed(7)
    N, d, Ns = 30, 3, 6
    x = torch.randn(N, d, generator=g)
    xs = torch.randn(Ns, d, generator=g)
    out = {'x': x, 'xs': xs}
    k = gpk.ARDKernel(d, 1.1, 0.9)
    with torch.no_grad():
        S = k(x, x) + 0.2 * torch.eye(N)
        Ks = k(x, xs)
        Kss = k(xs, xs)
    out.update(Sigma=S, K_s=Ks, K_ss=Kss)
    for D in (1, 3):
        y = torch.randn(N, D, generator=g)
        out[f'y{D}'] = y
        for meth in ('cholesky1', 'cholesky2', 'cholesky3', 'direct'):
            if D > 1 and meth != 'cholesky3':
                continue
            out[f'gll_{meth}_D{D}'] = pack.Gaussian_log_likelihood(y, S, meth)
        for meth in ('cholesky1', 'cholesky3', 'direct'):
            mu, cov = pack.conditional_Gaussian(y, S, Ks, Kss, meth)
            out[f'cg_{meth}_D{D}_mu'] = mu
            out[f'cg_{meth}_D{D}_cov'] = cov
    y = out['y3']
    lb = torch.tensor([0.7], requires_grad=True)
    k = gpk.ARDKernel(d, 1.1, 0.9)
    ll = pack.negative_log_likelihood(k, lb, x, y)
    (-ll).backward()
    out.update(pack_nll=ll, pack_nll_g_lb=lb.grad, pack_nll_g_ls=k.length_scales.grad, pack_nll_g_sv=k.signal_variance.grad)
    # GP_basic
    k = gpk.ARDKernel(d, 1.1, 0.9)
    gp = GP_basic(k, 0.4)
    for D in (1, 3):
        y = out[f'y{D}']
        gp.zero_grad()
        ll = gp.log_likelihood(x, y)
        (-ll.sum()).backward()
        out[f'gpb_ll_D{D}'] = ll
        out[f'gpb_g_noise_D{D}'] = gp.noise_variance.grad.clone()
        out[f'gpb_g_ls_D{D}'] = k.length_scales.grad.clone()
        with torch.no_grad():
            mu, cov = gp(x, y, xs)
        out[f'gpb_mu_D{D}'] = mu
        out[f'gpb_cov_D{D}'] = cov
    save('pack', **out)


# ---------------------------------------------------------------- C1-shaped AR (gen-2023)
def gen_c1():
    g = torch.Generator().manual_seed(1)
    n_all, d, D = 80, 2, 8
    xa = torch.rand(n_all, d, generator=g)
    w = torch.randn(d, D, generator=g)
    ylo = torch.sin(2 * np.pi * xa @ w)
    yhi = 1.2 * ylo + 0.1 * torch.cos(3 * xa[:, :1])
    x, xe = xa[:40], xa[40:]
    y0, y1 = ylo[:40], yhi[:40]
    m = G23.AR({'fidelity_shapes': [(D,), (D,)]}).double()
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    losses = []
    for it in range(5):
        opt.zero_grad()
        loss = m.compute_loss(x, [y0, y1])
        loss.backward()
        if it == 0:
            g0 = {k: v.clone() for k, v in grads(m).items()}
        losses.append(loss.item())
        opt.step()
    u, v = m(xe)
    save('c1_AR2023', x=x, xe=xe, y0=y0, y1=y1, losses=np.array(losses), u=u, var=v,
         **{'g0_' + k.replace('.', '_'): val for k, val in g0.items()},
         **{'p5_' + k.replace('.', '_'): val for k, val in m.named_parameters()})


# ---------------------------------------------------------------- C4-shaped GAR (gen-2023) + HOGP_simple copies
def smooth_field(x, shape, g, scale=1.0):
    grids = [torch.linspace(0, 1, s) for s in shape]
    out = 0
    for r in range(3):
        comp = (1 + x[:, r % x.shape[1]] * (r + 1) * scale)
        for k, gr in enumerate(grids):
            f = torch.sin((r + 1 + k) * np.pi * gr + r)
            comp = comp.unsqueeze(-1) * f
        out = out + comp
    return out


def gen_c4():
    g = torch.Generator().manual_seed(4)
    N, d, shape = 16, 5, (8, 8, 4)
    x = torch.rand(N, d, generator=g)
    Ylo = smooth_field(x, shape, g)
    Yhi = 1.1 * Ylo + 0.05 * smooth_field(x, shape, g, 2.0)
    m = G23.GAR({'fidelity_shapes': [torch.Size(shape)] * 2}).double()
    # make mapping trainable state visible: vectors are Parameters (matrix.py:64)
    loss = m.compute_loss(x, [Ylo, Yhi])
    loss.backward()
    gr = grads(m)
    xs = torch.rand(4, d, generator=g)
    u, v = m(xs)
    save('c4_GAR2023', x=x, Ylo=Ylo, Yhi=Yhi, xs=xs, loss=loss, u=u, var=v,
         **{'g_' + k.replace('.', '_'): val for k, val in gr.items()})

    # one HOGP at non-default params with y-grad (residual training path)
    h = G23.HOGP({'fidelity_shapes': [torch.Size(shape)]}).double()
    with torch.no_grad():
        h.noise_box.value.fill_(3.0)
        for i, k in enumerate(h.kernel_list):
            k.length_scale.fill_(0.2 * (i + 1) - 0.3)
            k.scale.fill_(0.1 * (i + 1))
    Y = Yhi.clone().requires_grad_(True)
    loss = h.compute_loss(x, Y)
    loss.backward()
    u, v = h.forward(xs)
    save('hogp2023_params', x=x, Y=Y, xs=xs, loss=loss, gY=Y.grad, u=u, var=v, A=h.A, g=h.g,
         exp_format=np.array(h.kernel_list[0].noise_exp_format is True),
         **{'g_' + k.replace('.', '_'): val for k, val in grads(h).items()})

    # HOGP_simple, both copies (shared kernel object across modes: SURVEY A-10)
    for tag, cls in (('ffm', HOGP_simple_ffm), ('gp', HOGP_simple_gp)):
        k = gpk.SquaredExponentialKernel(0.2, 0.1)
        hs = cls(k, 2.0, list(shape)).double()
        loss = hs.log_likelihood(x, Yhi)
        loss.backward()
        with torch.no_grad():
            u, v = hs.forward(x, xs)
        save(f'hogp_simple_{tag}', x=x, Y=Yhi, xs=xs, loss=loss, u=u, var=v,
             g_noise=hs.noise_variance.grad, g_length_scale=k.length_scale.grad, g_signal_variance=k.signal_variance.grad)


# ---------------------------------------------------------------- couplings
def gen_couplings():
    g = torch.Generator().manual_seed(9)
    mm = Matrix_Mapping({'low_fidelity_shape': (4, 3), 'high_fidelity_shape': (8, 3), 'matrix_init_method': 'smooth',
                         'rho_value_init': 0.8, 'trainable_rho': True}).double()
    lo = torch.randn(5, 4, 3, generator=g)
    hi = torch.randn(5, 8, 3, generator=g)
    res = mm.forward(lo, hi)
    res.pow(2).sum().backward()
    back = mm.backward(lo, res.detach())
    mm2 = Matrix_Mapping({'low_fidelity_shape': (4,), 'high_fidelity_shape': (8,), 'matrix_init_method': 'eye'}).double()
    r = Residual({'rho_value_init': 0.7}).double()
    save('couplings', lo=lo, hi=hi, w0=mm.vectors[0], w1=mm.vectors[1], rho=mm.rho, res=res, back=back,
         g_w0=mm.vectors[0].grad, g_rho=mm.rho.grad, eye_init=mm2.vectors[0],
         resid_fwd=r.forward(lo, lo * 2), resid_bwd=r.backward(lo, lo * 2))


# ---------------------------------------------------------------- C5-shaped batch of independent small GPs
def gen_c5():
    B, N, d, Ns = 6, 64, 8, 16
    out = {}
    for b in range(B):
        g = torch.Generator().manual_seed(5000 + b)
        x = torch.rand(N, d, generator=g)
        w = torch.randn(d, 1, generator=g)
        y = torch.sin(3 * x @ w) + 0.05 * torch.randn(N, 1, generator=g)
        ls = torch.exp(torch.rand(d, generator=g) * 2 - 1)
        lb = float(torch.rand(1, generator=g) * 3)
        xs = torch.rand(Ns, d, generator=g)
        m = cigp(gpk.ARDKernel(d), lb)
        with torch.no_grad():
            m.kernel.length_scales.copy_(ls)
        ll = m.negative_log_likelihood(x, y)
        (-ll).backward()
        with torch.no_grad():
            mean, cov = m(x, y, xs)
        out.update({f'x{b}': x, f'y{b}': y, f'ls{b}': ls, f'lb{b}': np.array(lb), f'xs{b}': xs, f'll{b}': ll,
                    f'g_ls{b}': m.kernel.length_scales.grad, f'g_sv{b}': m.kernel.signal_variance.grad,
                    f'g_lb{b}': m.log_beta.grad, f'mean{b}': mean, f'vdiag{b}': cov.diag()})
    save('c5_batch', **out)


def gen_predict_dx():
    """Posterior differentiated w.r.t. the test points through the reference's own cigp.forward (autograd), plus the
    ARD kernel differentiated w.r.t. both inputs."""
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, N, d, D, Ns in (('a', 150, 5, 2, 9), ('b', 300, 8, 1, 40)):
        x = torch.rand(N, d, generator=g)
        y = torch.sin(3 * x @ torch.randn(d, D, generator=g)) + 0.05 * torch.randn(N, D, generator=g)
        xs = torch.rand(Ns, d, generator=g).requires_grad_(True)
        ls = torch.exp(torch.rand(d, generator=g) - 0.5)
        m = cigp(gpk.ARDKernel(d), 2.0)
        with torch.no_grad():
            m.kernel.length_scales.copy_(ls)
            m.kernel.signal_variance.fill_(1.3)
        wm = torch.randn(Ns, D, generator=g)
        wc = torch.randn(Ns, Ns, generator=g)
        mean, cov = m(x, y, xs)
        ((mean * wm).sum() + (cov * wc).sum()).backward()
        g_full = xs.grad.clone()
        xs.grad = None
        mean, cov = m(x, y, xs)
        wd = torch.randn(Ns, generator=g)
        ((mean * wm).sum() + (cov.diag() * wd).sum()).backward()
        out.update({f'x_{tag}': x, f'y_{tag}': y, f'xs_{tag}': xs, f'ls_{tag}': ls, f'wm_{tag}': wm, f'wc_{tag}': wc,
                    f'wd_{tag}': wd, f'mean_{tag}': mean, f'cov_{tag}': cov, f'gxs_full_{tag}': g_full,
                    f'gxs_diag_{tag}': xs.grad})
    # kernel matrix w.r.t. both inputs
    x1 = torch.randn(37, 4, generator=g).requires_grad_(True)
    x2 = torch.randn(29, 4, generator=g).requires_grad_(True)
    k = gpk.ARDKernel(4)
    with torch.no_grad():
        k.length_scales.copy_(torch.tensor([0.7, -1.2, 2.0, 0.9]))
        k.signal_variance.fill_(0.8)
    W = torch.randn(37, 29, generator=g)
    (k(x1, x2) * W).sum().backward()
    out.update(kx1=x1, kx2=x2, kW=W, k_ls=k.length_scales, k_sv=k.signal_variance, g_kx1=x1.grad, g_kx2=x2.grad)
    save('predict_dx', **out)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'predict_dx':
        gen_predict_dx()
        sys.exit(0)
    gen_kernels()
    gen_kats()
    gen_c2()
    gen_c3()
    gen_pack()
    gen_c1()
    gen_c4()
    gen_couplings()
    gen_c5()
    gen_predict_dx()
