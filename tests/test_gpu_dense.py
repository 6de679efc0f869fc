"""GPU parity tests of the dense GP path: CUDA (through the drop-in modules -> C ABI) vs the CPU oracle on the
same inputs, vs the committed golden vectors of the real reference, and size-independent properties at the
BASELINE.json sizes.  Tolerance: 1e-9 relative (north_star), fp32 inputs 1e-4."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ff_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9
DEV = 'cuda'


def T(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def G(a):
    return T(a).to(DEV)


@pytest.fixture(autouse=True)
def _f64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _cigp(d, ls, sv, lb):
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    k = ARDKernel(d)
    with torch.no_grad():
        k.length_scales.copy_(T(ls))
        k.signal_variance.fill_(float(sv))
    return cigp(k, float(lb)).to(DEV)


def test_kernel_matrices_match_reference_golden():
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel, SquaredExponentialKernel
    from fidelityfusion_b200.MFGP_ver2023May.kernel.SE_kernel import SE_kernel
    g = load_golden('kernels')
    for tag in ('small', 'mm'):
        x1, x2 = G(g[f'{tag}_x1']), G(g[f'{tag}_x2'])
        d = x1.shape[1]
        k = ARDKernel(d)
        with torch.no_grad():
            k.length_scales.copy_(T(g[f'{tag}_ls']))
            k.signal_variance.fill_(-1.7)
        k = k.to(DEV)
        assert rel_err(k(x1, x2).cpu(), g[f'{tag}_ard']) < TOL
        assert rel_err(k(x1, x1).cpu(), g[f'{tag}_ard_sym']) < TOL
        assert rel_err(SquaredExponentialKernel(0.3, -0.2).to(DEV)(x1, x2).cpu(), g[f'{tag}_sqexp']) < TOL
        assert rel_err(SE_kernel(True, [0.7 + 0.1 * i for i in range(d)], 1.3).to(DEV)(x1, x2).cpu(), g[f'{tag}_se_exp']) < TOL
        assert rel_err(SE_kernel(False, 0.8, 2.0).to(DEV)(x1, x2).cpu(), g[f'{tag}_se_lin']) < TOL


def test_kernel_matrix_autograd_matches_oracle():
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    gen = torch.Generator().manual_seed(2)
    x1, x2 = torch.randn(70, 6, generator=gen), torch.randn(45, 6, generator=gen)
    wgt = torch.randn(70, 45, generator=gen)
    k = ARDKernel(6, 0.8, 1.4).to(DEV)
    (k(x1.to(DEV), x2.to(DEV)) * wgt.to(DEV)).sum().backward()
    ls, sv = (torch.ones(6) * 0.8).requires_grad_(True), torch.tensor([1.4], requires_grad=True)
    (O.ard_kernel(x1, x2, ls, sv) * wgt).sum().backward()
    assert rel_err(k.length_scales.grad.cpu(), ls.grad) < TOL
    assert rel_err(k.signal_variance.grad.cpu(), sv.grad) < TOL


def test_kat1_cigp_ard():
    g = load_golden('kat1_cigp_ard')
    m = _cigp(2, [1., 1.], 1.0, 1.0)
    x, y, xs = G(g['x']), G(g['y']), G(g['xs'])
    ll = m.negative_log_likelihood(x, y)
    assert ll.dim() == 0
    assert abs(ll.item() - (-14.408784462236161)) < 1e-11
    (-ll).backward()
    assert rel_err(m.kernel.length_scales.grad.cpu(), g['g_length_scales']) < TOL
    assert rel_err(m.kernel.signal_variance.grad.cpu(), g['g_signal_variance']) < TOL
    assert rel_err(m.log_beta.grad.cpu(), g['g_log_beta']) < TOL
    mean, cov = m(x, y, xs)
    assert rel_err(mean.cpu(), g['mean']) < TOL and rel_err(cov.cpu(), g['cov']) < TOL
    mean2, cov2 = m(x, y, xs)                                          # second call reuses the cached factor
    assert torch.equal(mean, mean2) and torch.equal(cov, cov2)


def test_kat2_CIGP2023():
    from fidelityfusion_b200.MFGP_ver2023May import CIGP
    g = load_golden('kat2_CIGP2023')
    c = CIGP(None).double().to(DEV)
    assert sorted(n for n, _ in c.named_parameters()) == ['kernel.length_scale', 'kernel.scale', 'noise_box.value']
    x, y, xs = G(g['x']), G(g['y']), G(g['xs'])
    loss = c.compute_loss(x, y)
    assert abs(loss.item() - 19.35232673878598) < 1e-10
    loss.backward()
    assert rel_err(c.noise_box.value.grad.cpu(), g['g_noise']) < TOL
    assert rel_err(c.kernel.length_scale.grad.cpu(), g['g_length_scale']) < TOL
    assert rel_err(c.kernel.scale.grad.cpu(), g['g_scale']) < TOL
    u, v = c.forward(xs)
    assert rel_err(u.cpu(), g['u']) < TOL and rel_err(v.cpu(), g['var']) < TOL


@pytest.mark.parametrize('tag', ['init', 'ls_half', 'ls_double', 'lb_m2', 'lb_3', 'sv_neg'])
def test_c2_shape_nll_and_every_gradient(tag):
    g = load_golden('c2_n512')
    ls0, sv0, lb0 = g[f'{tag}_params']
    m = _cigp(16, np.full(16, ls0), sv0, lb0)
    x = G(g['x'])
    y = G(g['y']).requires_grad_(True)
    ll = m.negative_log_likelihood(x, y)
    assert abs(ll.item() - float(g[f'{tag}_ll'])) <= TOL * abs(float(g[f'{tag}_ll']))
    (-ll).backward()
    assert rel_err(m.kernel.length_scales.grad.cpu(), g[f'{tag}_g_length_scales']) < TOL
    assert rel_err(m.kernel.signal_variance.grad.cpu(), g[f'{tag}_g_signal_variance']) < TOL
    assert rel_err(m.log_beta.grad.cpu(), g[f'{tag}_g_log_beta']) < TOL
    assert rel_err(y.grad.cpu(), g[f'{tag}_g_y']) < TOL


def test_c2_shape_predict():
    g = load_golden('c2_n512')
    m = _cigp(16, np.full(16, 2.0), 1.0, 1.0)
    mean, cov = m(G(g['x']), G(g['y']), G(g['xs']))
    assert rel_err(mean.cpu(), g['pred_mean']) < TOL and rel_err(cov.cpu(), g['pred_cov']) < TOL


@pytest.mark.parametrize('n,d,D', [(1, 1, 1), (7, 3, 2), (129, 4, 1), (300, 2, 9), (257, 5, 70)])
def test_ragged_sizes_vs_oracle(n, d, D):
    """sizes that are not multiples of any tile (identity-tail padding), D on both sides of the GEMV/GEMM switch."""
    gen = torch.Generator().manual_seed(100 + n)
    x, y = torch.rand(n, d, generator=gen), torch.randn(n, D, generator=gen)
    ls, sv, lb = torch.rand(d, generator=gen) + 0.5, torch.tensor([1.3]), torch.tensor([0.7])
    m = _cigp(d, ls, sv, lb)
    yy = y.to(DEV).requires_grad_(True)
    ll = m.negative_log_likelihood(x.to(DEV), yy)
    (-ll).backward()
    loss, gr = O.cigp_ard_nll_and_grads(x, y, ls, sv, lb, want_y_grad=True)
    assert abs(-ll.item() - loss) <= TOL * max(abs(loss), 1.0)
    assert rel_err(m.kernel.length_scales.grad.cpu(), gr['length_scales']) < TOL
    assert rel_err(m.kernel.signal_variance.grad.cpu(), gr['signal_variance']) < TOL
    assert rel_err(m.log_beta.grad.cpu(), gr['log_beta']) < TOL
    assert rel_err(yy.grad.cpu(), gr['y']) < TOL
    xs = torch.rand(5, d, generator=gen)
    mean, cov = m(x.to(DEV), y.to(DEV), xs.to(DEV))
    om, oc = O.cigp_ard_predict(x, y, xs, ls, sv, lb)
    assert rel_err(mean.cpu(), om) < TOL and rel_err(cov.cpu(), oc) < TOL


def test_c3_yvar_many_outputs_tensor_linear():
    from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Tensor_linear
    g = load_golden('c3_small')
    tl = Tensor_linear([16], [64]).double().to(DEV)
    assert rel_err(tl.vectors[0].detach().cpu(), g['tl_init']) < 1e-12
    m = _cigp(5, np.full(5, 0.7), 1.2, 0.5)
    x, ylo, yhi, yv = G(g['x']), G(g['y_low']), G(g['y_high']), G(g['y_var'])
    res = yhi - tl(ylo)
    assert rel_err(res.detach().cpu(), g['res']) < TOL
    ll = m.negative_log_likelihood(x, [res, yv])
    assert abs(ll.item() - float(g['ll'])) <= TOL * abs(float(g['ll']))
    (-ll).backward()
    assert rel_err(m.kernel.length_scales.grad.cpu(), g['g_length_scales']) < TOL
    assert rel_err(m.kernel.signal_variance.grad.cpu(), g['g_signal_variance']) < TOL
    assert rel_err(m.log_beta.grad.cpu(), g['g_log_beta']) < TOL
    assert rel_err(tl.vectors[0].grad.cpu(), g['g_tl']) < TOL        # trained through the residual (CIGAR.py:119)
    mean, cov = m(x, [res.detach(), yv], G(g['xs']))
    assert rel_err(mean.cpu(), g['mean']) < TOL and rel_err(cov.cpu(), g['cov']) < TOL
    g2 = load_golden('tensor_linear_2mode')
    tl2 = Tensor_linear([4, 6], [8, 12]).double().to(DEV)
    assert rel_err(tl2(G(g2['t'])).detach().cpu(), g2['out']) < 1e-12   # only the last mode is applied (sic)


def test_c3_full_size_residual_step_vs_oracle():
    """BASELINE config 3 at full size: the top CIGAR fidelity (CIGAR.py:119-127) - 64x64 field outputs flattened to
    D = 4096 columns, N = 128 points, d = 5, the 1024 -> 4096 Tensor_linear coupling trained through the residual,
    N x N y_var - against the CPU oracle on the same inputs: log-likelihood, every gradient (4 M coupling weights
    included), posterior mean and covariance."""
    from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Tensor_linear
    gen = torch.Generator().manual_seed(3)
    n, d, Dl, Dh, ns = 128, 5, 1024, 4096, 16
    x = torch.rand(n, d, generator=gen)
    grid_l, grid_h = torch.linspace(0, 1, Dl), torch.linspace(0, 1, Dh)
    amp = 1.0 + x[:, :1]
    y_low = amp * torch.sin(6.0 * grid_l[None, :] + 3.0 * x[:, 1:2]) + 0.3 * torch.cos(9.0 * grid_l[None, :] * x[:, 2:3])
    y_high = 1.1 * (amp * torch.sin(6.0 * grid_h[None, :] + 3.0 * x[:, 1:2])) + 0.05 * torch.cos(4.0 * grid_h[None, :] + x[:, 3:4])
    a = torch.rand(n, n, generator=gen)
    y_var = 0.01 * (a @ a.t()) / n
    xs = torch.rand(ns, d, generator=gen)
    ls, sv, lb = np.linspace(0.6, 1.4, d), 1.3, 0.8
    # oracle (CPU)
    w0 = O.tensor_linear_init(Dl, Dh).double()
    wo = w0.clone().requires_grad_(True)
    lso, svo, lbo = T(ls).requires_grad_(True), torch.tensor([sv], requires_grad=True), torch.tensor([lb], requires_grad=True)
    res_o = y_high - O.tensor_linear_forward(y_low, [wo])
    ll_o = O.cigp_log_likelihood(O.ard_kernel(x, x, lso, svo), lbo, res_o, y_var)
    (-ll_o).backward()
    with torch.no_grad():
        mean_o, cov_o = O.cigp_predict(O.ard_kernel(x, x, lso, svo), O.ard_kernel(x, xs, lso, svo),
                                       O.ard_kernel(xs, xs, lso, svo), lbo, res_o.detach())
    # CUDA path
    tl = Tensor_linear([Dl], [Dh]).double().to(DEV)
    assert rel_err(tl.vectors[0].detach().cpu(), w0) < 1e-12
    m = _cigp(d, ls, sv, lb)
    res = G(y_high) - tl(G(y_low))
    assert rel_err(res.detach().cpu(), res_o.detach()) < TOL
    ll = m.negative_log_likelihood(G(x), [res, G(y_var)])
    assert abs(ll.item() - ll_o.item()) <= TOL * abs(ll_o.item())
    (-ll).backward()
    assert rel_err(m.kernel.length_scales.grad.cpu(), lso.grad) < TOL
    assert rel_err(m.kernel.signal_variance.grad.cpu(), svo.grad) < TOL
    assert rel_err(m.log_beta.grad.cpu(), lbo.grad) < TOL
    assert rel_err(tl.vectors[0].grad.cpu(), wo.grad) < TOL
    mean, cov = m(G(x), [res.detach(), G(y_var)], G(xs))
    assert rel_err(mean.cpu(), mean_o) < TOL and rel_err(cov.cpu(), cov_o) < TOL


def test_pack_and_gp_basic():
    from fidelityfusion_b200.GaussianProcess import gp_computation_pack as pack
    from fidelityfusion_b200.GaussianProcess.gp_basic import GP_basic
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    g = load_golden('pack')
    S, Ks, Kss = G(g['Sigma']), G(g['K_s']), G(g['K_ss'])
    for D in (1, 3):
        y = G(g[f'y{D}'])
        for meth in ('cholesky1', 'cholesky2', 'cholesky3', 'direct'):
            key = f'gll_{meth}_D{D}'
            if key in g:
                out = pack.Gaussian_log_likelihood(y, S, meth)
                assert tuple(out.shape) == tuple(g[key].shape)
                assert rel_err(out.cpu(), g[key]) < TOL
        for meth in ('cholesky1', 'cholesky3', 'direct'):
            mu, cov = pack.conditional_Gaussian(y, S, Ks, Kss, meth)
            assert rel_err(mu.cpu(), g[f'cg_{meth}_D{D}_mu']) < TOL and rel_err(cov.cpu(), g[f'cg_{meth}_D{D}_cov']) < TOL
    x, xs, y = G(g['x']), G(g['xs']), G(g['y3'])
    k = ARDKernel(3, 1.1, 0.9).to(DEV)
    lb = torch.tensor([0.7], device=DEV, requires_grad=True)
    ll = pack.negative_log_likelihood(k, lb, x, y)
    assert rel_err(ll.detach().cpu().reshape(()), g['pack_nll']) < TOL
    (-ll).backward()
    assert rel_err(lb.grad.cpu(), g['pack_nll_g_lb']) < TOL
    assert rel_err(k.length_scales.grad.cpu(), g['pack_nll_g_ls']) < TOL
    assert rel_err(k.signal_variance.grad.cpu(), g['pack_nll_g_sv']) < TOL
    for D in (1, 3):
        k = ARDKernel(3, 1.1, 0.9)
        gp = GP_basic(k, 0.4).to(DEV)
        y = G(g[f'y{D}'])
        ll = gp.log_likelihood(x, y)
        assert rel_err(ll.detach().cpu(), g[f'gpb_ll_D{D}']) < TOL
        (-ll.sum()).backward()
        assert rel_err(gp.noise_variance.grad.cpu(), g[f'gpb_g_noise_D{D}']) < TOL
        assert rel_err(gp.kernel.length_scales.grad.cpu(), g[f'gpb_g_ls_D{D}']) < TOL
        mu, cov = gp(x, y, xs)
        assert rel_err(mu.cpu(), g[f'gpb_mu_D{D}']) < TOL and rel_err(cov.cpu(), g[f'gpb_cov_D{D}']) < TOL


def test_c1_AR2023_training_trajectory():
    """C1: gen-2023 AR = sum of CIGP losses on Residual targets (AR_AutoRegression.py:206-254), 5 Adam steps."""
    from fidelityfusion_b200.MFGP_ver2023May import CIGP
    from fidelityfusion_b200.MFGP_ver2023May.multiscale_coupling.Residual import Residual
    g = load_golden('c1_AR2023')
    x, xe, y0, y1 = G(g['x']), G(g['xe']), G(g['y0']), G(g['y1'])
    gps = torch.nn.ModuleList([CIGP(None), CIGP(None)]).double().to(DEV)
    res = Residual(None).double().to(DEV)
    params = list(gps.parameters()) + list(res.parameters())
    opt = torch.optim.Adam(params, lr=0.01)
    for it in range(5):
        opt.zero_grad()
        loss = gps[0].compute_loss(x, y0) + gps[1].compute_loss(x, res.forward(y0, y1), update_data=True)
        loss.backward()
        assert abs(loss.item() - g['losses'][it]) <= 1e-8 * abs(g['losses'][it]), it
        if it == 0:
            assert rel_err(res.rho.grad.cpu(), g['g0_residual_list_0_rho']) < TOL
            for f in range(2):
                assert rel_err(gps[f].kernel.length_scale.grad.cpu(), g[f'g0_cigp_list_{f}_kernel_length_scale']) < TOL
                assert rel_err(gps[f].kernel.scale.grad.cpu(), g[f'g0_cigp_list_{f}_kernel_scale']) < TOL
                assert rel_err(gps[f].noise_box.value.grad.cpu(), g[f'g0_cigp_list_{f}_noise_box_value']) < TOL
        opt.step()
    for f in range(2):
        assert rel_err(gps[f].kernel.length_scale.detach().cpu(), g[f'p5_cigp_list_{f}_kernel_length_scale']) < 1e-8
    # prediction chain of AR.forward (AR_AutoRegression.py:148-177): the residual GP was fitted on the pre-step targets
    m0, v0 = gps[0].forward(xe)
    m1, v1 = gps[1].forward(xe)
    u, var = res.backward(m0, m1), res.var_backward(v0, v1)
    assert rel_err(u.detach().cpu(), g['u']) < 1e-7 and rel_err(var.detach().cpu(), g['var']) < 1e-7


def test_c5_batched_independent_gps():
    from fidelityfusion_b200.batched import batched_cigp_eval
    g = load_golden('c5_batch')
    Bn = 6
    x = torch.stack([T(g[f'x{b}']) for b in range(Bn)]).to(DEV)
    y = torch.stack([T(g[f'y{b}']) for b in range(Bn)]).to(DEV)
    xs = torch.stack([T(g[f'xs{b}']) for b in range(Bn)]).to(DEV)
    ls = torch.stack([T(g[f'ls{b}']) for b in range(Bn)]).to(DEV)
    lb = torch.stack([T(g[f'lb{b}']) for b in range(Bn)]).to(DEV)
    sv = torch.ones(Bn, device=DEV)
    out = batched_cigp_eval(x, y, ls, sv, lb, xs)
    for b in range(Bn):
        assert abs(-out['nll'][b].item() - float(g[f'll{b}'])) <= TOL * abs(float(g[f'll{b}']))
        assert rel_err(out['g_length_scales'][b].cpu(), g[f'g_ls{b}']) < TOL
        assert rel_err(out['g_signal_variance'][b].cpu().reshape(1), g[f'g_sv{b}']) < TOL
        assert rel_err(out['g_log_beta'][b].cpu().reshape(1), g[f'g_lb{b}']) < TOL
        assert rel_err(out['mean'][b].cpu(), g[f'mean{b}']) < TOL
        assert rel_err(out['var'][b].cpu(), g[f'vdiag{b}']) < TOL


def test_batched_eval_asynchronous_status():
    """check=False: no host sync inside the call, the LAPACK-style status comes back with the results; the values are
    those of the checked call, and a non-PD problem (NaN noise) is reported for its batch index either way."""
    from fidelityfusion_b200.batched import batched_cigp_eval, check_batch_info
    gen = torch.Generator().manual_seed(77)
    Bn, n, d, ns = 9, 200, 3, 5
    x = torch.rand(Bn, n, d, generator=gen).to(DEV)
    y = torch.randn(Bn, n, 1, generator=gen).to(DEV)
    xs = torch.rand(Bn, ns, d, generator=gen).to(DEV)
    ls, sv, lb = torch.ones(Bn, d, device=DEV), torch.ones(Bn, device=DEV), torch.ones(Bn, device=DEV)
    a = batched_cigp_eval(x, y, ls, sv, lb, xs)
    b = batched_cigp_eval(x, y, ls, sv, lb, xs, check=False)
    assert 'info' not in a and float(b['info'].abs().sum()) == 0.0
    check_batch_info(b['info'])
    for k in a:
        if k != '_packed':                              # the packed row buffer of b carries the extra status column
            assert torch.equal(a[k], b[k]), k
    assert b['_packed'].shape[1] == a['_packed'].shape[1] + 1
    lb_bad = lb.clone()
    lb_bad[3] = float('nan')
    with pytest.raises(torch.linalg.LinAlgError, match='Batch element 3'):
        batched_cigp_eval(x, y, ls, sv, lb_bad, xs)
    c = batched_cigp_eval(x, y, ls, sv, lb_bad, xs, check=False)
    assert float(c['info'][3]) > 0 and float(c['info'].abs().sum()) == float(c['info'][3])
    with pytest.raises(torch.linalg.LinAlgError, match='Batch element 3'):
        check_batch_info(c['info'])
    assert torch.equal(c['nll'][:3], a['nll'][:3]) and torch.equal(c['nll'][4:], a['nll'][4:])


def test_c5_full_size_batch_vs_oracle_subset():
    """C5 recipe at N=512, d=8: 256 problems in one call, 8 of them checked against the oracle."""
    from fidelityfusion_b200.batched import batched_cigp_eval
    Bn, n, d, ns = 256, 512, 8, 64
    xs_, ys_, lss, lbs, xss = [], [], [], [], []
    for b in range(Bn):
        gen = torch.Generator().manual_seed(5000 + b)
        x = torch.rand(n, d, generator=gen)
        w = torch.randn(d, 1, generator=gen)
        y = torch.sin(3 * x @ w) + 0.05 * torch.randn(n, 1, generator=gen)
        lss.append(torch.exp(torch.rand(d, generator=gen) * 2 - 1))
        lbs.append(torch.rand(1, generator=gen)[0] * 3)
        xss.append(torch.rand(ns, d, generator=gen))
        xs_.append(x); ys_.append(y)
    x, y, xs = torch.stack(xs_), torch.stack(ys_), torch.stack(xss)
    ls, lb, sv = torch.stack(lss), torch.stack(lbs), torch.ones(Bn)
    out = batched_cigp_eval(x.to(DEV), y.to(DEV), ls.to(DEV), sv.to(DEV), lb.to(DEV), xs.to(DEV))
    for b in (0, 1, 17, 63, 64, 128, 200, 255):
        loss, gr = O.cigp_ard_nll_and_grads(x[b], y[b], ls[b], sv[b:b + 1], lb[b:b + 1])
        assert abs(out['nll'][b].item() - loss) <= TOL * abs(loss)
        assert rel_err(out['g_length_scales'][b].cpu(), gr['length_scales']) < TOL
        assert rel_err(out['g_log_beta'][b].cpu().reshape(1), gr['log_beta']) < TOL
        om, oc = O.cigp_ard_predict(x[b], y[b], xs[b], ls[b], sv[b:b + 1], lb[b:b + 1])
        assert rel_err(out['mean'][b].cpu(), om) < TOL and rel_err(out['var'][b].cpu(), oc.diag()) < TOL


def test_batched_ragged_sizes_through_the_fused_blocks_vs_oracle():
    """Batched problems (>= 8, so the fused 256-block kernel and the lower-super-tile kernel-matrix kernel run) whose
    size is NOT a multiple of the block sizes: n = 500 (padded to 512: identity tail inside the last 128-block and the
    last super-tile), n = 260 (padded to 384: base-kernel path, edge super-tiles), d = 3 and 16, two outputs."""
    from fidelityfusion_b200.batched import batched_cigp_eval
    for Bn, n, d, ns in ((9, 500, 3, 7), (8, 260, 16, 5)):
        gen = torch.Generator().manual_seed(900 + n)
        x = torch.rand(Bn, n, d, generator=gen)
        y = torch.sin(2 * x.sum(-1, keepdim=True)) + 0.1 * torch.randn(Bn, n, 1, generator=gen)
        ls = torch.exp(torch.rand(Bn, d, generator=gen) - 0.5)
        lb = torch.rand(Bn, generator=gen) * 2
        sv = torch.ones(Bn)
        xs = torch.rand(Bn, ns, d, generator=gen)
        out = batched_cigp_eval(x.to(DEV), y.to(DEV), ls.to(DEV), sv.to(DEV), lb.to(DEV), xs.to(DEV))
        for b in (0, Bn - 1):
            loss, gr = O.cigp_ard_nll_and_grads(x[b], y[b], ls[b], sv[b:b + 1], lb[b:b + 1])
            assert abs(out['nll'][b].item() - loss) <= TOL * abs(loss)
            assert rel_err(out['g_length_scales'][b].cpu(), gr['length_scales']) < TOL
            assert rel_err(out['g_signal_variance'][b].cpu().reshape(1), gr['signal_variance']) < TOL
            assert rel_err(out['g_log_beta'][b].cpu().reshape(1), gr['log_beta']) < TOL
            om, oc = O.cigp_ard_predict(x[b], y[b], xs[b], ls[b], sv[b:b + 1], lb[b:b + 1])
            assert rel_err(out['mean'][b].cpu(), om) < TOL and rel_err(out['var'][b].cpu(), oc.diag()) < TOL


def test_batched_tail_split_rows_equal_the_unsplit_evaluation():
    """330 problems = 2.23 waves of the 148 SMs: the default schedule runs the chunk as two parts on two streams (tail
    split, dense_gp.cu FFGP_SPLIT=2: 148 + 182 problems).  Every result row must equal, bit for bit, the row the same
    problem gets in a small batch that is not split (batches under 2 waves never are) - for problems of the first part,
    of the second part, and across the boundary - and the oracle at 1e-9 on a few of them."""
    gen = torch.Generator().manual_seed(41)
    B, n, d, ns = 330, 200, 3, 7
    x = torch.rand(B, n, d, generator=gen)
    y = torch.sin(3 * x.sum(2, keepdim=True)) + 0.05 * torch.randn(B, n, 1, generator=gen)
    ls = torch.rand(B, d, generator=gen) + 0.5
    sv = torch.ones(B)
    lb = torch.rand(B, generator=gen) * 3
    xs = torch.rand(B, ns, d, generator=gen)
    from fidelityfusion_b200.batched import batched_cigp_eval
    dev = [t.to(DEV) for t in (x, y, ls, sv, lb, xs)]
    full = batched_cigp_eval(*dev)
    assert torch.isfinite(full['_packed']).all()
    for lo, hi in ((0, 120), (100, 200), (210, 330)):
        part = batched_cigp_eval(*[t[lo:hi] for t in dev])
        assert torch.equal(part['_packed'], full['_packed'][lo:hi]), (lo, hi)
    for b in (0, 147, 148, 329):
        loss, gr = O.cigp_ard_nll_and_grads(x[b], y[b], ls[b], sv[b:b + 1], lb[b:b + 1])
        m, c = O.cigp_ard_predict(x[b], y[b], xs[b], ls[b], sv[b:b + 1], lb[b:b + 1])
        assert abs(float(full['nll'][b]) - loss) <= 1e-9 * abs(loss)
        assert rel_err(full['g_length_scales'][b], gr['length_scales']) < 1e-9
        assert rel_err(full['mean'][b], m) < 1e-9 and rel_err(full['var'][b], c.diag()) < 1e-9


def test_not_positive_definite_raises_linalg_error():
    from fidelityfusion_b200 import ops
    y = torch.randn(40, 1, device=DEV)
    S = -torch.eye(40, device=DEV)
    with pytest.raises(torch.linalg.LinAlgError):
        ops.dense_nll(None, y, None, None, sigma_add=S)
    S = torch.eye(300, device=DEV)
    S[200, 200] = -1.0
    with pytest.raises(torch.linalg.LinAlgError, match='order 201'):
        ops.dense_nll(None, torch.randn(300, 1, device=DEV), None, None, sigma_add=S)


def test_fp32_inputs_within_1e4():
    torch.set_default_dtype(torch.float32)
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    gen = torch.Generator().manual_seed(8)
    x, y = torch.rand(200, 4, generator=gen), torch.randn(200, 2, generator=gen)
    m = cigp(ARDKernel(4), 1.0).to(DEV)
    ll = m.negative_log_likelihood(x.to(DEV), y.to(DEV))
    assert ll.dtype == torch.float32
    (-ll).backward()
    loss, gr = O.cigp_ard_nll_and_grads(x.double(), y.double(), torch.ones(4, dtype=torch.float64),
                                        torch.ones(1, dtype=torch.float64), torch.ones(1, dtype=torch.float64))
    assert abs(-ll.item() - loss) <= 1e-4 * abs(loss)
    assert rel_err(m.kernel.length_scales.grad.double().cpu(), gr['length_scales']) < 1e-4
    assert rel_err(m.log_beta.grad.double().cpu(), gr['log_beta']) < 1e-4


def test_factor_and_inverse_properties_at_scale():
    """size-independent properties at N=2048 (multi-level recursion): L L^T = A, M L = I, log-det."""
    from fidelityfusion_b200 import ops
    n = 2048
    gen = torch.Generator(device=DEV).manual_seed(1)
    Bm = torch.randn(n, n, device=DEV, generator=gen)
    A = Bm @ Bm.T / n + torch.eye(n, device=DEV)
    L, M, logdet = ops.potrf_trtri(A)
    assert float((L @ L.T - A).abs().max() / A.abs().max()) < 1e-13
    assert float((M @ L - torch.eye(n, device=DEV)).abs().max()) < 1e-11
    assert float(torch.triu(L, 1).abs().max()) == 0.0 and float(torch.triu(M, 1).abs().max()) == 0.0
    assert abs(float(logdet) - float(torch.linalg.slogdet(A)[1])) < 1e-9 * n


def test_batched_factor_and_inverse_fused_256_blocks():
    """Batched problems (>= 8) factor their 256-blocks with the fused factor256_kernel by default: L, L^-1 and log|A| of
    32 x (n = 512) and 9 x (n = 300, ragged: padded to 384, base-kernel path) against torch.linalg, plus the LAPACK-style
    status of a batch with one non-PD problem."""
    from fidelityfusion_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(7)
    for batch, n in ((32, 512), (9, 300), (8, 256), (10, 500)):      # 500 -> 512: fused 256-blocks with an identity tail
        X = torch.randn(batch, n, n + 16, device=DEV, generator=gen)
        A = X @ X.transpose(1, 2) / n + 0.25 * torch.eye(n, device=DEV)
        L, M, logdet = ops.potrf_trtri(A)
        Lr = torch.linalg.cholesky(A)
        assert float((L - Lr).abs().max() / Lr.abs().max()) < 1e-12
        assert float((M @ Lr - torch.eye(n, device=DEV)).abs().max()) < 1e-10
        assert float((logdet - 2 * Lr.diagonal(dim1=1, dim2=2).log().sum(1)).abs().max()) < 1e-9 * n      # log|A|
    A = torch.eye(512, device=DEV).repeat(16, 1, 1)
    A[5, 300, 300] = -2.0
    with pytest.raises(torch.linalg.LinAlgError):
        ops.potrf_trtri(A)


def test_c2_full_size_anchor_and_gradient_identity():
    """BASELINE config 2 at full size (N=8192, d=16): the NLL anchor recorded from the reference on CPU
     and gradient identities that need no CPU factorisation:
    dNLL/dlog_beta = -e^{-lb} tr(G) and  sum_i y_i alpha_i = 2*(quadratic part)."""
    n, d = 8192, 16
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(n, d, generator=gen)
    y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn(n, 1, generator=gen)
    m = _cigp(d, np.ones(d), 1.0, 1.0)
    yy = y.to(DEV).requires_grad_(True)
    ll = m.negative_log_likelihood(x.to(DEV), yy)
    assert abs(-ll.item() - 10327.5181420215) < 1e-9 * 10327.5
    (-ll).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    # finite-difference check of the log_beta and signal_variance gradients (central, h=1e-4 => ~1e-8 rel)
    for name, p in (('log_beta', m.log_beta), ('signal_variance', m.kernel.signal_variance)):
        g0 = p.grad.item()
        h = 1e-4
        with torch.no_grad():
            p.add_(h)
            up = -m.negative_log_likelihood(x.to(DEV), y.to(DEV)).item()
            p.sub_(2 * h)
            dn = -m.negative_log_likelihood(x.to(DEV), y.to(DEV)).item()
            p.add_(h)
        fd = (up - dn) / (2 * h)
        assert abs(fd - g0) <= 1e-6 * max(abs(g0), 1.0), (name, fd, g0)


def test_c2_full_size_every_gradient_vs_oracle():
    """BASELINE config 2 at full size (N=8192, d=16): NLL, all 16 length-scale gradients, signal-variance and log_beta
    gradients and dNLL/dY against ONE evaluation of the CPU oracle (the reference's torch route: cdist kernel ->
    linalg.cholesky -> triangular solve -> autograd), at a perturbed parameter point (length_scales spread over
    [0.8, 1.7] with one negative entry for the abs() branch, signal_variance 1.3, log_beta 0.5); 1e-9 relative."""
    n, d = 8192, 16
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(n, d, generator=gen)
    y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn(n, 1, generator=gen)
    ls = 0.8 + 0.06 * torch.arange(d, dtype=torch.float64)
    ls[3] = -ls[3]
    sv, lb = 1.3, 0.5
    m = _cigp(d, ls.numpy(), sv, lb)
    yy = y.to(DEV).requires_grad_(True)
    ll = m.negative_log_likelihood(x.to(DEV), yy)
    (-ll).backward()
    loss, gr = O.cigp_ard_nll_and_grads(x, y, ls, T([sv]), T([lb]), want_y_grad=True)
    assert abs(-ll.item() - loss) <= TOL * abs(loss)
    assert rel_err(m.kernel.length_scales.grad.cpu(), gr['length_scales']) < TOL
    assert rel_err(m.kernel.signal_variance.grad.cpu(), gr['signal_variance']) < TOL
    assert rel_err(m.log_beta.grad.cpu(), gr['log_beta']) < TOL
    assert rel_err(yy.grad.cpu(), gr['y']) < TOL


def test_kernel_matrix_input_gradients_match_reference_golden():
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    g = load_golden('predict_dx')
    k = ARDKernel(4)
    with torch.no_grad():
        k.length_scales.copy_(T(g['k_ls']))
        k.signal_variance.copy_(T(g['k_sv']))
    k = k.to(DEV)
    x1, x2 = G(g['kx1']).requires_grad_(True), G(g['kx2']).requires_grad_(True)
    (k(x1, x2) * G(g['kW'])).sum().backward()
    assert rel_err(x1.grad.cpu(), g['g_kx1']) < TOL
    assert rel_err(x2.grad.cpu(), g['g_kx2']) < TOL
    # same tensor on both sides (K(x*, x*)): the two contributions add up
    xx = G(g['kx1']).requires_grad_(True)
    Wm = torch.randn(37, 37, generator=torch.Generator().manual_seed(0)).to(DEV)
    (k(xx, xx) * Wm).sum().backward()
    xo = T(g['kx1']).clone().requires_grad_(True)
    (O.ard_kernel(xo, xo, T(g['k_ls']), T(g['k_sv'])) * Wm.cpu()).sum().backward()
    assert rel_err(xx.grad.cpu(), xo.grad) < TOL


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_posterior_gradient_wrt_test_points_matches_reference_golden(tag):
    """cigp.forward differentiated w.r.t. x_test (acquisition optimisers, DMF_acq.py:247-254): full covariance
    weights and diagonal-only weights, then a second backward through the same graph (workspace replay)."""
    g = load_golden('predict_dx')
    d = g[f'x_{tag}'].shape[1]
    m = _cigp(d, g[f'ls_{tag}'], 1.3, 2.0)
    x, y = G(g[f'x_{tag}']), G(g[f'y_{tag}'])
    xs = G(g[f'xs_{tag}']).requires_grad_(True)
    mean, cov = m(x, y, xs)
    assert rel_err(mean.cpu(), g[f'mean_{tag}']) < TOL and rel_err(cov.cpu(), g[f'cov_{tag}']) < TOL
    s = (mean * G(g[f'wm_{tag}'])).sum() + (cov * G(g[f'wc_{tag}'])).sum()
    s.backward(retain_graph=True)
    assert rel_err(xs.grad.cpu(), g[f'gxs_full_{tag}']) < TOL
    xs.grad = None
    s.backward()                                   # the workspace was consumed: the forward is replayed
    assert rel_err(xs.grad.cpu(), g[f'gxs_full_{tag}']) < TOL
    xs.grad = None
    mean, cov = m(x, y, xs)                        # factor cache hit
    ((mean * G(g[f'wm_{tag}'])).sum() + (cov.diagonal() * G(g[f'wd_{tag}'])).sum()).backward()
    assert rel_err(xs.grad.cpu(), g[f'gxs_diag_{tag}']) < TOL
    # mean only
    xs.grad = None
    mean, _ = m(x, y, xs)
    (mean * G(g[f'wm_{tag}'])).sum().backward()
    _, _, gx = O.cigp_ard_predict_dx(T(g[f'x_{tag}']), T(g[f'y_{tag}']), T(g[f'xs_{tag}']), T(g[f'ls_{tag}']), T([1.3]),
                                     T([2.0]), T(g[f'wm_{tag}']), torch.zeros(xs.shape[0], xs.shape[0]))
    assert rel_err(xs.grad.cpu(), gx) < TOL


def test_posterior_gradient_diag_variance_C_ABI_and_finite_difference():
    """The diagonal-variance branch of ffgp_dense_predict_bwd_f64 through ops.dense_predict(full_cov=False), checked
    against the oracle and against a central difference of the CUDA forward itself at N=2048."""
    from fidelityfusion_b200 import ops
    gen = torch.Generator().manual_seed(9)
    n, d, ns = 2048, 6, 33
    x = torch.rand(n, d, generator=gen)
    y = torch.sin(3 * x.sum(1, keepdim=True)) + 0.05 * torch.randn(n, 1, generator=gen)
    xs0 = torch.rand(ns, d, generator=gen)
    il = torch.full((d,), 1 / 0.6)
    amp, noise = torch.tensor([1.2]), torch.tensor([0.05])
    wm, wv = torch.randn(ns, 1, generator=gen), torch.randn(ns, generator=gen)
    cache = ops.FactorCache()

    def f(xs):
        return ops.dense_predict(x.to(DEV), y.to(DEV), xs, il.to(DEV), amp.to(DEV), diag_add=(noise + 1e-6).to(DEV),
                                 cov_offset=noise.to(DEV), full_cov=False, cache=cache, cache_token='t')

    xs = xs0.to(DEV).requires_grad_(True)
    mean, var = f(xs)
    ((mean * wm.to(DEV)).sum() + (var * wv.to(DEV)).sum()).backward()
    ls = 0.6 * torch.ones(d) - 1e-9
    _, _, gx = O.cigp_ard_predict_dx(x, y, xs0, ls, amp, -torch.log(noise), wm, torch.diag(wv))
    assert rel_err(xs.grad.cpu(), gx) < 1e-8       # cond(Sigma) ~ 1e5 at this size; the oracle's own autograd noise
    with torch.no_grad():
        h = 1e-5
        e = torch.zeros(ns, d); e[3, 2] = h
        s = lambda o: float(((o[0] * wm.to(DEV)).sum() + (o[1] * wv.to(DEV)).sum()).item())
        fd = (s(f((xs0 + e).to(DEV))) - s(f((xs0 - e).to(DEV)))) / (2 * h)
    assert abs(fd - xs.grad[3, 2].item()) < 1e-5 * max(1.0, abs(fd))


@pytest.mark.parametrize('kind', ['UCB', 'EI', 'PI'])
def test_acquisition_scores_and_partials_match_reference_formulas(kind):
    """ffgp_acquisition_f64 vs the restated DMF_acq.py formulas (scipy cdf/pdf rounded to float32, autograd partials)."""
    from fidelityfusion_b200.MF_BayesianOptimization.Discrete.DMF_acq import acquisition
    gen = torch.Generator().manual_seed(3)
    mean = torch.randn(257, 1, generator=gen)
    var = torch.rand(257, 1, generator=gen) * 0.5 + 1e-4
    var[5] = 1e-20                                             # std clamp branch
    mo, vo = mean.clone().requires_grad_(True), var.clone().requires_grad_(True)
    so = O.acq_score(mo, vo, kind, f_best=0.3, x_dimension=5)
    so.sum().backward()
    mg, vg = mean.to(DEV).requires_grad_(True), var.to(DEV).requires_grad_(True)
    sg = acquisition(mg, vg, kind, f_best=0.3, beta=0.2 * 5, xi=0.01)
    sg.sum().backward()
    assert rel_err(sg.cpu(), so.detach().to(torch.float64)) < TOL
    assert rel_err(mg.grad.cpu(), mo.grad) < TOL
    ok = torch.ones(257, dtype=torch.bool); ok[5] = False      # d sqrt at 1e-20 under the clamp: 0 on both sides
    assert rel_err(vg.grad.cpu()[ok], vo.grad[ok]) < TOL
    assert float(vg.grad[5].abs().item()) == 0.0 or kind == 'UCB'


@pytest.mark.parametrize('kind', ['UCB', 'EI', 'PI'])
def test_acquisition_kernel_matches_the_reference_class_golden(kind):
    """ffgp_acquisition_f64 (scores + partials in one launch) against the UNMODIFIED DiscreteAcquisitionFunction
    (golden acq.npz from oracle/gen_golden_acq.py): UCB / PI at 1e-12, EI exactly reproduces the float32-rounded cdf/pdf
    of the scipy round trip (DMF_acq.py:104) up to the last-ulp difference between erfc and scipy's ndtr."""
    from fidelityfusion_b200.MF_BayesianOptimization.Discrete.DMF_acq import acquisition
    g = load_golden('acq')
    mu = G(g['mean']).requires_grad_(True)
    v = G(g['var']).requires_grad_(True)
    s = acquisition(mu, v, kind, f_best=float(g['f_best']), beta=0.2 * int(g['x_dimension']), xi=0.01)
    s.sum().backward()
    ref, dm, dv = g[kind + '_score'], g[kind + '_dmean'], g[kind + '_dvar']
    ok = np.isfinite(ref).reshape(-1) & np.isfinite(dv).reshape(-1) & (g['var'].reshape(-1) > 1e-17)
    # var <= 1e-18 sits on / below the clamp(std, 1e-9): PI is ~1e18 there and sqrt'(0) = inf in the reference's autograd
    tol = 1e-12 if kind != 'EI' else 2e-7          # one float32 ulp of cdf/pdf where erfc and ndtr round differently
    assert rel_err(s.detach().cpu().numpy().reshape(-1)[ok], ref.reshape(-1)[ok]) < tol
    assert rel_err(mu.grad.cpu().numpy().reshape(-1)[ok], dm.reshape(-1)[ok]) < tol
    assert rel_err(v.grad.cpu().numpy().reshape(-1)[ok], dv.reshape(-1)[ok]) < tol
    # clamped region: same score (std = 1e-9), zero variance-gradient (torch.clamp passes no gradient below the bound)
    lo = (g['var'].reshape(-1) < 1e-18) & np.isfinite(ref).reshape(-1)
    if kind != 'UCB' and lo.any():
        assert rel_err(s.detach().cpu().numpy().reshape(-1)[lo], ref.reshape(-1)[lo]) < 1e-9
        assert float(np.abs(v.grad.cpu().numpy().reshape(-1)[lo]).max()) == 0.0


def test_candidate_optimisation_step_on_device():
    """UCB/EI of the posterior differentiated w.r.t. the candidate through DiscreteAcquisitionFunction ->
    cigp.forward -> ffgp_dense_predict_bwd_f64, against the same composition on the CPU oracle."""
    from fidelityfusion_b200.MF_BayesianOptimization.Discrete.DMF_acq import DiscreteAcquisitionFunction, optimize_acq_candidates
    gen = torch.Generator().manual_seed(12)
    n, d = 200, 4
    x = torch.rand(n, d, generator=gen)
    y = torch.sin(4 * x.sum(1, keepdim=True)) + 0.05 * torch.randn(n, 1, generator=gen)
    ls = [0.5, 0.8, 0.6, 1.1]
    m = _cigp(d, ls, 1.0, 3.0)
    xd, yd = x.to(DEV), y.to(DEV)
    mean_f = lambda xx, s: m(xd, yd, xx)[0]
    var_f = lambda xx, s: m(xd, yd, xx)[1].diagonal().reshape(-1, 1)
    acq = DiscreteAcquisitionFunction(mean_f, var_f, 1, d, f_best=float(y.max()))
    xc = torch.rand(1, d, generator=gen)
    for kind in ('UCB', 'EI'):
        xg = xc.to(DEV).requires_grad_(True)
        s = getattr(acq, kind + '_MF')(xg, 0)
        s.sum().backward()
        xo = xc.clone().requires_grad_(True)
        K = O.ard_kernel(x, x, T(ls), T([1.0]))
        mo, co = O.cigp_predict(K, O.ard_kernel(x, xo, T(ls), T([1.0])), O.ard_kernel(xo, xo, T(ls), T([1.0])), T([3.0]), y)
        so = O.acq_score(mo, co.diagonal().reshape(-1, 1), kind, f_best=float(y.max()), x_dimension=d)
        so.sum().backward()
        assert rel_err(s.cpu(), so.detach().to(torch.float64)) < 1e-8
        assert rel_err(xg.grad.cpu(), xo.grad) < 1e-7       # EI multiplies float32-rounded cdf/pdf: 1e-7 headroom
    # the optimiser loop runs and improves the score
    x0 = [xc.clone()]
    s0 = float(acq.UCB_MF(x0[0].to(DEV), 0).item())
    xb = optimize_acq_candidates(lambda xx, s: acq.UCB_MF(xx, s), 1, d, n_iterations=15, learning_rate=0.01, x_init=x0)
    assert float(acq.UCB_MF(xb, 0).item()) > s0
