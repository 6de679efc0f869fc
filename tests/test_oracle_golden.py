"""Pin the CPU oracle (oracle/ff_oracle.py) against golden vectors produced by the real
reference (oracle/gen_golden.py) and the known-answer tests of SURVEY.md appendix B."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ff_oracle as O

torch.set_default_dtype(torch.float64)
T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
TOL = 1e-12


def P(v):
    return T(v).clone().requires_grad_(True)


def test_kernels():
    g = load_golden('kernels')
    for tag in ('small', 'mm'):
        x1, x2, ls = T(g[f'{tag}_x1']), T(g[f'{tag}_x2']), T(g[f'{tag}_ls'])
        d = x1.shape[1]
        assert rel_err(O.ard_kernel(x1, x2, ls, T([-1.7])), g[f'{tag}_ard']) < TOL
        assert rel_err(O.ard_kernel(x1, x1, ls, T([-1.7])), g[f'{tag}_ard_sym']) < TOL
        assert rel_err(O.sqexp_kernel(x1, x2, T([0.3]), T([-0.2])), g[f'{tag}_sqexp']) < TOL
        lsl = torch.log(T([0.7 + 0.1 * i for i in range(d)]))
        assert rel_err(O.se_kernel(x1, x2, lsl, torch.log(T(1.3)), True), g[f'{tag}_se_exp']) < TOL
        assert rel_err(O.se_kernel(x1, x2, T(0.8), T(2.0), False), g[f'{tag}_se_lin']) < TOL
        assert not bool(g[f'{tag}_se_cfg_is_exp'])      # dict config => linear format (kernel_utils.py:12)


def test_kat1_cigp_ard_values_from_survey():
    g = load_golden('kat1_cigp_ard')
    x, y, xs = T(g['x']), T(g['y']), T(g['xs'])
    ls, sv, lb = P([1., 1.]), P([1.]), P([1.])
    ll = O.cigp_log_likelihood(O.ard_kernel(x, x, ls, sv), lb, y)
    assert abs(ll.item() - (-14.408784462236161)) < 1e-12
    assert rel_err(ll.detach().reshape(()), g['ll']) < TOL
    (-ll).backward()
    assert rel_err(ls.grad, g['g_length_scales']) < 1e-11
    assert rel_err(ls.grad, [-0.8810011435841414, -0.5206954919592008]) < 1e-11
    assert rel_err(sv.grad, g['g_signal_variance']) < 1e-11
    assert rel_err(lb.grad, g['g_log_beta']) < 1e-11
    mean, cov = O.cigp_ard_predict(x, y, xs, ls.detach(), sv.detach(), lb.detach())
    assert rel_err(mean, g['mean']) < TOL and rel_err(cov, g['cov']) < TOL
    assert abs(cov[1, 2].item() - 0.34143245709563463) < 1e-12


def test_kat2_CIGP2023():
    g = load_golden('kat2_CIGP2023')
    x, y, xs = T(g['x']), T(g['y']), T(g['xs'])
    assert not bool(g['exp_format'])
    l, s, nv = P(1.), P(1.), P(0.)                              # noise 'exp' format, init log(1.)
    K = O.se_kernel(x, x, l, s, False)
    loss = O.CIGP_loss(K, nv.exp(), y)
    assert abs(loss.item() - 19.35232673878598) < 1e-11
    loss.backward()
    assert rel_err(nv.grad, g['g_noise']) < 1e-11
    assert rel_err(l.grad, g['g_length_scale']) < 1e-11
    assert rel_err(s.grad, g['g_scale']) < 1e-11
    with torch.no_grad():
        u, v = O.CIGP_predict(K, O.se_kernel(x, xs, l, s, False), O.se_kernel(xs, xs, l, s, False).diag(), nv.exp(), y)
    assert rel_err(u, g['u']) < TOL and rel_err(v, g['var']) < TOL


def _hogp_from_golden(g, n_modes, params=None):
    x, Y, xs = T(g['x']), T(g['Y']), T(g['xs'])
    shape = Y.shape[1:]
    if params is None:
        params = [(1., 1.)] * (n_modes + 1)
    ps = [(P(a), P(b)) for a, b in params]
    grids = [torch.arange(s, dtype=torch.float64).reshape(-1, 1) for s in shape]
    Ks = [O.se_kernel(x, x, ps[0][0], ps[0][1], False)]
    for k, gr in enumerate(grids):
        Ks.append(O.se_kernel(gr, gr, ps[k + 1][0], ps[k + 1][1], False))
    return x, Y, xs, ps, Ks


def test_kat3_HOGP2023():
    g = load_golden('kat3_HOGP2023')
    x, Y, xs, ps, Ks = _hogp_from_golden(g, 2)
    nv = P(1.)
    loss, A, gg = O.hogp_loss(Ks, nv.pow(-1), Y)
    assert abs(loss.item() - 1.193667663383912) < 1e-12
    loss.backward()
    assert rel_err(nv.grad, g['g_noise_box_value']) < 1e-9
    for k in range(3):
        assert rel_err(ps[k][0].grad, g[f'g_kernel_list_{k}_length_scale']) < 1e-8
        assert rel_err(ps[k][1].grad, g[f'g_kernel_list_{k}_scale']) < 1e-9
    with torch.no_grad():
        Kst = O.se_kernel(xs, x, ps[0][0], ps[0][1], False)
        kss = O.se_kernel(xs, xs, ps[0][0], ps[0][1], False).diag()
        u, v = O.hogp_predict([k.detach() for k in Ks], Kst, kss, A.detach(), gg.detach(), 'hogp2023')
    assert rel_err(u, g['u']) < 1e-10 and rel_err(v, g['var']) < 1e-10


def test_hogp2023_params_and_ygrad():
    g = load_golden('hogp2023_params')
    params = [(0.2 * (i + 1) - 0.3, 0.1 * (i + 1)) for i in range(4)]
    x, Y, xs, ps, Ks = _hogp_from_golden(g, 3, params)
    Y = Y.clone().requires_grad_(True)
    nv = P(3.0)
    loss, A, gg = O.hogp_loss(Ks, nv.pow(-1), Y)
    assert rel_err(loss.detach(), g['loss']) < 1e-12
    loss.backward()
    assert rel_err(Y.grad, g['gY']) < 1e-10
    assert rel_err(A.detach(), g['A']) < 1e-10 and rel_err(gg.detach(), g['g']) < 1e-9
    assert rel_err(nv.grad, g['g_noise_box_value']) < 1e-9


def test_c2_dense_nll_grad_and_predict():
    g = load_golden('c2_n512')
    x, y = T(g['x']), T(g['y'])
    for tag in ('init', 'ls_half', 'ls_double', 'lb_m2', 'lb_3', 'sv_neg'):
        ls0, sv0, lb0 = g[f'{tag}_params']
        loss, gr = O.cigp_ard_nll_and_grads(x, y, T(np.full(16, ls0)), T([sv0]), T([lb0]), want_y_grad=True)
        assert abs(-loss - float(g[f'{tag}_ll'])) <= 1e-12 * abs(float(g[f'{tag}_ll']))
        for k in ('length_scales', 'signal_variance', 'log_beta', 'y'):
            assert rel_err(gr[k], g[f'{tag}_g_{k}']) < 1e-10, (tag, k)
        # the closed-form route through Sigma^-1 (what the CUDA path implements) agrees with autograd
        nll, ga = O.dense_nll_grads_analytic_numpy(g['x'], g['y'], np.full(16, ls0), sv0, lb0)
        assert abs(nll - loss) <= 1e-11 * abs(loss)
        assert rel_err(ga['length_scales'], g[f'{tag}_g_length_scales']) < 1e-9
        assert rel_err([ga['signal_variance']], g[f'{tag}_g_signal_variance']) < 1e-9
        assert rel_err([ga['log_beta']], g[f'{tag}_g_log_beta']) < 1e-9
        assert rel_err(ga['y'], g[f'{tag}_g_y']) < 1e-9
    # SURVEY appendix B anchor for the N=512 recipe
    assert abs(float(g['init_ll']) + 647.12268694965) < 1e-8
    mean, cov = O.cigp_ard_predict(x, y, T(g['xs']), T(np.full(16, 2.0)), T([1.0]), T([1.0]))
    assert rel_err(mean, g['pred_mean']) < TOL and rel_err(cov, g['pred_cov']) < TOL


def test_c3_yvar_tensor_linear():
    g = load_golden('c3_small')
    x, ylo, yhi, yv = T(g['x']), T(g['y_low']), T(g['y_high']), T(g['y_var'])
    w = O.tensor_linear_init(16, 64).double()
    assert rel_err(w, g['tl_init']) < TOL
    w = w.clone().requires_grad_(True)
    ls, sv, lb = P(np.full(5, 0.7)), P([1.2]), P([0.5])
    res = yhi - O.tensor_linear_forward(ylo, [w])
    assert rel_err(res.detach(), g['res']) < TOL
    ll = O.cigp_log_likelihood(O.ard_kernel(x, x, ls, sv), lb, res, yv)
    assert rel_err(ll.detach().reshape(()), g['ll']) < TOL
    (-ll).backward()
    assert rel_err(ls.grad, g['g_length_scales']) < 1e-10
    assert rel_err(sv.grad, g['g_signal_variance']) < 1e-10
    assert rel_err(lb.grad, g['g_log_beta']) < 1e-10
    assert rel_err(w.grad, g['g_tl']) < 1e-10
    mean, cov = O.cigp_ard_predict(x, res.detach(), T(g['xs']), ls.detach(), sv.detach(), lb.detach())
    assert rel_err(mean, g['mean']) < TOL and rel_err(cov, g['cov']) < TOL
    g2 = load_golden('tensor_linear_2mode')
    assert rel_err(O.tensor_linear_forward(T(g2['t']), [T(g2['w0']), T(g2['w1'])]), g2['out']) < TOL


def test_pack_and_gp_basic():
    g = load_golden('pack')
    S, Ks, Kss = T(g['Sigma']), T(g['K_s']), T(g['K_ss'])
    for D in (1, 3):
        y = T(g[f'y{D}'])
        for meth in ('cholesky1', 'cholesky2', 'cholesky3', 'direct'):
            key = f'gll_{meth}_D{D}'
            if key in g:
                assert rel_err(O.gaussian_log_likelihood(y, S, meth), g[key]) < 1e-11
        for meth in ('cholesky1', 'cholesky3', 'direct'):
            mu, cov = O.conditional_gaussian(y, S, Ks, Kss, meth)
            assert rel_err(mu, g[f'cg_{meth}_D{D}_mu']) < 1e-10
            assert rel_err(cov, g[f'cg_{meth}_D{D}_cov']) < 1e-10
    x, xs, y = T(g['x']), T(g['xs']), T(g['y3'])
    ls, sv, lb = P(np.full(3, 1.1)), P([0.9]), P([0.7])
    ll = O.pack_negative_log_likelihood(lambda a, b: O.ard_kernel(a, b, ls, sv), lb, x, y)
    assert rel_err(ll.detach().reshape(()), g['pack_nll']) < TOL
    (-ll).backward()
    assert rel_err(lb.grad, g['pack_nll_g_lb']) < 1e-10 and rel_err(ls.grad, g['pack_nll_g_ls']) < 1e-10
    for D in (1, 3):
        y = T(g[f'y{D}'])
        ls, sv, nv = P(np.full(3, 1.1)), P([0.9]), P([0.4])
        cov = O.gp_basic_cov(O.ard_kernel(x, x, ls, sv), nv)
        ll = O.gaussian_log_likelihood(y, cov, 'cholesky3')
        assert rel_err(ll.detach(), g[f'gpb_ll_D{D}']) < 1e-11
        (-ll.sum()).backward()
        assert rel_err(nv.grad, g[f'gpb_g_noise_D{D}']) < 1e-10
        assert rel_err(ls.grad, g[f'gpb_g_ls_D{D}']) < 1e-10
        with torch.no_grad():
            mu, c = O.conditional_gaussian(y, cov, O.ard_kernel(x, xs, ls, sv), O.ard_kernel(xs, xs, ls, sv))
        assert rel_err(mu.squeeze(), g[f'gpb_mu_D{D}']) < 1e-10 and rel_err(c, g[f'gpb_cov_D{D}']) < 1e-10


def test_c1_AR2023_first_step():
    g = load_golden('c1_AR2023')
    x, y0, y1 = T(g['x']), T(g['y0']), T(g['y1'])
    pr = [(P(1.), P(1.), P(0.)) for _ in range(2)]
    rho = P(1.)
    loss = 0.
    for f, (l, s, nv) in enumerate(pr):
        tgt = y0 if f == 0 else y1 - y0 * rho          # Residual.forward, Residual.py:20-22
        loss = loss + O.CIGP_loss(O.se_kernel(x, x, l, s, False), nv.exp(), tgt)
    assert abs(loss.item() - g['losses'][0]) <= 1e-12 * abs(g['losses'][0])
    loss.backward()
    assert rel_err(rho.grad, g['g0_residual_list_0_rho']) < 1e-10
    for f, (l, s, nv) in enumerate(pr):
        assert rel_err(l.grad, g[f'g0_cigp_list_{f}_kernel_length_scale']) < 1e-10
        assert rel_err(s.grad, g[f'g0_cigp_list_{f}_kernel_scale']) < 1e-10
        assert rel_err(nv.grad, g[f'g0_cigp_list_{f}_noise_box_value']) < 1e-10


def test_c4_GAR2023_loss():
    g = load_golden('c4_GAR2023')
    x, Ylo, Yhi = T(g['x']), T(g['Ylo']), T(g['Yhi'])
    shape = Ylo.shape[1:]
    grids = [torch.arange(s, dtype=torch.float64).reshape(-1, 1) for s in shape]
    Ws = [P(np.eye(s)) for s in shape]
    loss = 0.
    allp = []
    for f in range(2):
        ps = [(P(1.), P(1.)) for _ in range(4)]
        nv = P(1.)
        allp.append((ps, nv))
        Ks = [O.se_kernel(x, x, ps[0][0], ps[0][1], False)] + \
             [O.se_kernel(gr, gr, ps[k + 1][0], ps[k + 1][1], False) for k, gr in enumerate(grids)]
        tgt = Ylo if f == 0 else O.matrix_mapping_forward(Ylo, Yhi, Ws, T(1.))
        loss = loss + O.hogp_loss(Ks, nv.pow(-1), tgt)[0]
    assert rel_err(loss.detach(), g['loss']) < 1e-12
    loss.backward()
    for k in range(3):
        assert rel_err(Ws[k].grad, g[f'g_matrix_list_0_vectors_{k}']) < 1e-8
    for f in range(2):
        assert rel_err(allp[f][1].grad, g[f'g_hogp_list_{f}_noise_box_value']) < 1e-8


def test_hogp_simple_copies():
    for tag in ('ffm', 'gp'):
        g = load_golden(f'hogp_simple_{tag}')
        x, Y, xs = T(g['x']), T(g['Y']), T(g['xs'])
        l, s, nv = P([0.2]), P([0.1]), P([2.0])
        grids = [torch.arange(n, dtype=torch.float64).reshape(-1, 1) for n in Y.shape[1:]]
        Ks = [O.sqexp_kernel(x, x, l, s)] + [O.sqexp_kernel(gr, gr, l, s) for gr in grids]
        loss, A, gg = O.hogp_loss(Ks, nv.pow(-1), Y)
        assert rel_err(loss.detach(), g['loss']) < 1e-12
        loss.backward()
        assert rel_err(nv.grad, g['g_noise']) < 1e-8
        assert rel_err(l.grad, g['g_length_scale']) < 1e-7
        with torch.no_grad():
            u, v = O.hogp_predict([k.detach() for k in Ks], O.sqexp_kernel(xs, x, l, s),
                                  O.sqexp_kernel(xs, xs, l, s).diag(), A.detach(), gg.detach(), tag)
        assert rel_err(u, g['u']) < 1e-9
        assert rel_err(v, g['var']) < (1e-5 if tag == 'ffm' else 1e-9)   # 'ffm' inverts K0 explicitly (cond ~1e10)


def test_couplings():
    g = load_golden('couplings')
    lo, hi = T(g['lo']), T(g['hi'])
    w0 = O.smooth_mapping_matrix(4, 8)
    assert rel_err(w0, g['w0']) < 1e-12
    w0 = w0.clone().requires_grad_(True)
    w1 = T(g['w1'])
    rho = P(g['rho'])
    res = O.matrix_mapping_forward(lo, hi, [w0, w1], rho)
    assert rel_err(res.detach(), g['res']) < TOL
    res.pow(2).sum().backward()
    assert rel_err(w0.grad, g['g_w0']) < 1e-11 and rel_err(rho.grad, g['g_rho']) < 1e-11
    assert rel_err(O.matrix_mapping_backward(lo, res.detach(), [w0.detach(), w1], rho.detach()), g['back']) < TOL
    assert rel_err(O.tensor_linear_init(4, 8).double(), g['eye_init']) < 1e-7   # reference init is fp32-rounded


def test_c5_batch():
    g = load_golden('c5_batch')
    for b in range(6):
        x, y, xs = T(g[f'x{b}']), T(g[f'y{b}']), T(g[f'xs{b}'])
        loss, gr = O.cigp_ard_nll_and_grads(x, y, T(g[f'ls{b}']), T([1.0]), T([float(g[f'lb{b}'])]))
        assert abs(-loss - float(g[f'll{b}'])) <= 1e-12 * abs(float(g[f'll{b}']))
        assert rel_err(gr['length_scales'], g[f'g_ls{b}']) < 1e-10
        assert rel_err(gr['signal_variance'], g[f'g_sv{b}']) < 1e-10
        assert rel_err(gr['log_beta'], g[f'g_lb{b}']) < 1e-10
        mean, cov = O.cigp_ard_predict(x, y, xs, T(g[f'ls{b}']), T([1.0]), T([float(g[f'lb{b}'])]))
        assert rel_err(mean, g[f'mean{b}']) < TOL and rel_err(cov.diag(), g[f'vdiag{b}']) < TOL


def test_posterior_gradient_wrt_test_points():
    """d posterior / d x* and d K / d inputs: oracle (autograd restatement) vs the reference's own autograd."""
    g = load_golden('predict_dx')
    for tag in ('a', 'b'):
        x, y, xs, ls = T(g[f'x_{tag}']), T(g[f'y_{tag}']), T(g[f'xs_{tag}']), T(g[f'ls_{tag}'])
        mean, cov, gx = O.cigp_ard_predict_dx(x, y, xs, ls, T([1.3]), T([2.0]), T(g[f'wm_{tag}']), T(g[f'wc_{tag}']))
        assert rel_err(mean, g[f'mean_{tag}']) < TOL and rel_err(cov, g[f'cov_{tag}']) < 1e-11
        assert rel_err(gx, g[f'gxs_full_{tag}']) < 1e-10
        _, _, gx = O.cigp_ard_predict_dx(x, y, xs, ls, T([1.3]), T([2.0]), T(g[f'wm_{tag}']), torch.diag(T(g[f'wd_{tag}'])))
        assert rel_err(gx, g[f'gxs_diag_{tag}']) < 1e-10
    x1, x2 = P(g['kx1']), P(g['kx2'])
    (O.ard_kernel(x1, x2, T(g['k_ls']), T(g['k_sv'])) * T(g['kW'])).sum().backward()
    assert rel_err(x1.grad, g['g_kx1']) < 1e-11 and rel_err(x2.grad, g['g_kx2']) < 1e-11


@pytest.mark.parametrize('kind', ['UCB', 'EI', 'PI'])
def test_acquisition_scores_pinned_on_the_reference_class(kind):
    """O.acq_score against the unmodified DiscreteAcquisitionFunction (DMF_acq.py:49-128; golden from
    oracle/gen_golden_acq.py), scores and autograd partials, including the float32-rounded cdf/pdf of EI (:104) and the
    clamp(std, min=1e-9) region."""
    g = load_golden('acq')
    mu = T(g['mean']).requires_grad_(True)
    v = T(g['var']).requires_grad_(True)
    s = O.acq_score(mu, v, kind, f_best=float(g['f_best']), x_dimension=int(g['x_dimension']))
    assert s.shape == g[kind + '_score'].shape
    assert np.array_equal(s.detach().numpy(), g[kind + '_score'])           # same operations in the same order: bit-exact
    s.sum().backward()
    assert np.array_equal(mu.grad.numpy(), g[kind + '_dmean'])
    assert np.array_equal(np.nan_to_num(v.grad.numpy(), nan=-7.0), np.nan_to_num(g[kind + '_dvar'], nan=-7.0))


def test_single_fidelity_acquisition_scores_pinned_on_the_reference_module():
    """O.acq_sf_score against the unmodified Bayesian_optimization/acq.py classes (UCB :135-149, EI :166-181, PI :211-231,
    PF :279-294; golden from oracle/gen_golden_acq_sf.py): scores and autograd partials, the float32 cdf / pdf tensors of
    EI / PI and the clamp(std, min=1e-9) region.  Same operations in the same order: bit-exact."""
    g = load_golden('acq_sf')
    f_best, kappa, xi = float(g['f_best']), float(g['kappa']), float(g['xi'])
    for kind in ('UCB', 'EI'):
        mu = T(g['mean']).requires_grad_(True)
        v = T(g['var']).requires_grad_(True)
        s = O.acq_sf_score(kind, mu, v, f_best=f_best, kappa=kappa, xi=xi)
        assert np.array_equal(s.detach().numpy(), g[kind + '_score'])
        s.sum().backward()
        assert np.array_equal(mu.grad.numpy(), g[kind + '_dmean'])
        assert np.array_equal(np.nan_to_num(v.grad.numpy(), nan=-7.0), np.nan_to_num(g[kind + '_dvar'], nan=-7.0))
    pi = O.acq_sf_score('PI', T(g['mean']), T(g['var']), f_best=f_best, xi=xi)
    assert pi.dtype == torch.float32 and np.array_equal(pi.numpy(), g['PI_score'])
    pf = O.acq_sf_score('PF', T(g['pf_mean']), T(g['pf_var']), thresholds=list(g['pf_thresholds']))
    assert np.array_equal(pf, g['PF_score'])


@pytest.mark.parametrize('tag', ['full', 'bcast'])
def test_hogp2023_tensor_valued_y_var(tag):
    """O.hogp_loss with a tensor y_var (`A = A + y_var`, hogp.py:176) against the unmodified HOGP.compute_loss (golden
    from oracle/gen_golden_hogp_yvar.py): one value per element of A, and a per-sample variance broadcast over the grid."""
    g = load_golden('hogp2023_yvar')
    params = [(-1.2 + 0.25 * i, 0.1 * (i + 1)) for i in range(4)]
    x, Y, xs, ps, Ks = _hogp_from_golden(g, 3, params)
    Y = Y.clone().requires_grad_(True)
    yv = P(g[f'{tag}_y_var'])
    nv = P(2.5)
    loss, A, gg = O.hogp_loss(Ks, nv.pow(-1), Y, y_var=yv)
    assert rel_err(loss.detach(), g[f'{tag}_loss']) < 1e-12
    loss.backward()
    assert rel_err(Y.grad, g[f'{tag}_gY']) < 1e-10 and rel_err(yv.grad, g[f'{tag}_g_y_var']) < 1e-10
    assert rel_err(A.detach(), g[f'{tag}_A']) < 1e-10 and rel_err(gg.detach(), g[f'{tag}_g']) < 1e-9
    assert rel_err(nv.grad, g[f'{tag}_g_noise']) < 1e-9
    for k in range(4):
        assert rel_err(ps[k][0].grad, g[f'{tag}_g_ls{k}']) < 1e-7 and rel_err(ps[k][1].grad, g[f'{tag}_g_sc{k}']) < 1e-7
