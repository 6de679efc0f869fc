"""Single-fidelity acquisition functions (reference Bayesian_optimization/acq.py:118-294) on the device,
against golden vectors of the UNMODIFIED reference module, and the acquisition epilogue of the batched sweep."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def G(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64).cuda()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_ucb_ei_pi_pf_match_the_reference_module_golden():
    from fidelityfusion_b200.Bayesian_optimization import acq as A
    g = load_golden('acq_sf')
    f_best, kappa, xi = float(g['f_best']), float(g['kappa']), float(g['xi'])
    varr = g['var'].reshape(-1)
    for kind in ('UCB', 'EI'):
        mu, v = G(g['mean']).requires_grad_(True), G(g['var']).requires_grad_(True)
        fn = A.UCB(lambda X: mu, lambda X: v, kappa=kappa) if kind == 'UCB' else A.EI(lambda X: mu, lambda X: v, xi=xi)
        s = fn.forward(None) if kind == 'UCB' else fn.forward(None, f_best)
        assert s.is_cuda and s.dtype == torch.float64 and tuple(s.shape) == g[kind + '_score'].shape
        s.sum().backward()
        ref, dm, dv = g[kind + '_score'], g[kind + '_dmean'], g[kind + '_dvar']
        # var <= 1e-18 sits on / below clamp(std, 1e-9) (EI) or at sqrt'(0) = inf (UCB): compared separately
        ok = np.isfinite(dv).reshape(-1) & (varr > 1e-17)
        tol = 1e-12 if kind == 'UCB' else 2e-7       # EI: one float32 ulp of cdf / pdf where erfc and scipy's ndtr round differently
        assert rel_err(s.detach().cpu().numpy().reshape(-1)[ok], ref.reshape(-1)[ok]) < tol
        assert rel_err(mu.grad.cpu().numpy().reshape(-1)[ok], dm.reshape(-1)[ok]) < tol
        assert rel_err(v.grad.cpu().numpy().reshape(-1)[ok], dv.reshape(-1)[ok]) < tol
        lo = varr < 1e-18
        assert rel_err(s.detach().cpu().numpy().reshape(-1)[lo], ref.reshape(-1)[lo]) < 1e-9
        if kind == 'EI':
            assert float(np.abs(v.grad.cpu().numpy().reshape(-1)[lo]).max()) == 0.0      # torch.clamp passes no gradient below the bound
    pi = A.PI(lambda X: G(g['mean']), lambda X: G(g['var']), sita=xi).forward(None, f_best)
    assert pi.dtype == torch.float32 and not pi.requires_grad
    assert float(np.abs(pi.cpu().numpy() - g['PI_score']).max()) < 2e-7                  # float32 cdf, one ulp
    pf = A.PF(lambda X: G(g['pf_mean']), lambda X: G(g['pf_var']), list(g['pf_thresholds'])).forward(torch.zeros(301, 2))
    assert rel_err(pf.cpu().numpy(), g['PF_score']) < 1e-12


def _toy_posterior():
    """cigp drop-in on a 1-d toy problem: mean / variance callables as the reference's acq_demo.py builds them."""
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    torch.manual_seed(3)
    x = torch.rand(40, 1, dtype=torch.float64).cuda() * 4 - 2
    y = torch.sin(2 * x) + 0.05 * torch.randn(40, 1, dtype=torch.float64).cuda()
    m = cigp(ARDKernel(1), 2.0).cuda().double()

    def post(X):
        mean, cov = m.forward(x, y, X.to(torch.float64))
        return mean, cov.diag().reshape(-1, 1)
    return (lambda X: post(X)[0]), (lambda X: post(X)[1]), float(y.max())


def test_candidate_search_and_optimisation_run_on_the_device():
    from fidelityfusion_b200.Bayesian_optimization import acq as A
    mean_f, var_f, f_best = _toy_posterior()
    bounds = torch.tensor([[-2.0, 2.0]], device='cuda')
    torch.manual_seed(11)
    for fn in (A.UCB(mean_f, var_f, kappa=2.0), A.EI(mean_f, var_f), A.PI(mean_f, var_f), A.KG(mean_f, var_f)):
        nxt = A.find_next_batch(fn, bounds, batch_size=2, n_samples=200, f_best=f_best)
        assert nxt.is_cuda and tuple(nxt.shape) == (2, 1) and bool(((nxt >= -2) & (nxt <= 2)).all())
    # the picked point really is the arg-max of the score over the drawn set
    ucb = A.UCB(mean_f, var_f, kappa=2.0)
    torch.manual_seed(5)
    pick = A.find_next_batch(ucb, bounds, batch_size=1, n_samples=300)
    torch.manual_seed(5)
    X = torch.empty(300, 1, dtype=torch.float32, device='cuda').uniform_(-2.0, 2.0)
    assert torch.equal(pick[0], X[torch.argmax(ucb.forward(X))])
    # gradient ascent on the candidates through the kernel's partials and the fused posterior gradient: the summed score
    # of the returned iterate is not worse than that of the start points
    torch.manual_seed(7)
    best = A.optimize_acqf(ucb, raw_samples=16, bounds=bounds, num_restarts=10)
    torch.manual_seed(7)
    start = torch.rand((16, 1), dtype=torch.float32, device='cuda') * 4.0 - 2.0
    assert tuple(best.shape) == (16, 1) and float(ucb.forward(best).sum()) >= float(ucb.forward(start).sum()) - 1e-12


@pytest.mark.parametrize('kind', ['EI', 'UCB', 'PI', 'UCB_STD', 'PI_CDF'])
def test_batched_sweep_writes_the_scores_in_its_epilogue(kind):
    """out['score'] of batched_cigp_eval(acq=...) (ffgp_batched_pack_acq_f64: the score block of the packed result row)
    equals the stand-alone kernel applied to the sweep's own mean / variance (1e-14: the same device function inlined in
    two kernels, the compiler may contract its multiply-adds differently), and the other fields are unchanged, bit for
    bit, by the extra block."""
    from fidelityfusion_b200.batched import batched_cigp_eval, ACQ_KINDS
    from fidelityfusion_b200.MF_BayesianOptimization.Discrete.DMF_acq import acquisition
    g = torch.Generator().manual_seed(9)
    B, n, d, ns = 12, 96, 3, 17
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64).cuda()
    y = torch.sin(3 * x.sum(2, keepdim=True)) + 0.05 * torch.randn(B, n, 1, generator=g, dtype=torch.float64).cuda()
    ls = (torch.rand(B, d, generator=g, dtype=torch.float64) + 0.5).cuda()
    sv = torch.ones(B, dtype=torch.float64).cuda()
    lb = (torch.rand(B, generator=g, dtype=torch.float64) * 3).cuda()
    xs = torch.rand(B, ns, d, generator=g, dtype=torch.float64).cuda()
    a = dict(kind=kind, f_best=0.4, beta=1.7, xi=0.02)
    base = batched_cigp_eval(x, y, ls, sv, lb, xs)
    out = batched_cigp_eval(x, y, ls, sv, lb, xs, acq=a)
    for k in ('nll', 'g_length_scales', 'g_signal_variance', 'g_log_beta', 'mean', 'var'):
        assert torch.equal(out[k], base[k])
    ref = acquisition(out['mean'][..., 0], out['var'], ACQ_KINDS[kind], f_best=0.4, beta=1.7, xi=0.02)
    assert tuple(out['score'].shape) == (B, ns)
    assert rel_err(out['score'].cpu().numpy(), ref.cpu().numpy()) < 1e-14
    out2 = batched_cigp_eval(x, y, ls, sv, lb, xs, want_grad=False, check=False, acq=a)
    # prediction-only sweep (alpha through the two triangular passes instead of S y: last-bit different predictions)
    ref2 = acquisition(out2['mean'][..., 0], out2['var'], ACQ_KINDS[kind], f_best=0.4, beta=1.7, xi=0.02)
    assert rel_err(out2['score'].cpu().numpy(), ref2.cpu().numpy()) < 1e-14 and float(out2['info'].abs().max()) == 0.0
    assert rel_err(out2['score'].cpu().numpy(), out['score'].cpu().numpy()) < 1e-9
