"""BASELINE configs 1, 3, 4 at FULL size against the CPU trajectories of the unmodified reference
(tests/golden/l4_*.npz, written by oracle/gen_golden_l4.py), without needing the reference tree at run time.

tests/test_binding.py runs the reference's own L4 files on the drop-ins when a checkout is present.  On a box without
one (the reference cannot travel) these tests replay the same compositions - the fidelity loops of train_CIGAR
(CIGAR.py:84-134), gen-2023 AR.compute_loss / forward (AR_AutoRegression.py:148-254), gen-2023 GAR.compute_loss /
forward (GAR_GeneralizedAutoAR.py:153-250), gen-2024 train_GAR (GAR.py:76-126) - on the drop-in operator modules and
compare every recorded quantity at the north-star tolerance (1e-9 relative, fp64).  Inputs come from the same seeded
recipes (oracle/l4_cases.py data_* functions)."""
import contextlib
import io

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-9


def _cases():
    import os
    import sys
    from conftest import ROOT
    p = os.path.join(ROOT, 'oracle')
    if p not in sys.path:
        sys.path.insert(0, p)
    import l4_cases
    return l4_cases


def _check_params(model, g, tol=TOL, prefix='p_'):
    lc = _cases()
    for name, p in model.named_parameters():
        key = prefix + name.replace('.', '_')
        if p.numel() > 8192:
            smp, nrm = lc._sample(p)
            assert rel_err(smp.cpu(), g[key + '_sample']) < tol, key
            assert rel_err(nrm.cpu(), g[key + '_norm']) < tol, key
        else:
            assert rel_err(p.detach().cpu().reshape(g[key].shape), g[key]) < tol, key


def _check_big(t, g, key, tol=TOL):
    lc = _cases()
    if key in g:
        assert rel_err(t.detach().cpu().reshape(g[key].shape), g[key]) < tol, key
    else:
        smp, nrm = lc._sample(t)
        assert rel_err(smp.cpu(), g[key + '_sample']) < tol, key
        assert rel_err(nrm.cpu(), g[key + '_norm']) < tol, key


def test_c1_full_size_AR2023_50_adam_steps():
    """g1: N = 100, d = 2, D = 64, 2 fidelities, 50 Adam steps (lr 0.01), prediction on the 100 held-out points."""
    from fidelityfusion_b200.MFGP_ver2023May import CIGP
    from fidelityfusion_b200.MFGP_ver2023May.multiscale_coupling.Residual import Residual
    g = load_golden('l4_ar2023_c1')
    x, xe, y0, y1 = (t.to(DEV) for t in _cases().data_c1())

    class AR(torch.nn.Module):                       # attribute names of the reference class => same parameter keys
        def __init__(self):
            super().__init__()
            self.cigp_list = torch.nn.ModuleList([CIGP(None), CIGP(None)])
            self.residual_list = torch.nn.ModuleList([Residual(None)])

    m = AR().double().to(DEV)
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    for it in range(50):
        opt.zero_grad()
        loss = m.cigp_list[0].compute_loss(x, y0) + \
            m.cigp_list[1].compute_loss(x, m.residual_list[0].forward(y0, y1), update_data=True)
        loss.backward()
        assert abs(loss.item() - g['losses'][it]) <= TOL * abs(g['losses'][it]), it
        if it == 0:
            for name, p in m.named_parameters():
                assert rel_err(p.grad.cpu().reshape(-1), g['g0_' + name.replace('.', '_')].reshape(-1)) < TOL, name
        opt.step()
    _check_params(m, g)
    with torch.no_grad():
        m0, v0 = m.cigp_list[0].forward(xe)
        m1, v1 = m.cigp_list[1].forward(xe)
        u, var = m.residual_list[0].backward(m0, m1), m.residual_list[0].var_backward(v0, v1)
    assert rel_err(u.cpu(), g['u']) < TOL and rel_err(var.cpu(), g['var']) < TOL


def test_c3_full_size_train_CIGAR_three_fidelities():
    """g3: CIGAR, N = (512, 256, 128), outputs 256 / 1024 / 4096 columns, 20 Adam steps per fidelity (lr 1e-3) with a
    fresh Adam over ALL parameters per fidelity (SURVEY A-14), N x N y_var, Tensor_linear trained through the residual;
    then CIGAR.forward's chain (incl. the `var_low.diag()` quirk, A-11) on 64 test points."""
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Tensor_linear
    from fidelityfusion_b200.FidelityFusion_Models.MF_data import MultiFidelityDataManager
    g = load_golden('l4_cigar3_c3')
    data, xt = _cases().data_cigar3(DEV)
    dm = MultiFidelityDataManager(data)
    shapes = [(256,), (1024,), (4096,)]

    class CIGAR(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gpr_list = torch.nn.ModuleList([cigp(ARDKernel(5), 1.0) for _ in range(3)])
            self.Tensor_linear_list = torch.nn.ModuleList([Tensor_linear(shapes[i], shapes[i + 1]) for i in range(2)])

        def forward(self, dm, x_test, to_fidelity=None):          # CIGAR.py:40-82
            level = 2 if to_fidelity is None else to_fidelity
            for f in range(level + 1):
                if f == 0:
                    xtr, ytr = dm.get_data(0, normal=True)
                    mean_low, var_low = self.gpr_list[0].forward(xtr, ytr, x_test)
                    var_low = var_low.diag().unsqueeze(1).expand_as(mean_low)
                    mean_high, var_high = mean_low, var_low
                else:
                    xtr, ytr = dm.get_data_by_name('res-{}'.format(f))
                    mean_res, _ = self.gpr_list[f].forward(xtr, ytr, x_test)
                    var_res = var_low.diag().unsqueeze(1).expand_as(mean_res)
                    mean_high = self.Tensor_linear_list[f - 1](mean_low) + mean_res
                    var_high = self.Tensor_linear_list[f - 1](var_low) + var_res
                    mean_low, var_low = mean_high, var_high
            return mean_high, var_high

    m = CIGAR().to(DEV)
    iters = 20
    for f in range(3):
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        if f == 0:
            xl, yl = dm.get_data(0, normal=True)
            for i in range(iters):
                opt.zero_grad()
                loss = -m.gpr_list[0].negative_log_likelihood(xl, yl)
                assert abs(loss.item() - g['losses'][f, i]) <= TOL * abs(g['losses'][f, i]), (f, i)
                loss.backward()
                opt.step()
        else:
            with torch.no_grad():
                sx, y_low, y_high = dm.get_nonsubset_fill_data(m, f - 1, f)
            for i in range(iters):
                opt.zero_grad()
                res_mean = y_high[0] - m.Tensor_linear_list[f - 1](y_low[0])
                res_var = abs(y_high[1] - y_low[1])
                if i == iters - 1:
                    dm.add_data(raw_fidelity_name='res-{}'.format(f), fidelity_index=None, x=sx.detach(),
                                y=[res_mean.detach(), res_var.detach()])
                loss = -m.gpr_list[f].negative_log_likelihood(sx, [res_mean, res_var])
                assert abs(loss.item() - g['losses'][f, i]) <= TOL * abs(g['losses'][f, i]), (f, i)
                loss.backward()
                opt.step()
    _check_params(m, g)
    with torch.no_grad():
        mean, var = m(dm, dm.normalizelayer[0].normalize_x(xt.to(DEV)))
    _check_big(mean, g, 'mean')
    _check_big(var, g, 'var')


def _gar2023(shape):
    from fidelityfusion_b200.MFGP_ver2023May import HOGP
    from fidelityfusion_b200.MFGP_ver2023May.multiscale_coupling.matrix import Matrix_Mapping

    class GAR(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.hogp_list = torch.nn.ModuleList([HOGP({'fidelity_shapes': [shape]}) for _ in range(2)])
            self.matrix_list = torch.nn.ModuleList([Matrix_Mapping({'low_fidelity_shape': shape, 'high_fidelity_shape': shape})])

        def compute_loss(self, x, y_list):                        # GAR_GeneralizedAutoAR.py:207-250, aligned inputs
            return self.hogp_list[0].compute_loss(x, y_list[0]) + \
                self.hogp_list[1].compute_loss(x, self.matrix_list[0].forward(y_list[0], y_list[1]), update_data=True)

        def forward(self, x):                                     # :153-177
            mean, var = self.hogp_list[0].forward(x)
            rm, rv = self.hogp_list[1].forward(x)
            return self.matrix_list[0].backward(mean, rm), self.matrix_list[0].var_backward(var, rv)

    return GAR()


def test_c4_full_size_GAR2023_loss_grads_steps_predict():
    """g4: gen-2023 GAR on 128 x (32 x 32 x 16), 2 aligned fidelities: first-step loss and ALL gradients (9 HOGP
    parameters x 2 + the three trainable mapping matrices + rho), 5 Adam steps, prediction on 32 points."""
    g = load_golden('l4_gar2023_c4')
    x, xt, ylo, yhi = (t.to(DEV) for t in _cases().data_c4())
    m = _gar2023(torch.Size([32, 32, 16])).double().to(DEV)
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    for it in range(5):
        opt.zero_grad()
        loss = m.compute_loss(x, [ylo, yhi])
        loss.backward()
        assert abs(loss.item() - g['losses'][it]) <= TOL * abs(g['losses'][it]), it
        if it == 0:
            lc = _cases()
            for name, p in m.named_parameters():
                if p.grad is None:
                    continue
                key = 'g0_' + name.replace('.', '_')
                a = p.grad if p.numel() <= 8192 else lc._sample(p.grad)[0]
                assert rel_err(a.cpu().reshape(g[key].shape), g[key]) < TOL, name
        opt.step()
    _check_params(m, g)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        u, var = m.forward(xt)
    _check_big(u, g, 'u')
    _check_big(var, g, 'var')


def test_c4_full_size_train_GAR_gen2024():
    """C4 through gen-2024: GAR.py:76-126 (HOGP_simple of two_fidelity_models per fidelity, Tensor_linear residual,
    8 Adam steps per fidelity, lr 1e-2) and GAR.forward (:39-74) on 32 points."""
    from fidelityfusion_b200.FidelityFusion_Models.two_fidelity_models.hogp_simple import HOGP_simple
    from fidelityfusion_b200.GaussianProcess.kernel import SquaredExponentialKernel
    from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Tensor_linear
    from fidelityfusion_b200.FidelityFusion_Models.MF_data import MultiFidelityDataManager
    g = load_golden('l4_gar2_c4')
    shape = (32, 32, 16)
    x, xt, ylo, yhi = (t.to(DEV) for t in _cases().data_c4())
    dm = MultiFidelityDataManager([{'raw_fidelity_name': '0', 'fidelity_indicator': 0, 'X': x, 'Y': ylo},
                                   {'raw_fidelity_name': '1', 'fidelity_indicator': 1, 'X': x, 'Y': yhi}])

    class GAR(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.hogp_list = torch.nn.ModuleList([HOGP_simple(kernel=SquaredExponentialKernel(), noise_variance=1.0,
                                                              output_shape=shape) for _ in range(2)])
            self.Tensor_linear_list = torch.nn.ModuleList([Tensor_linear(shape, shape)])

    m = GAR().double().to(DEV)
    iters = 8
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    xl, yl = dm.get_data(0, normal=True)
    for i in range(iters):
        opt.zero_grad()
        loss = m.hogp_list[0].log_likelihood(xl, yl)
        assert abs(loss.item() - g['losses'][0, i]) <= TOL * abs(g['losses'][0, i]), i
        loss.backward()
        opt.step()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    with torch.no_grad():
        sx, y_low, y_high = dm.get_nonsubset_fill_data(m, 0, 1)        # aligned inputs: no model.forward needed
    for i in range(iters):
        opt.zero_grad()
        res_mean = y_high[0] - m.Tensor_linear_list[0](y_low[0])
        res_var = abs(y_high[1] - y_low[1])
        if i == iters - 1:
            dm.add_data(raw_fidelity_name='res-1', fidelity_index=None, x=sx.detach(), y=[res_mean.detach(), res_var.detach()])
        loss = m.hogp_list[1].log_likelihood(sx, [res_mean, res_var])
        assert abs(loss.item() - g['losses'][1, i]) <= TOL * abs(g['losses'][1, i]), i
        loss.backward()
        opt.step()
    _check_params(m, g)
    with torch.no_grad():
        xtn = dm.normalizelayer[0].normalize_x(xt)
        xtr, _ = dm.get_data(0, normal=True)
        mean_low, var_low = m.hogp_list[0].forward(xtr, xtn)
        xtr, _ = dm.get_data_by_name('res-1')
        mean_res, var_res = m.hogp_list[1].forward(xtr, xtn)
        mean = m.Tensor_linear_list[0](mean_low) + mean_res
        var = m.Tensor_linear_list[0](var_low) + var_res
    _check_big(mean, g, 'mean')
    _check_big(var, g, 'var')
