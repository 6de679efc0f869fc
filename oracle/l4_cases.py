"""The reference's own L4 entry points (train_CIGAR / train_AR / train_GAR / train_ResGP / train_NAR of
FidelityFusion_Models/, and gen-2023 AR / GAR / CIGAR .compute_loss + .forward of MFGP_ver2023May/) driven on the
synthetic BASELINE configs of SURVEY.md 8(d).

TEST INFRASTRUCTURE.  The case functions import the L4 classes by their REFERENCE module names, so they run
  * the unmodified reference on torch-CPU  (oracle/gen_golden_l4.py -> tests/golden/l4_*.npz), and
  * the unmodified reference L4 code on the CUDA drop-ins after fidelityfusion_b200.binding.install()
    (tests/test_binding.py, tools/run_l4_on_gpu.py),
with identical inputs.  Nothing here is used by the product.

Every case returns {name: tensor/array}: per-iteration losses, final parameters (large coupling matrices as a fixed
strided sample + Frobenius norm), predictions."""
import contextlib
import io
import math

import numpy as np
import torch


class Recorder:
    """Stands in for Experiments.log_debugger.log_debugger: train_*() call get_status(model, optimizer, i, loss)
    before every backward (CIGAR.py:102-103); we only record the loss."""

    def __init__(self):
        self.losses = []

    def get_status(self, model, optimizer, epoch, loss):
        self.losses.append(float(loss.detach().reshape(-1)[0].item()))


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _sample(t, k=4096):
    """A fixed strided sample of a big tensor + its norm (keeps the fixtures small)."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // k)
    return f[::step][:k].clone(), f.norm().reshape(1)


def _put(out, key, t):
    """Store t whole if small, else as a strided sample + norm."""
    if t.numel() > 8192:
        out[key + '_sample'], out[key + '_norm'] = _sample(t)
    else:
        out[key] = t.detach().clone()


def _params(model, out, prefix='p_'):
    for name, p in model.named_parameters():
        key = prefix + name.replace('.', '_')
        if p.numel() > 8192:
            out[key + '_sample'], out[key + '_norm'] = _sample(p)
        else:
            out[key] = p.detach().clone()


def sep_field(x, shape, modes=4, phase=0.0, scale=1.0):
    """Smooth field on a regular grid: sum of `modes` separable sine modes with x-dependent amplitudes."""
    grids = [torch.linspace(0, 1, s, dtype=x.dtype, device=x.device) for s in shape]
    out = 0
    for r in range(modes):
        comp = 1.0 + scale * (r + 1) * 0.5 * x[:, r % x.shape[1]] + 0.3 * torch.sin(3.0 * x[:, (r + 1) % x.shape[1]] + r)
        for k, gr in enumerate(grids):
            comp = comp.unsqueeze(-1) * torch.sin((r + 1 + k) * math.pi * gr + 0.7 * r + phase)
        out = out + comp
    return out


# ------------------------------------------------------------------------------------------------------------------
# gen-2024  (FidelityFusion_Models/)
# ------------------------------------------------------------------------------------------------------------------
def data_cigar3(dev, sizes=(512, 256, 128), grids=(16, 32, 64), n_test=64):
    """C3 inputs (SURVEY 8d recipe): nested subsets of U[0,1]^{512x5}, smooth separable fields on 16^2/32^2/64^2 grids."""
    g = torch.Generator(device='cpu').manual_seed(3)
    d = 5
    x_all = torch.rand(sizes[0], d, generator=g, device='cpu')
    xt = torch.rand(n_test, d, generator=g, device='cpu')
    data = []
    for f, (n, s) in enumerate(zip(sizes, grids)):
        xf = x_all[:n]
        yf = (1.0 + 0.1 * f) * sep_field(xf, (s, s)) + 0.05 * f * sep_field(xf, (s, s), modes=2, phase=1.3)
        data.append({'raw_fidelity_name': str(f), 'fidelity_indicator': f, 'X': xf.to(dev), 'Y': yf.reshape(n, -1).to(dev)})
    return data, xt


def case_cigar3_c3(dev, max_iter=20, lr=1e-3, sizes=(512, 256, 128), grids=(16, 32, 64), n_test=64):
    """BASELINE config 3: CIGAR on a 64x64 field (4096 output columns), 3 fidelities with different output sizes
    (16^2 / 32^2 / 64^2, flattened), nested inputs, if_nonsubset=True (the only branch of train_CIGAR that runs,
    SURVEY 8c), unmodified train_CIGAR (CIGAR.py:84-134) + CIGAR.forward (:40-82)."""
    from FidelityFusion_Models.CIGAR import CIGAR, train_CIGAR
    from FidelityFusion_Models.MF_data import MultiFidelityDataManager
    import GaussianProcess.kernel as kernel
    d = 5
    data, xt = data_cigar3(dev, sizes, grids, n_test)
    dm = MultiFidelityDataManager(data)
    shapes = [(s * s,) for s in grids]
    model = CIGAR(len(sizes), [kernel.ARDKernel(d) for _ in sizes], shapes, if_nonsubset=True).to(dev)
    rec = Recorder()
    _quiet(train_CIGAR, model, dm, max_iter=max_iter, lr_init=lr, debugger=rec)
    with torch.no_grad():
        xtn = dm.normalizelayer[0].normalize_x(xt.to(dev))
        mean, var = model(dm, xtn)
    out = {'losses': torch.tensor(rec.losses).reshape(len(sizes), max_iter)}
    _put(out, 'mean', mean)
    _put(out, 'var', var)
    _params(model, out)
    return out


def _mf_scalar_data(g, d, sizes, shared, dev):
    """D = 1 multi-fidelity toy data whose fidelity-f inputs share only `shared[f]` points with fidelity f-1 (the
    rest are new points): exercises get_nonsubset_fill_data's mixed branch (MF_data.py:286-303)."""
    f_lo = lambda x: torch.sin(4.0 * x.sum(1, keepdim=True)) + 0.3 * x[:, :1]
    data, prev = [], None
    for f, n in enumerate(sizes):
        if prev is None:
            xf = torch.rand(n, d, generator=g, device='cpu')
        else:
            keep = prev[torch.randperm(prev.shape[0], generator=g, device='cpu')[:shared[f]]]
            xf = torch.cat([keep, torch.rand(n - shared[f], d, generator=g, device='cpu')], 0)
            xf = xf[torch.randperm(n, generator=g, device='cpu')]
        yf = (1.0 + 0.25 * f) * f_lo(xf) + 0.1 * f * torch.cos(5.0 * xf[:, 1:2]) + 0.01 * torch.randn(n, 1, generator=g, device='cpu')
        data.append({'raw_fidelity_name': str(f), 'fidelity_indicator': f, 'X': xf.to(dev), 'Y': yf.to(dev)})
        prev = xf
    return data


def _train_cigp_family(dev, cls_name, train_name, module, sizes, shared, max_iter, lr, seed, kernel_kind='SE', **ctor):
    import importlib
    mod = importlib.import_module(module)
    Model, train = getattr(mod, cls_name), getattr(mod, train_name)
    from FidelityFusion_Models.MF_data import MultiFidelityDataManager
    import GaussianProcess.kernel as kernel
    g = torch.Generator(device='cpu').manual_seed(seed)
    d = 2
    data = _mf_scalar_data(g, d, sizes, shared, dev)
    xt = torch.rand(50, d, generator=g, device='cpu')
    dm = MultiFidelityDataManager(data)
    if cls_name == 'NAR':
        # the fidelity-f GP of NAR sees [x, y_low] (NAR.py:97): input_dim grows by one per level
        kl = [kernel.ARDKernel(d + (1 if f else 0)) for f in range(len(sizes))]
    elif kernel_kind == 'SE':
        kl = [kernel.SquaredExponentialKernel() for _ in sizes]
    else:
        kl = [kernel.ARDKernel(d) for _ in sizes]
    model = Model(len(sizes), kl, if_nonsubset=True, **ctor).to(dev)
    rec = Recorder()
    _quiet(train, model, dm, max_iter=max_iter, lr_init=lr, debugger=rec)
    with torch.no_grad():
        xtn = dm.normalizelayer[0].normalize_x(xt.to(dev))
        mean, cov = model(dm, xtn)
    out = {'losses': torch.tensor(rec.losses).reshape(len(sizes), max_iter), 'mean': mean, 'cov': cov}
    _params(model, out)
    return out


def case_ar3_nonsubset(dev):
    """train_AR (AR_autoRegression.py:92-140) with 3 fidelities whose inputs only partly overlap: the non-subset fill
    (MF_data.py:253-303) calls AR.forward for the missing low-fidelity outputs and passes their posterior covariance
    on as y_var; rho trains through the residual."""
    return _train_cigp_family(dev, 'AR', 'train_AR', 'FidelityFusion_Models.AR_autoRegression', (120, 60, 30),
                              (None, 40, 18), 30, 1e-2, seed=21, rho_init=1.0)


def case_resgp2_nonsubset(dev):
    """train_ResGP (ResGP.py:67-112), 2 fidelities, partly overlapping inputs."""
    return _train_cigp_family(dev, 'ResGP', 'train_ResGP', 'FidelityFusion_Models.ResGP', (90, 45), (None, 30), 20, 1e-2,
                              seed=22, kernel_kind='ARD')


def case_nar2_nonsubset(dev):
    """train_NAR (NAR.py:63-110), 2 fidelities, partly overlapping inputs; the high-fidelity GP's input is [x, y_low]."""
    return _train_cigp_family(dev, 'NAR', 'train_NAR', 'FidelityFusion_Models.NAR', (90, 45), (None, 30), 20, 1e-2, seed=23)


def data_c4(N=128, shape=(32, 32, 16), n_test=32):
    """C4 inputs (SURVEY 8d recipe): N = 128, d = 5, smooth 32x32x16 fields, Y_hi = 1.1 Y_lo + 0.05 field_2 (CPU tensors)."""
    g = torch.Generator(device='cpu').manual_seed(4)
    d = 5
    x = torch.rand(N, d, generator=g, device='cpu')
    xt = torch.rand(n_test, d, generator=g, device='cpu')
    ylo = sep_field(x, shape)
    yhi = 1.1 * ylo + 0.05 * sep_field(x, shape, modes=3, phase=0.9, scale=2.0)
    return x, xt, ylo, yhi


def case_gar2_c4(dev, max_iter=8, lr=1e-2, N=128, shape=(32, 32, 16), n_test=32):
    """BASELINE config 4 through gen-2024: GAR (GAR.py:13-126) on 32x32x16 tensor outputs, 2 aligned fidelities,
    HOGP_simple (two_fidelity_models copy) per fidelity, Tensor_linear coupling, unmodified train_GAR."""
    from FidelityFusion_Models.GAR import GAR, train_GAR
    from FidelityFusion_Models.MF_data import MultiFidelityDataManager
    import GaussianProcess.kernel as kernel
    x, xt, ylo, yhi = data_c4(N, shape, n_test)
    data = [{'raw_fidelity_name': '0', 'fidelity_indicator': 0, 'X': x.to(dev), 'Y': ylo.to(dev)},
            {'raw_fidelity_name': '1', 'fidelity_indicator': 1, 'X': x.to(dev), 'Y': yhi.to(dev)}]
    dm = MultiFidelityDataManager(data)
    model = GAR(2, [kernel.SquaredExponentialKernel() for _ in range(2)], [shape, shape], if_nonsubset=True).double().to(dev)
    rec = Recorder()
    _quiet(train_GAR, model, dm, max_iter=max_iter, lr_init=lr, debugger=rec)
    with torch.no_grad():
        xtn = dm.normalizelayer[0].normalize_x(xt.to(dev))
        mean, var = model(dm, xtn)
    out = {'losses': torch.tensor(rec.losses).reshape(2, max_iter)}
    _put(out, 'mean', mean)
    _put(out, 'var', var)
    _params(model, out)
    return out


# ------------------------------------------------------------------------------------------------------------------
# gen-2023  (MFGP_ver2023May/)
# ------------------------------------------------------------------------------------------------------------------
def _adam_loop(model, loss_fn, steps, lr):
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses, g0 = [], None
    for it in range(steps):
        opt.zero_grad()
        loss = loss_fn()
        loss.backward()
        if it == 0:
            g0 = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        losses.append(float(loss.item()))
        opt.step()
    return torch.tensor(losses), g0


def data_c1():
    """C1 inputs (SURVEY 8d recipe, stand-in for the toy data missing from the tree): x ~ U[0,1]^{200x2}, first 100
    train / last 100 eval, y_lo = sin(2 pi x w), y_hi = 1.2 y_lo + 0.1 cos(3 x_0), z-normalised (CPU tensors)."""
    g = torch.Generator(device='cpu').manual_seed(1)
    xa = torch.rand(200, 2, generator=g, device='cpu')
    w = torch.randn(2, 64, generator=g, device='cpu')
    ylo = torch.sin(2 * math.pi * xa @ w)
    yhi = 1.2 * ylo + 0.1 * torch.cos(3 * xa[:, :1])
    norm = lambda t, ref: (t - ref.mean()) / ref.std()                   # z-normalisation, mfgp_demo.py:25-33 style
    return (norm(xa[:100], xa[:100]), norm(xa[100:], xa[:100]), norm(ylo[:100], ylo[:100]), norm(yhi[:100], yhi[:100]))


def case_ar2023_c1(dev, steps=50, lr=0.01):
    """BASELINE config 1 (SURVEY 8d recipe): gen-2023 AR, 2 fidelities, N = 100, d = 2, D = 64, the epoch loop of
    mfgp_demo.py:122-127 (Adam lr 0.01, 50 steps on AR.compute_loss, AR_AutoRegression.py:206-254), then AR.forward
    on the 100 held-out points."""
    from MFGP_ver2023May import AR
    x, xe, y0, y1 = (t.to(dev) for t in data_c1())
    m = AR({'fidelity_shapes': [(64,), (64,)]}).double().to(dev)
    losses, g0 = _quiet(_adam_loop, m, lambda: m.compute_loss(x, [y0, y1]), steps, lr)
    with torch.no_grad():
        u, v = _quiet(m, xe)
    out = {'losses': losses, 'u': u, 'var': v}
    out.update({'g0_' + k.replace('.', '_'): t for k, t in g0.items()})
    _params(m, out)
    return out


def case_gar2023_c4(dev, steps=5, lr=0.01, N=128, shape=(32, 32, 16), n_test=32):
    """BASELINE config 4 (SURVEY 8d recipe): gen-2023 GAR({'fidelity_shapes': [32x32x16] * 2}), N = 128, d = 5:
    GAR.compute_loss (GAR_GeneralizedAutoAR.py:207-250: two HOGPs + Matrix_Mapping residual), all gradients of the
    first step, 5 Adam steps, GAR.forward on 32 points."""
    from MFGP_ver2023May import GAR
    x, xt, ylo, yhi = data_c4(N, shape, n_test)
    x, xt, ylo, yhi = x.to(dev), xt.to(dev), ylo.to(dev), yhi.to(dev)
    m = GAR({'fidelity_shapes': [torch.Size(shape)] * 2}).double().to(dev)
    losses, g0 = _quiet(_adam_loop, m, lambda: m.compute_loss(x, [ylo, yhi]), steps, lr)
    with torch.no_grad():
        u, v = _quiet(m, xt)
    out = {'losses': losses, 'u_sample': _sample(u)[0], 'u_norm': _sample(u)[1], 'var_sample': _sample(v)[0],
           'var_norm': _sample(v)[1]}
    for k, t in g0.items():
        key = 'g0_' + k.replace('.', '_')
        out[key] = t if t.numel() <= 8192 else _sample(t)[0]
    _params(m, out)
    return out


def case_cigar2023(dev, steps=10, lr=0.01):
    """gen-2023 CIGAR (CIGAR_ConditionalIndependentGAR.py:211-255), 2 fidelities of equal output shape (the aliased
    mapping config of SURVEY A-7 forbids more): CIGP per fidelity on the flattened field + Matrix_Mapping residual."""
    from MFGP_ver2023May import CIGAR
    g = torch.Generator(device='cpu').manual_seed(6)
    N, d, shape = 96, 3, (12, 10)
    x = torch.rand(N, d, generator=g, device='cpu')
    xt = torch.rand(20, d, generator=g, device='cpu')
    ylo = sep_field(x, shape, modes=3).reshape(N, -1)                    # CIGP wants [N, D] (cigp.py:114)
    yhi = 0.9 * ylo + 0.1 * sep_field(x, shape, modes=2, phase=0.4).reshape(N, -1)
    x, xt, ylo, yhi = x.to(dev), xt.to(dev), ylo.to(dev), yhi.to(dev)
    m = CIGAR({'fidelity_shapes': [torch.Size([ylo.shape[1]])] * 2}).double().to(dev)
    losses, g0 = _quiet(_adam_loop, m, lambda: m.compute_loss(x, [ylo, yhi]), steps, lr)
    with torch.no_grad():
        u, v = _quiet(m, xt)
    out = {'losses': losses, 'u': u, 'var': v}
    out.update({'g0_' + k.replace('.', '_'): t for k, t in g0.items()})
    _params(m, out)
    return out


def case_family2023(dev, steps=10, lr=0.01):
    """The other gen-2023 multi-fidelity callers on the C1 data (2 fidelities, N = 100, d = 2, D = 64): ResGP
    (ResGP.py:200-245 compute_loss, :145-198 forward), NAR (NAR_NonlinearAR.py:188-233, :133-186: the high-fidelity CIGP's
    input is [x, y_low]) and CAR (CAR_ContinuAR.py:189-233, :134-187: FIDES + Kernel_res on the residual) - compute_loss,
    first-step gradients, 10 Adam steps, forward on the 100 held-out points."""
    import MFGP_ver2023May as G23
    x, xe, y0, y1 = (t.to(dev) for t in data_c1())
    out = {}
    for name in ('ResGP', 'NAR', 'CAR'):
        torch.manual_seed(7)                                   # Kernel_res reseeds the global generator itself (App. A-13)
        m = getattr(G23, name)({'fidelity_shapes': [(64,), (64,)]}).double().to(dev)
        losses, g0 = _quiet(_adam_loop, m, lambda: m.compute_loss(x, [y0, y1]), steps, lr)
        with torch.no_grad():
            u, v = _quiet(m, xe)
        out.update({f'{name}_losses': losses, f'{name}_u': u, f'{name}_var': v})
        out.update({f'{name}_g0_' + k.replace('.', '_'): t for k, t in g0.items()})
        _params(m, out, prefix=f'{name}_p_')
    return out


def case_bo_cigp_acq(dev, steps=30, lr=1e-2, n=48, n_cand=64):
    """One single-fidelity Bayesian-optimisation step built from the reference's own pieces (SURVEY 8f rank 2 callers):
    Bayesian_optimization/cigp.py CIGP_withMean - unmodified: ARDKernel + gp_computation_pack.Gaussian_log_likelihood
    (log_likelihood, :86-91) / conditional_Gaussian (forward, :61-84), input / output normalisation in torch - trained with
    Adam, then Bayesian_optimization/acq.py UCB / EI / PI (:118-231) on its posterior at a candidate set, and the arg-max
    candidate of each score."""
    from Bayesian_optimization.cigp import CIGP_withMean
    from Bayesian_optimization import acq as A
    import GaussianProcess.kernel as kernel
    g = torch.Generator(device='cpu').manual_seed(31)
    xtr = (torch.rand(n, 2, generator=g, device='cpu') * 4 - 2).to(dev)
    ytr = (torch.sin(2 * xtr[:, :1]) * torch.cos(xtr[:, 1:]) + 0.05 * torch.randn(n, 1, generator=g, device='cpu').to(dev))
    cand = (torch.rand(n_cand, 2, generator=g, device='cpu') * 4 - 2).to(dev)
    model = CIGP_withMean(2, 1, kernel=kernel.ARDKernel(2), noise_variance=0.5).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        loss = -model.log_likelihood(xtr, ytr)
        losses.append(float(loss.detach().reshape(-1)[0].item()))
        loss.backward()
        opt.step()
    out = {'losses': torch.tensor(losses, dtype=torch.float64)}
    _params(model, out)
    with torch.no_grad():
        mean_f = lambda X: model.forward(xtr, ytr, X)[0]
        var_f = lambda X: model.forward(xtr, ytr, X)[1]
        mu, var = model.forward(xtr, ytr, cand)
        f_best = float(ytr.max())
        ucb = _quiet(A.UCB(mean_f, var_f, kappa=2.0).forward, cand)
        ei = _quiet(A.EI(mean_f, var_f, xi=0.01).forward, cand, f_best)
        pi = _quiet(A.PI(mean_f, var_f, sita=0.01).forward, cand, f_best)
    out.update(mu=mu, var=var, ucb=ucb, ei=ei, pi=pi.to(torch.float64),
               next_ucb=cand[torch.argmax(ucb)], next_ei=cand[torch.argmax(ei)])
    return out


CASES = {
    'l4_cigar3_c3': case_cigar3_c3,
    'l4_ar3_nonsubset': case_ar3_nonsubset,
    'l4_resgp2_nonsubset': case_resgp2_nonsubset,
    'l4_nar2_nonsubset': case_nar2_nonsubset,
    'l4_gar2_c4': case_gar2_c4,
    'l4_ar2023_c1': case_ar2023_c1,
    'l4_gar2023_c4': case_gar2023_c4,
    'l4_cigar2023': case_cigar2023,
    'l4_bo_cigp_acq': case_bo_cigp_acq,
    'l4_family2023': case_family2023,
}


def to_numpy(out):
    res = {}
    for k, v in out.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        res[k] = np.asarray(v)
    return res
