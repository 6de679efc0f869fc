"""Device-side training loop: FusedAdam against torch.optim.Adam, the CUDA-graph trainer against the
eager loop (bit for bit) and against the CPU oracle trained with torch's own Adam (reference loop
GaussianProcess/cigp_v10.py:160-175).  All through the Python mirror -> C ABI."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _data(n, d, D=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 2
    w = torch.randn(d, D, generator=g, dtype=torch.float64)
    y = torch.sin(2 * x @ w) + 0.1 * torch.randn(n, D, generator=g, dtype=torch.float64)
    return x, y


def _model(d):
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    return cigp(ARDKernel(d), 1.0).double().cuda()


def _params(m):
    return torch.cat([p.detach().reshape(-1).cpu() for p in m.parameters()])


def _oracle_training(x, y, d, iters, lr):
    """The reference's loop on the CPU: oracle loss (torch CPU fp64 + autograd) and torch.optim.Adam."""
    from oracle import ff_oracle as O
    ls = torch.ones(d, dtype=torch.float64, requires_grad=True)
    sv = torch.ones(1, dtype=torch.float64, requires_grad=True)
    lb = torch.ones(1, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([ls, sv, lb], lr=lr)
    curve = []
    for _ in range(iters):
        opt.zero_grad()
        loss = -O.cigp_log_likelihood(O.ard_kernel(x, x, ls, sv), lb, y)
        loss.backward()
        opt.step()
        curve.append(float(loss))
    return torch.cat([ls.detach(), sv.detach(), lb.detach()]), torch.tensor(curve, dtype=torch.float64)


def test_fused_adam_matches_torch_adam():
    from fidelityfusion_b200.training import FusedAdam
    x, y = _data(200, 3, 2)
    xc, yc = x.cuda(), y.cuda()
    ma, mb = _model(3), _model(3)
    oa = torch.optim.Adam(ma.parameters(), lr=0.05)
    ob = FusedAdam(mb.parameters(), lr=0.05, history=32)
    for _ in range(30):
        for m, o in ((ma, oa), (mb, ob)):
            o.zero_grad()
            loss = -m.negative_log_likelihood(xc, yc)
            loss.backward()
            if o is ob:
                o.step(loss=loss)
            else:
                o.step()
    pa, pb = _params(ma), _params(mb)
    assert float((pa - pb).abs().max() / pa.abs().max()) < 1e-11
    assert ob.losses(30).shape == (30,) and float(ob.losses(30)[-1]) == pytest.approx(float(loss.detach()), rel=1e-12)
    # parameters without a gradient are skipped, like torch's Adam (CIGAR.py:97: one optimiser over ALL parameters)
    extra = torch.nn.Parameter(torch.ones(3, dtype=torch.float64, device='cuda'))
    oc = FusedAdam(list(mb.parameters()) + [extra], lr=0.1)
    oc.zero_grad()
    (-mb.negative_log_likelihood(xc, yc)).backward()
    oc.step()
    assert torch.equal(extra.detach().cpu(), torch.ones(3, dtype=torch.float64)) and 'step' not in oc.state[extra]


def test_fused_adam_rejects_what_it_cannot_update():
    from fidelityfusion_b200.training import FusedAdam
    p32 = torch.nn.Parameter(torch.ones(2, device='cuda', dtype=torch.float32))
    p32.grad = torch.ones_like(p32)
    with pytest.raises(TypeError):
        FusedAdam([p32]).step()
    with pytest.raises(NotImplementedError):
        FusedAdam([p32], weight_decay=0.1)


@pytest.mark.parametrize('n,d', [(150, 2), (1024, 5)])      # 1024: the look-ahead side streams are part of the capture
def test_graphed_trainer_equals_eager_loop_bit_for_bit(n, d):
    from fidelityfusion_b200.training import GraphedTrainer, train
    x, y = _data(n, d)
    xc, yc = x.cuda(), y.cuda()
    iters = 12
    me, mg = _model(d), _model(d)
    curve_e = train(lambda: -me.negative_log_likelihood(xc, yc), me.parameters(), iters, lr=0.02, graphed=False)
    tr = GraphedTrainer(lambda: -mg.negative_log_likelihood(xc, yc), mg.parameters(), lr=0.02, history=iters)
    assert torch.equal(_params(mg), _params(_model(d)))        # warm-up iterations were rolled back
    tr.run(5).run(iters - 5)
    assert torch.equal(_params(me), _params(mg))
    assert torch.equal(curve_e.cpu(), tr.losses().cpu())
    assert tr.losses().shape == (iters,)


def test_graphed_training_follows_the_reference_loop_on_the_cpu_oracle():
    from fidelityfusion_b200.training import train
    n, d, iters, lr = 120, 2, 25, 0.03
    x, y = _data(n, d, 1, seed=3)
    m = _model(d)
    xc, yc = x.cuda(), y.cuda()
    curve = train(lambda: -m.negative_log_likelihood(xc, yc), m.parameters(), iters, lr=lr)
    p_ref, curve_ref = _oracle_training(x, y, d, iters, lr)
    p = torch.cat([m.kernel.length_scales.detach().cpu(), m.kernel.signal_variance.detach().cpu().reshape(-1),
                   m.log_beta.detach().cpu()])
    assert float((p - p_ref).abs().max() / p_ref.abs().max()) < 1e-9
    assert float(((curve.cpu() - curve_ref).abs() / curve_ref.abs()).max()) < 1e-9


def test_graphed_trainer_reports_a_non_pd_covariance():
    from fidelityfusion_b200.training import GraphedTrainer
    x, y = _data(64, 2)
    xc, yc = x.cuda(), y.cuda()
    m = _model(2)
    bad_var = -5.0 * torch.eye(64, dtype=torch.float64, device='cuda')      # Sigma + diag(y_var) is indefinite
    with pytest.raises(torch.linalg.LinAlgError):
        GraphedTrainer(lambda: -m.negative_log_likelihood(xc, [yc, bad_var]), m.parameters(), lr=0.01).run(2)


def test_graphed_trainer_on_the_kronecker_model():
    """gen-2023 HOGP.compute_loss (hogp.py:140-198) trained by graph replay equals the eager loop."""
    from fidelityfusion_b200.MFGP_ver2023May import HOGP
    from fidelityfusion_b200.training import GraphedTrainer, train
    g = torch.Generator().manual_seed(7)
    x = torch.rand(24, 3, generator=g, dtype=torch.float64).cuda()
    Y = torch.randn(24, 6, 5, generator=g, dtype=torch.float64).cuda()
    def make():
        torch.manual_seed(0)
        return HOGP({'fidelity_shapes': [torch.Size([6, 5])]}).double().cuda()
    he, hg = make(), make()
    curve_e = train(lambda: he.compute_loss(x, Y), he.parameters(), 8, lr=0.01, graphed=False)
    tr = GraphedTrainer(lambda: hg.compute_loss(x, Y), hg.parameters(), lr=0.01, history=8).run(8)
    assert torch.equal(_params(he), _params(hg))
    assert torch.equal(curve_e.cpu(), tr.losses().cpu())


def test_prediction_after_graphed_training_uses_the_trained_parameters():
    """cigp.forward keeps (L^-1, alpha) resident between calls, keyed on the parameters' version counters
    (ops.state_token); training by graph replay must invalidate it like torch's in-place optimiser update would."""
    from fidelityfusion_b200.training import GraphedTrainer
    x, y = _data(90, 2, seed=5)
    xs = _data(11, 2, seed=6)[0].cuda()
    xc, yc = x.cuda(), y.cuda()
    m = _model(2)
    with torch.no_grad():
        before = m(xc, yc, xs)[0].clone()
    GraphedTrainer(lambda: -m.negative_log_likelihood(xc, yc), m.parameters(), lr=0.05).run(10)
    fresh = _model(2)
    fresh.load_state_dict(m.state_dict())
    with torch.no_grad():
        after, want = m(xc, yc, xs)[0], fresh(xc, yc, xs)[0]
    assert torch.equal(after, want) and not torch.equal(after, before)
