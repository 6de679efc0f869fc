"""Host-side logic that needs no GPU: module surface / state_dict keys mirror the reference, loud failure
without CUDA, batching normalisation, result packing and shard ranges."""
import pytest
import torch

from fidelityfusion_b200 import _lib, batched, ops


def test_state_dict_keys_match_reference_modules():
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.gp_basic import GP_basic
    from fidelityfusion_b200.GaussianProcess.hogp_simple import HOGP_simple
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel, SquaredExponentialKernel
    from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Tensor_linear
    from fidelityfusion_b200.MFGP_ver2023May import CIGP, HOGP
    from fidelityfusion_b200.MFGP_ver2023May.multiscale_coupling.matrix import Matrix_Mapping
    assert list(cigp(ARDKernel(3), 1.0).state_dict()) == ['log_beta', 'kernel.length_scales', 'kernel.signal_variance']
    assert list(GP_basic(SquaredExponentialKernel(), 0.1).state_dict()) == ['noise_variance', 'kernel.length_scale', 'kernel.signal_variance']
    assert list(CIGP(None).state_dict()) == ['noise_box.value', 'kernel.length_scale', 'kernel.scale']
    h = HOGP({'fidelity_shapes': [torch.Size([3, 2])]})
    assert list(h.state_dict()) == ['noise_box.value', 'kernel_list.0.length_scale', 'kernel_list.0.scale',
                                    'kernel_list.1.length_scale', 'kernel_list.1.scale', 'kernel_list.2.length_scale',
                                    'kernel_list.2.scale', 'grid.0', 'grid.1', 'mapping_vector.0', 'mapping_vector.1']
    assert [p.requires_grad for p in h.grid] == [False, False]
    hs = HOGP_simple(SquaredExponentialKernel(), 1.0, [3, 2])
    assert 'noise_variance' in hs.state_dict() and 'kernel_list.0.length_scale' in hs.state_dict()
    assert hs.kernel_list[0] is hs.kernel_list[2]                       # ONE kernel object shared by all modes
    assert list(Tensor_linear([4], [8]).state_dict()) == ['vectors.0']
    mm = Matrix_Mapping({'low_fidelity_shape': (4,), 'high_fidelity_shape': (8,)})
    assert sorted(mm.state_dict()) == ['rho', 'vectors.0'] and mm.rho.dtype == torch.float32 and not mm.rho.requires_grad


def test_reference_defaults():
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    from fidelityfusion_b200.MFGP_ver2023May import CIGP
    from fidelityfusion_b200.MFGP_ver2023May.kernel.SE_kernel import SE_kernel
    k = ARDKernel(4, initial_length_scale=2.0)
    assert k.eps == 1e-9 and torch.all(k.length_scales == 2.0) and k.signal_variance.shape == (1,)
    c = CIGP(None)
    assert c.kernel.noise_exp_format is not True           # dict config => linear format (kernel_utils.py:12)
    assert float(c.noise_box.value) == 0.0 and c.noise_box.format == 'exp'
    assert float(SE_kernel(True, 2.0, 3.0).length_scale) == pytest.approx(torch.log(torch.tensor(2.0)).item())


def test_no_cpu_fallback():
    """The product fails loudly on CPU tensors instead of silently computing somewhere else."""
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    m = cigp(ARDKernel(2), 1.0)
    with pytest.raises(_lib.FFGPError, match='CUDA tensors only'):
        m.negative_log_likelihood(torch.rand(5, 2), torch.rand(5, 1))
    from fidelityfusion_b200 import tensorly_compat as tl
    with pytest.raises(_lib.FFGPError):
        tl.mode_dot(torch.rand(3, 4), torch.rand(2, 4), 1)


def test_product_does_not_import_the_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'fidelityfusion_b200')
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_norm_batch_shapes():
    x, y = torch.rand(5, 2), torch.rand(5, 3)
    out = ops._norm_batch(x, y, torch.rand(2), torch.rand(1), torch.tensor(0.3), None)
    _, yb, il, amp, dg, _, n, d, D, batch, pb, squeeze = out
    assert (n, d, D, batch, pb, squeeze) == (5, 2, 3, 1, 0, True) and dg.shape == (5,) and yb.shape == (1, 5, 3)
    xb, yb = torch.rand(4, 5, 2), torch.rand(4, 5, 1)
    out = ops._norm_batch(xb, yb, torch.rand(4, 2), torch.rand(1), torch.rand(5), None)
    _, _, il, amp, dg, _, n, d, D, batch, pb, squeeze = out
    assert pb == 1 and amp.shape == (4,) and dg.shape == (4, 5) and il.shape == (4, 2) and not squeeze
    out = ops._norm_batch(xb, yb, torch.rand(1), torch.rand(1), None, None)      # scalar length scale broadcast over d
    assert out[2].shape == (2,) and out[10] == 0


def test_shard_ranges_cover_everything_once():
    for total in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            rs = [batched.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    res = {'nll': torch.rand(5), 'g_length_scales': torch.rand(5, 3), 'mean': torch.rand(5, 4, 2), 'var': torch.rand(5, 4)}
    keys = ['nll', 'g_length_scales', 'mean', 'var']
    buf = batched.pack_results(res, keys)
    assert buf.shape == (5, 1 + 3 + 8 + 4)
    back = batched.unpack_results(buf, {k: tuple(res[k].shape[1:]) for k in keys}, keys)
    assert all(torch.equal(back[k], res[k]) for k in keys)


def test_factor_cache_token_logic():
    c = ops.FactorCache()
    buf, reuse = c.workspace(64, torch.device('cpu'), ('a', 1))
    assert reuse == 0
    c.mark_valid()
    assert c.workspace(64, torch.device('cpu'), ('a', 1))[1] == 1
    assert c.workspace(64, torch.device('cpu'), ('a', 2))[1] == 0       # parameters changed
    c.mark_valid()
    assert c.workspace(128, torch.device('cpu'), ('a', 2))[1] == 0      # grew: buffer reallocated
    c.mark_valid(); c.invalidate()
    assert c.workspace(128, torch.device('cpu'), ('a', 2))[1] == 0


def test_next_row_modules_fail_loudly_without_cuda():
    """SURVEY 8f rows (training loop, data matching, FIDES) follow the same rule as the hot path: no CPU fallback."""
    from fidelityfusion_b200 import data_match
    from fidelityfusion_b200.MFGP_ver2023May import FIDES
    from fidelityfusion_b200.training import FusedAdam, GraphedTrainer
    p = torch.nn.Parameter(torch.ones(3, dtype=torch.float64))
    p.grad = torch.ones_like(p)
    with pytest.raises(TypeError, match='CUDA'):
        FusedAdam([p], lr=0.1).step()
    with pytest.raises(_lib.FFGPError):
        GraphedTrainer(lambda: (p * p).sum(), [p])
    with pytest.raises(NotImplementedError):
        FusedAdam([p], amsgrad=True)
    with pytest.raises(_lib.FFGPError):
        data_match.row_match(torch.rand(4, 2, dtype=torch.float64), torch.rand(3, 2, dtype=torch.float64))
    with pytest.raises(ValueError):
        data_match.row_match(torch.rand(4, 2), torch.rand(3, 5))
    f = FIDES({}).double()
    assert list(f.state_dict()) == ['noise_box.value', 'kernel.length_scale', 'kernel.scale', 'kernel.length_scale_z', 'kernel.b']
    assert f.kernel.noise_exp_format is not True and f.kernel.seed == 1024 and f.forward(torch.rand(2, 2)) is None
    f.set_fidelity(0., 1., 0., 2.)
    with pytest.raises(_lib.FFGPError):
        f.compute_loss(torch.rand(5, 2, dtype=torch.float64), torch.rand(5, 1, dtype=torch.float64))


def test_fused_adam_table_cache_is_bounded_and_capture_safe_api():
    """Host-only behaviour of FusedAdam that does not touch the device: hyper-parameter validation and defaults."""
    from fidelityfusion_b200.training import FusedAdam
    p = torch.nn.Parameter(torch.ones(2, dtype=torch.float64))
    o = FusedAdam([p], lr=0.01)
    g = o.param_groups[0]
    assert g['lr'] == 0.01 and g['betas'] == (0.9, 0.999) and g['eps'] == 1e-8 and g['maximize'] is False
    assert o.losses().numel() == 0
    o.step()                                     # no gradient anywhere: nothing to do, nothing launched (torch's Adam too)
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([p], betas=(1.0, 0.9))


def test_acquisition_mirrors_fail_loudly_without_cuda_and_keep_the_reference_surface():
    """Bayesian_optimization.acq (reference acq.py:118-294): constructor defaults and forward arities of the reference, no
    CPU fallback for the scores, and the packed-row layout of a batched sweep with the score block."""
    import inspect
    from fidelityfusion_b200 import _lib
    from fidelityfusion_b200.Bayesian_optimization import acq as A
    from fidelityfusion_b200.batched import result_layout, result_shapes, empty_result, unpack_results
    mu, var = torch.zeros(5, 1, dtype=torch.float64), torch.ones(5, 1, dtype=torch.float64)
    ucb, ei, pi, kg = A.UCB(lambda X: mu, lambda X: var), A.EI(lambda X: mu, lambda X: var), A.PI(lambda X: mu, lambda X: var), \
        A.KG(lambda X: mu, lambda X: var)
    assert (ucb.kappa, ei.xi, pi.sita, kg.num_fantasies) == (2.0, 0.01, 0.01, 10)
    assert [len(inspect.signature(f.forward).parameters) for f in (ucb, ei, pi, kg)] == [1, 2, 2, 2]     # acq.py:31-42 dispatches on this
    assert list(inspect.signature(A.optimize_acqf).parameters) == ['acq', 'raw_samples', 'bounds', 'f_best', 'num_restarts', 'options']
    assert list(inspect.signature(A.find_next_batch).parameters) == ['acq', 'bounds', 'batch_size', 'n_samples', 'f_best']
    for call in (lambda: ucb.forward(None), lambda: ei.forward(None, 0.0), lambda: pi.forward(None, 0.0)):
        with pytest.raises(_lib.FFGPError):
            call()
    # KG and PF are torch expressions of the caller's posterior (Monte-Carlo fantasies / a product of normal cdfs)
    assert tuple(kg.forward(None, 0.0).shape) == (1,)
    pf = A.PF(lambda X: torch.zeros(4, 2, dtype=torch.float64), lambda X: torch.ones(4, 2, dtype=torch.float64), [0.0, 0.0])
    assert torch.allclose(pf.forward(torch.zeros(4, 3)), torch.full((4,), 0.25, dtype=torch.float64))
    # packed row with the score block: [nll | grads | mean | var | score | info]
    keys, widths = result_layout(3, 1, 4, True, False, True)
    assert keys == ['nll', 'g_length_scales', 'g_signal_variance', 'g_log_beta', 'mean', 'var', 'score', 'info']
    assert widths == [1, 3, 1, 1, 4, 4, 4, 1]
    buf = torch.arange(2 * sum(widths), dtype=torch.float64).reshape(2, -1)
    out = unpack_results(buf, result_shapes(3, 1, 4), keys)
    assert tuple(out['score'].shape) == (2, 4) and float(out['score'][0, 0]) == 14.0 and float(out['info'][1]) == 2 * 19 - 1
    assert tuple(empty_result(3, 1, 4, True, True, 'cpu', True)['score'].shape) == (0, 4)
