"""Discrete multi-fidelity acquisition functions on the device (reference
MF_BayesianOptimization/Discrete/DMF_acq.py:15-262).  Same class / method names and argument meaning;
the score and its partial derivatives come from ONE CUDA kernel (ffgp_acquisition_f64) instead of a host round trip
through scipy.stats.norm (DMF_acq.py:104), and the chain to the candidate x goes through the fused posterior
gradient (ops.dense_predict -> ffgp_dense_predict_bwd_f64)."""
import torch
import torch.nn as nn

from ... import _lib as B
from ... import ops

PI = 3.1415926
from ...batched import ACQ_KINDS as KINDS           # {'UCB': 0, 'EI': 1, 'PI': 2, 'UCB_STD': 3, 'PI_CDF': 4}


class _Acq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, var, kind, f_best, beta, xi, round_f32):
        L = B.lib()
        shape, dt = mean.shape, mean.dtype
        mc = mean.detach().to(torch.float64).contiguous().reshape(-1)
        vc = var.detach().to(torch.float64).expand(shape).contiguous().reshape(-1)
        m = mc.numel()
        score, dm, dv = torch.empty_like(mc), torch.empty_like(mc), torch.empty_like(mc)
        rc = L.ffgp_acquisition_f64(B.ptr(mc), B.ptr(vc), m, int(kind), float(f_best), float(beta), float(xi),
                                    int(bool(round_f32)), B.ptr(score), B.ptr(dm), B.ptr(dv), B.stream_ptr())
        B.check(rc, 'ffgp_acquisition_f64')
        ctx.save_for_backward(dm, dv)
        ctx.meta = (shape, tuple(var.shape), dt, var.dtype)
        return score.reshape(shape).to(dt)

    @staticmethod
    def backward(ctx, g):
        dm, dv = ctx.saved_tensors
        shape, vshape, dt, vdt = ctx.meta
        g = g.to(torch.float64).reshape(-1)
        gm = (g * dm).reshape(shape).to(dt)
        gv = (g * dv).reshape(shape)
        if vshape != tuple(shape):
            gv = gv.sum_to_size(vshape)
        return gm, gv.to(vdt), None, None, None, None, None


def acquisition(mean, var, kind, f_best=0.0, beta=1.0, xi=0.01, round_f32=True):
    """score(mean, var) of kind 'UCB' | 'EI' | 'PI', differentiable in mean and var.  `var` broadcasts to `mean`."""
    return _Acq.apply(mean, var, KINDS[kind] if isinstance(kind, str) else kind, f_best, beta, xi, round_f32)


class DiscreteAcquisitionFunction(nn.Module):
    """reference DMF_acq.py:15-166: mean_function(x, s) / variance_function(x, s) are the model's posterior at
    fidelity s; every *_MF method returns the score tensor of the same shape as the mean."""

    def __init__(self, mean_function, variance_function, fidelity_num, x_dimension, f_best):
        super().__init__()
        self.mean_function = mean_function
        self.variance_function = variance_function
        self.fidelity_num = fidelity_num
        self.x_dimension = x_dimension
        self.f_best = f_best if f_best is not None else None
        self.beta = 0.2 * int(x_dimension)

    def _fb(self):
        fb = self.f_best
        return float(fb.item() if isinstance(fb, torch.Tensor) else fb)

    def UCB_MF(self, x, s):
        self.beta = 0.2 * int(self.x_dimension)
        return acquisition(self.mean_function(x, s), self.variance_function(x, s), 'UCB', beta=self.beta)

    def EI_MF(self, x, s):
        self.beta = 0.2 * int(self.x_dimension)
        return acquisition(self.mean_function(x, s), self.variance_function(x, s), 'EI', f_best=self._fb(), xi=0.01)

    def PI_MF(self, x, s):
        self.beta = 0.2 * int(self.x_dimension)
        return acquisition(self.mean_function(x, s), self.variance_function(x, s), 'PI', f_best=self._fb(), xi=0.01)

    def KG_MF(self, x, s):
        """Monte-Carlo knowledge gradient with 10 fantasies (DMF_acq.py:130-165); random by construction."""
        self.beta = 0.2 * int(self.x_dimension)
        mean = self.mean_function(x, s)
        std = torch.nan_to_num(torch.clamp(torch.sqrt(self.variance_function(x, s)), min=1e-6), nan=1e-6)
        fantasies = torch.distributions.Normal(mean, std.expand_as(mean)).rsample(sample_shape=torch.Size([10]))
        best, _ = fantasies.max(dim=0)
        return (best - self.f_best).mean(dim=0)

    def acq_selection_fidelity(self, gamma, new_x):
        new_s = 0
        for i in range(self.fidelity_num):
            v = self.variance_function(new_x, i)
            new_s = i + 1 if bool((self.beta * v > gamma[i]).all()) else i
        return new_s


def optimize_acq_mf(fidelity_manager, acq_mf, n_iterations=10, learning_rate=0.001):
    """reference DMF_acq.py:226-262, same signature: Adam on a candidate x per fidelity, the fidelity whose final
    negative score is smallest wins.  `fidelity_num` and `x_dimension` are read from the data manager as the reference
    does (:240-241); the candidate is a [d, 1] column (:246, as written) created on the data's device; the gradient is
    never zeroed between steps (optimizer.zero_grad() is commented out, :247-249) - kept.  `acq_mf(x, s)` is a
    DiscreteAcquisitionFunction method; the chain d score / d x runs through ffgp_acquisition_f64's partials and the
    fused posterior gradient."""
    fidelity_num = int((len(fidelity_manager.data_dict) + 1) / 2)
    X0 = fidelity_manager.data_dict['0']['X']
    x_dimension = X0.shape[1]
    scores, xs = [], []
    for i in range(fidelity_num):
        X_initial = nn.Parameter(torch.rand(x_dimension, device=X0.device, dtype=X0.dtype).reshape(-1, 1),
                                 requires_grad=True)
        optimizer = torch.optim.Adam([X_initial], lr=learning_rate)
        for j in range(n_iterations):
            loss = -1 * acq_mf(X_initial, i)
            loss.backward()
            optimizer.step()
            print('iter', j, 'x:', X_initial, 'Negative Acquisition Function:', loss.item(), end='\n')
        xs.append(X_initial.detach())
        scores.append(loss.item())
    return xs[scores.index(min(scores))]


def optimize_acq_candidates(acq, fidelity_num, x_dimension, n_iterations=10, learning_rate=0.001, x_init=None,
                            device='cuda', dtype=torch.float64):
    """The same optimisation with explicit sizes and row candidates x [1, d] (no data manager); `x_init[s]` fixes the
    start point of fidelity s (tests)."""
    best_x, best_loss = None, None
    for s in range(fidelity_num):
        x0 = x_init[s] if x_init is not None else torch.rand(1, x_dimension, device=device, dtype=dtype)
        X = nn.Parameter(x0.clone().to(device=device, dtype=dtype))
        opt = torch.optim.Adam([X], lr=learning_rate)
        loss = None
        for _ in range(n_iterations):
            loss = (-1 * acq(X, s)).sum()
            loss.backward()
            opt.step()
        lv = float(loss.item())
        if best_loss is None or lv < best_loss:
            best_loss, best_x = lv, X.detach().clone()
    return best_x


def batched_candidate_scores(x, y, length_scales, signal_variance, log_beta, xs, kind='EI', f_best=0.0, beta=1.0, xi=0.01):
    """BASELINE config 5 as its real consumer uses it: B independent candidate GPs (v1/CFKG.py:124-129 re-fits one GP
    per candidate), each scored at its own test points, nothing leaves the device.  Returns scores [B, N*]."""
    from ...batched import batched_cigp_eval
    # the score is written by the epilogue of the sweep (ffgp_batched_pack_acq_f64): no launch of its own
    out = batched_cigp_eval(x, y, length_scales, signal_variance, log_beta, xs, want_grad=False,
                            acq=dict(kind=kind, f_best=f_best, beta=beta, xi=xi))
    return out['score']
