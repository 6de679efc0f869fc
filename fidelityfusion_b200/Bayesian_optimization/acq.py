"""Single-fidelity acquisition functions on the device (reference Bayesian_optimization/acq.py:118-294).

Same class names, constructor arguments and `forward` signatures as the reference; `mean_func` / `variance_func` are the
caller's posterior (e.g. `lambda X: model.forward(x, y, X)[0]` of a cigp drop-in) and must return CUDA tensors.  Scores
and their partial derivatives come from the fused kernel behind `ffgp_acquisition_f64` (kinds 1, 3, 4) instead of a host
round trip through scipy.stats.norm (acq.py:178, 230); candidate sets are drawn and ranked on the device.

Differences a user can observe (all keep the numbers): results stay on the device the posterior lives on (the reference
builds CPU float32 tensors for the cdf / pdf, so its EI / PI only work for CPU posteriors); `PF.forward` returns a
tensor, not a numpy array (`.cpu().numpy()` gives the reference's array); `optimize_acqf` runs - the reference's raises
"can't optimize a non-leaf Tensor" (acq.py:44-47 builds the start points as an expression of a Parameter).
"""
import inspect
import math

import torch
import torch.nn as nn

from ..MF_BayesianOptimization.Discrete.DMF_acq import acquisition


def _forward_arity(acq):
    return len(inspect.signature(acq.forward).parameters)


def optimize_acqf(acq, raw_samples, bounds, f_best=0, num_restarts=30, options=None):
    """reference acq.py:10-69: `raw_samples` start points uniform in `bounds` [d, 2], Adam (lr 0.1) on the negative summed
    score for `num_restarts` steps, the iterate with the smallest loss is returned.  The chain d score / d X goes
    through the kernel's partials and whatever `mean_func` / `variance_func` differentiate (the fused posterior
    gradient for the GP drop-ins)."""
    bounds = torch.as_tensor(bounds)
    nargs = _forward_arity(acq)

    def obj_func(X):
        return -(acq.forward(X) if nargs == 1 else acq.forward(X, f_best)).sum()

    lo, hi = bounds[:, 0], bounds[:, 1]
    X = nn.Parameter(torch.rand((raw_samples, len(bounds)), dtype=bounds.dtype if bounds.is_floating_point() else torch.float32,
                                device=bounds.device) * (hi - lo) + lo)
    optimizer = torch.optim.Adam([X], lr=0.1)
    best_x = X.clone().detach()
    with torch.no_grad():
        best_value = float(obj_func(best_x))
    for _ in range(num_restarts):
        optimizer.zero_grad()
        loss = obj_func(X)
        if loss.requires_grad:
            loss.backward()
            optimizer.step()
        if loss.item() < best_value:
            best_value = loss.item()
            best_x = X.clone().detach()
    return best_x


def find_next_batch(acq, bounds, batch_size=1, n_samples=1000, f_best=0):
    """reference acq.py:80-115: per batch slot, `n_samples` uniform points in [bounds[0, 0], bounds[0, 1]] (the first
    dimension's bounds for every dimension, as written, :105), the one with the highest score is taken.  Points are
    drawn on the device of `bounds`; the scores never leave it."""
    bounds = torch.as_tensor(bounds)
    nargs = _forward_arity(acq)
    picked = []
    for _ in range(batch_size):
        X = torch.empty(n_samples, bounds.shape[0], dtype=torch.float32, device=bounds.device).uniform_(
            float(bounds[0, 0]), float(bounds[0, 1]))
        with torch.no_grad():
            values = acq.forward(X) if nargs == 1 else acq.forward(X, f_best)
        picked.append(X[torch.argmax(values)])
    return torch.stack(picked)


class _MeanVar:
    def __init__(self, mean_func, variance_func):
        self.mean_func = mean_func
        self.variance_func = variance_func

    def _posterior(self, X):
        return self.mean_func(X), self.variance_func(X)


class UCB(_MeanVar):
    """mean + kappa * sqrt(variance)  (acq.py:118-149)"""

    def __init__(self, mean_func, variance_func, kappa=2.0):
        super().__init__(mean_func, variance_func)
        self.kappa = kappa

    def forward(self, X):
        mean, variance = self._posterior(X)
        return acquisition(mean, variance, 'UCB_STD', beta=self.kappa)


class EI(_MeanVar):
    """(mean - f_best - xi) Phi(Z) + std phi(Z), std clamped at 1e-9, Phi / phi float32 constants (acq.py:152-181)"""

    def __init__(self, mean_func, variance_func, xi=0.01):
        super().__init__(mean_func, variance_func)
        self.xi = xi

    def forward(self, X, f_best):
        mean, variance = self._posterior(X)
        return acquisition(mean, variance, 'EI', f_best=float(f_best), xi=self.xi)


class PI(_MeanVar):
    """Phi(Z) as a float32 tensor without gradient (acq.py:184-231; the reference builds it from a numpy array)"""

    def __init__(self, mean_func, variance_func, sita=0.01):
        super().__init__(mean_func, variance_func)
        self.sita = sita

    def forward(self, X, f_best):
        mean, variance = self._posterior(X)
        with torch.no_grad():
            return acquisition(mean, variance, 'PI_CDF', f_best=float(f_best), xi=self.sita).to(torch.float32)


class KG(_MeanVar):
    """Monte-Carlo knowledge gradient (acq.py:233-262): random by construction, sampled with the device generator."""

    def __init__(self, mean_func, variance_func, num_fantasies=10):
        super().__init__(mean_func, variance_func)
        self.num_fantasies = num_fantasies

    def forward(self, X, f_best):
        mean, variance = self._posterior(X)
        std = torch.nan_to_num(torch.clamp(torch.sqrt(variance), min=1e-6), nan=1e-6)
        fantasies = torch.distributions.Normal(mean, std.expand_as(mean)).rsample(torch.Size([self.num_fantasies]))
        best, _ = fantasies.max(dim=0)
        return (best - f_best).mean(dim=0)


class PF(_MeanVar):
    """Probability of feasibility prod_i Phi((threshold_i - mu_i) / sigma_i) over the constraint outputs (acq.py:264-294)."""

    def __init__(self, mean_func, variance_func, thresholds):
        super().__init__(mean_func, variance_func)
        self.thresholds = thresholds

    def forward(self, X):
        mu, variance = self._posterior(X)
        sigma = torch.sqrt(variance)
        pf = torch.ones(X.shape[0], dtype=torch.float64, device=mu.device)
        for i, thr in enumerate(self.thresholds):
            z = ((thr - mu[:, i]) / sigma[:, i]).to(torch.float64)
            pf = pf * (0.5 * torch.erfc(-z / math.sqrt(2.0)))
        return pf
