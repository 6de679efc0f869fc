"""gen-2024 GP operator `cigp` (reference GaussianProcess/cigp_v10.py:17-69), same signature and
state_dict keys (`kernel.*`, `log_beta`).  negative_log_likelihood() is ONE fused CUDA call:
kernel-matrix assembly -> blocked Cholesky + triangular inverse -> NLL -> analytic gradient."""
import math

import torch
import torch.nn as nn

from .. import ops
from .kernel import fused_or_none

JITTER = 1e-6
EPS = 1e-10
PI = 3.1415   # the reference's constant, cigp_v10.py:15 (sic) - part of the parity contract


class cigp(nn.Module):
    def __init__(self, kernel, log_beta):
        super().__init__()
        self.kernel = kernel
        self.log_beta = nn.Parameter(torch.tensor([log_beta]))
        self.factor_cache = ops.FactorCache()   # L^-1 and alpha stay on the device between forward() calls
        # forward() treats hyper-parameters and training data as constants of the posterior (fast path: resident factor,
        # gradient w.r.t. x_test only - what the acquisition optimisers use).  The reference's posterior is an ordinary
        # autograd expression in EVERYTHING (cigp_v10.py:24-48); set this to True to get that: the kernel matrices are
        # materialised and the conditional runs through the differentiable covariance-input path.
        self.differentiable_posterior = False

    @staticmethod
    def _split(y_train):
        if isinstance(y_train, list):
            return y_train[0], y_train[1]
        return y_train, None

    def forward(self, x_train, y_train, x_test):
        """Posterior mean [N*,D] and full covariance [N*,N*]; e^{-log_beta} is added to EVERY covariance
        entry and y_var is ignored, exactly as cigp_v10.py:24-48."""
        y_train, _ = self._split(y_train)
        noise = self.log_beta.exp().pow(-1)
        if self.differentiable_posterior and torch.is_grad_enabled():
            from .gp_computation_pack import conditional_Gaussian
            n = x_train.size(0)
            eye = torch.eye(n, dtype=y_train.dtype, device=y_train.device)
            Sigma = self.kernel(x_train, x_train) + noise * eye + JITTER * eye
            mean, cov = conditional_Gaussian(y_train, Sigma, self.kernel(x_train, x_test), self.kernel(x_test, x_test))
            return mean, cov + noise
        fp = fused_or_none(self.kernel)
        diag = (noise + JITTER).reshape(1)
        if fp is not None:
            # differentiable w.r.t. x_test when it requires grad (acquisition optimisers, DMF_acq.py:247-254);
            # parameters and training data are constants of the posterior
            inv_ls, amp, clamp = fp
            mean, cov = ops.dense_predict(x_train, y_train, x_test, inv_ls.detach(), amp.detach(), diag_add=diag.detach(),
                                          cov_offset=noise.detach(), full_cov=True, clamp=clamp, cache=self.factor_cache,
                                          cache_token=ops.state_token(self, x_train, y_train))
        else:
            with torch.no_grad():
                mean, cov = ops.dense_predict(None, y_train, None, None, None, diag_add=diag,
                                              sigma_add=self.kernel(x_train, x_train), Ks=self.kernel(x_train, x_test),
                                              Kss=self.kernel(x_test, x_test), cov_offset=noise, full_cov=True)
        return mean, cov

    def negative_log_likelihood(self, x_train, y_train):
        """Returns the LOG-likelihood (-nll), like the reference (cigp_v10.py:69); y_train may be [y, y_var]
        in which case diag(y_var) is added to Sigma (cigp_v10.py:59-60)."""
        y_train, y_var = self._split(y_train)
        n, D = y_train.shape
        diag = self.log_beta.exp().pow(-1) + JITTER
        diag = diag.expand(n)
        if y_var is not None:
            diag = diag + y_var.diag()
        fp = fused_or_none(self.kernel)
        if fp is not None:
            inv_ls, amp, clamp = fp
            core = ops.dense_nll(x_train, y_train, inv_ls, amp, diag_add=diag, clamp=clamp)
        else:
            core = ops.dense_nll(None, y_train, None, None, diag_add=diag, sigma_add=self.kernel(x_train, x_train))
        if self.factor_cache is not None:
            self.factor_cache.invalidate()
        nll = core + 0.5 * n * D * math.log(2 * PI)
        return -nll
