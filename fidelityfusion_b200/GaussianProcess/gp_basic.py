"""GP_basic (reference GaussianProcess/gp_basic.py:15-153): K + noise_variance^2 I (+ full y_var), no jitter."""
import torch
import torch.nn as nn

import numpy as np

from .gp_computation_pack import Gaussian_log_likelihood, conditional_Gaussian, _SolveLogdet, _mm


class GP_basic(nn.Module):
    def __init__(self, kernel, noise_variance):
        super().__init__()
        self.kernel = kernel
        self.noise_variance = nn.Parameter(torch.tensor([noise_variance]))

    def _cov(self, x_train, y_train):
        if isinstance(y_train, list):
            y_train_var = y_train[1]
            y_train = y_train[0]
        else:
            y_train_var = None
        n = len(x_train)
        K = self.kernel(x_train, x_train) + self.noise_variance.pow(2) * torch.eye(n, dtype=x_train.dtype, device=x_train.device)
        if y_train_var is not None:
            K = K + y_train_var
        return K, y_train

    def forward(self, x_train, y_train, x_test, Kinv_method='cholesky3'):
        K, y_train = self._cov(x_train, y_train)
        mu, var = conditional_Gaussian(y_train, K, self.kernel(x_train, x_test), self.kernel(x_test, x_test), Kinv_method)
        return mu.squeeze(), var

    def log_likelihood(self, x_train, y_train, Kinv_method='cholesky3'):
        K, y_train = self._cov(x_train, y_train)
        if Kinv_method == 'cholesky2':
            # gp_basic.py:122-125 sums the D x D matrix gamma^T gamma (the pack function returns the matrix, :62-65)
            n = len(x_train)
            alpha, logdet = _SolveLogdet.apply(y_train, K)
            return -0.5 * (_mm(alpha.T.contiguous(), alpha).sum() + 2 * logdet + n * np.log(2 * np.pi))
        return Gaussian_log_likelihood(y_train, K, Kinv_method)
