"""Kernel-matrix builders with the reference's nn.Module protocol `kernel(x1, x2) -> K`
(reference: GaussianProcess/kernel.py).  Each stationary kernel also exposes `fused_params()`
so that the GP operators can fuse assembly + Cholesky + gradient into one C call instead of
materialising K through autograd."""
import torch
import torch.nn as nn

from .. import ops

EPS = 1e-9


class ARDKernel(nn.Module):
    """reference GaussianProcess/kernel.py:65-105: l = |length_scales| + eps,
    K = |signal_variance| * exp(-0.5 * cdist(x1/l, x2/l)^2)."""

    def __init__(self, input_dim, initial_length_scale=1.0, initial_signal_variance=1.0, eps=EPS):
        super().__init__()
        self.length_scales = nn.Parameter(torch.ones(input_dim) * initial_length_scale)
        self.signal_variance = nn.Parameter(torch.tensor([initial_signal_variance]))
        self.eps = eps

    def fused_params(self):
        """(inv_ls[d], amp[1], clamp) as differentiable functions of the parameters."""
        return 1.0 / (torch.abs(self.length_scales) + self.eps), self.signal_variance.abs(), True

    def forward(self, x1, x2):
        inv_ls, amp, clamp = self.fused_params()
        return ops.kernel_matrix(x1, x2, inv_ls, amp, clamp)


class SquaredExponentialKernel(nn.Module):
    """reference GaussianProcess/kernel.py:239-272: scalar log-parameters,
    K = exp(signal_variance)^2 * exp(-0.5 * d2 / exp(length_scale)^2), explicit norm expansion, no clamp."""

    def __init__(self, length_scale=1.0, signal_variance=1.0):
        super().__init__()
        self.length_scale = nn.Parameter(torch.tensor([length_scale]))
        self.signal_variance = nn.Parameter(torch.tensor([signal_variance]))

    def fused_params(self):
        return torch.exp(-self.length_scale), self.signal_variance.exp().pow(2), False

    def forward(self, x1, x2):
        inv_ls, amp, clamp = self.fused_params()
        return ops.kernel_matrix(x1, x2, inv_ls, amp, clamp)


class SumKernel(nn.Module):
    """reference GaussianProcess/kernel.py:172-203 (composition happens on the materialised matrices)."""

    def __init__(self, kernel1, kernel2):
        super().__init__()
        self.kernel1 = kernel1
        self.kernel2 = kernel2

    def forward(self, x1, x2):
        return self.kernel1(x1, x2) + self.kernel2(x1, x2)


class ProductKernel(nn.Module):
    """reference GaussianProcess/kernel.py:205-236."""

    def __init__(self, kernel1, kernel2):
        super().__init__()
        self.kernel1 = kernel1
        self.kernel2 = kernel2

    def forward(self, x1, x2):
        return self.kernel1(x1, x2) * self.kernel2(x1, x2)


def fused_or_none(kernel):
    """fused_params() of a stationary SE-family kernel, else None (the caller then materialises K)."""
    f = getattr(kernel, 'fused_params', None)
    return f() if f is not None else None
