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


# ---------------------------------------------------------------------------------------------------------------------
# The remaining kernels of the reference file (LinearKernel :23-63, MaternKernel :109-169, RationalQuadraticKernel
# :275-310, MaternKernel_scalarLengthScale :312-347).  They are not on the path north_star names (ARD / RBF builders) and
# have no fused assembly kernel: the cross term x1 x2^T runs on libffgp's DMMA product (ops.matmul), the elementwise
# map is torch on the device, and the GP operators take the materialised matrix through the covariance-input mode of
# the C ABI (which differentiates a given Sigma).  Same constructor signatures, parameter names and formulas, so
# `import GaussianProcess.kernel as kernel; kernel.MaternKernel(...)` keeps working under the alias.
# ---------------------------------------------------------------------------------------------------------------------
def _sqdist(x1, x2):
    """||x1_i - x2_j||^2 by the norm expansion, cross term on the DMMA product."""
    return (x1 ** 2).sum(1).reshape(-1, 1) + (x2 ** 2).sum(1) - 2 * ops.matmul(x1, x2.T.contiguous())


class LinearKernel(nn.Module):
    def __init__(self, input_dim, initial_length_scale=1.0, initial_signal_variance=1.0):
        super().__init__()
        self.length_scales = nn.Parameter(torch.ones(input_dim) * initial_length_scale)
        self.signal_variance = nn.Parameter(torch.tensor([initial_signal_variance]))
        self.center = nn.Parameter(torch.zeros(input_dim))

    def forward(self, x1, x2):
        x1 = (x1 - self.center) / self.length_scales
        x2 = (x2 - self.center) / self.length_scales
        return ops.matmul(x1, x2.T.contiguous()) * self.signal_variance.abs()


class MaternKernel(nn.Module):
    def __init__(self, input_dim, initial_length_scale=1.0, initial_signal_variance=1.0, nu=2.5, rho=1, eps=EPS):
        super().__init__()
        self.length_scales = nn.Parameter(torch.ones(input_dim) * initial_length_scale)
        self.signal_variance = nn.Parameter(torch.tensor([initial_signal_variance]))
        self.eps = eps
        self.nu = nu
        self.rho = rho

    def forward(self, x1, x2):
        ls = torch.abs(self.length_scales) + self.eps
        # torch.cdist(...)**2 of the reference clamps the expansion at 0 before its sqrt (kernel.py:151)
        sqdist = _sqdist(x1 / ls, x2 / ls).clamp_min(0.0)
        amp = self.signal_variance.abs()
        if self.nu == 0.5:
            return amp * torch.exp(-torch.sqrt(sqdist) / self.rho)
        if self.nu == 1.5:
            r = torch.sqrt(3 * sqdist) / self.rho
            return amp * (1 + r) * torch.exp(-r)
        if self.nu == 2.5:
            r = torch.sqrt(5 * sqdist) / self.rho
            return amp * (1 + r + 5 / 3 * sqdist / self.rho ** 2) * torch.exp(-r)
        return None                                     # the reference falls off the end for other nu (:160-169)


class RationalQuadraticKernel(nn.Module):
    def __init__(self, length_scale=1., signal_variance=1., alpha=1.):
        super().__init__()
        self.length_scale = nn.Parameter(torch.tensor([length_scale]))
        self.signal_variance = nn.Parameter(torch.tensor([signal_variance]))
        self.alpha = nn.Parameter(torch.tensor([alpha]))

    def forward(self, x1, x2):
        return self.signal_variance.pow(2) * torch.pow(1 + 0.5 * _sqdist(x1, x2) / self.alpha / self.length_scale.pow(2), -self.alpha)


class MaternKernel_scalarLengthScale(nn.Module):
    def __init__(self, length_scale=1.0, signal_variance=1.0, nu=2.5):
        super().__init__()
        self.length_scale = nn.Parameter(torch.tensor([length_scale]))
        self.signal_variance = nn.Parameter(torch.tensor([signal_variance]))
        self.nu = nn.Parameter(torch.tensor([nu]))

    def forward(self, x1, x2):
        # as written (:345-346): (1 + sqrt(3 d^2) / l^2)^(-nu) - not a Matern form, reproduced
        return self.signal_variance.pow(2) * torch.pow(1 + torch.sqrt(3 * _sqdist(x1, x2)) / self.length_scale.pow(2), -self.nu)
