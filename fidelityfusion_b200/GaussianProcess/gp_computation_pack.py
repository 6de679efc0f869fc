"""Functional GP pack (reference GaussianProcess/gp_computation_pack.py): Gaussian_log_likelihood,
conditional_Gaussian, negative_log_likelihood and the Tensor_linear coupling, on libffgp kernels.
The covariance is given as a matrix here, so these run the CUDA path in covariance-input mode and
hand dNLL/dSigma back to autograd."""
import math

import numpy as np
import torch

from .. import ops
from ..tensorly_compat import mode_dot
from .kernel import fused_or_none

EPS = 1e-9
JITTER = 1e-6
PI = 3.1415


def _mm(A, B):
    """A @ B for 2-D operands on libffgp's DMMA mode-product kernels (autograd-aware): (A @ B) = B x_0 A."""
    return mode_dot(B, A, 0)


class _SolveLogdet(torch.autograd.Function):
    """One CUDA factorisation of a given covariance -> (alpha = Sigma^-1 Y, log|Sigma|), differentiable in Y and Sigma.
    Every `Kinv_method` variant of the reference is an expression in these two (gp_computation_pack.py:55-88)."""

    @staticmethod
    def forward(ctx, y, cov):
        from .. import _lib as B
        L = B.lib()
        n, D = y.shape
        dev = y.device
        want = any(ctx.needs_input_grad)
        yc, cc = ops._f64c(y).unsqueeze(0), ops._f64c(cov).unsqueeze(0)
        core = torch.empty(1, dtype=torch.float64, device=dev)
        logdet = torch.empty(1, dtype=torch.float64, device=dev)
        alpha = torch.empty(1, n, D, dtype=torch.float64, device=dev)
        G = torch.empty(1, n, n, dtype=torch.float64, device=dev) if want else None
        gdiag = torch.empty(1, n, dtype=torch.float64, device=dev) if want else None
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        wsb = L.ffgp_dense_workspace_bytes(n, 0, D, 0, 1)
        ws = ops._ws_cache.get(wsb, dev)
        rc = L.ffgp_dense_nll_f64(None, B.ptr(yc), None, None, None, B.ptr(cc), n, 0, D, 1, 0, 0, int(want), B.ptr(ws),
                                  wsb, B.ptr(core), B.ptr(logdet), B.ptr(alpha), None, None, B.ptr(gdiag), B.ptr(G),
                                  B.ptr(info), B.stream_ptr())
        B.check(rc, 'ffgp_dense_nll_f64')
        ops.check_info(info)
        a = alpha[0]
        if want:
            # the kernel returns G = 0.5 (D S - a a^T), S = Sigma^-1 (symmetric): recover S with our own product
            S = (2.0 * G[0] + _mm(a, a.T.contiguous())) / D
            ctx.save_for_backward(a, S)
        ctx.dt = (y.dtype, cov.dtype)
        return a.to(y.dtype), logdet[0].to(y.dtype)

    @staticmethod
    def backward(ctx, g_alpha, g_logdet):
        a, S = ctx.saved_tensors
        ydt, cdt = ctx.dt
        b = _mm(S, ops._f64c(g_alpha))                    # Sigma^-1 g_alpha
        gy = b.to(ydt) if ctx.needs_input_grad[0] else None
        gS = None
        if ctx.needs_input_grad[1]:
            gS = (-_mm(b, a.T.contiguous()) + g_logdet.to(torch.float64) * S).to(cdt)
        return gy, gS


def Gaussian_log_likelihood(y, cov, Kinv_method='cholesky3'):
    """reference gp_computation_pack.py:34-91.  All variants run the same CUDA factorisation; what differs is the
    expression the reference forms from it, reproduced as written: 'cholesky1' / 'direct' use y^T Sigma^-1 y (a D x D
    matrix for D columns), 'cholesky2' / 'cholesky3' put cholesky_solve(y, L) = Sigma^-1 y inside the square
    (SURVEY A-12), the torch_distribution variants evaluate a MultivariateNormal centred on y at y."""
    assert len(y.shape) == 2 and len(cov.shape) == 2, "y, mean, cov should be 2D tensors"
    n, D = y.shape
    if Kinv_method in ('torch_distribution_MN1', 'torch_distribution_MN2'):
        # MultivariateNormal(loc=y, ...).log_prob(y): the event size is y.shape[1] and must equal len(cov) (:85-88);
        # the deviation is zero, so every batch row gets -0.5 (n log 2 pi + log|cov|)
        if D != cov.shape[0]:
            raise ValueError(f'MultivariateNormal: loc has event size {D} but the covariance is {cov.shape[0]} x {cov.shape[1]}')
        _, logdet = _SolveLogdet.apply(y[:, :1], cov)
        return (-0.5 * (D * np.log(2 * np.pi) + logdet)).expand(n)
    alpha, logdet = _SolveLogdet.apply(y, cov)
    if Kinv_method in ('cholesky1', 'direct'):
        return -0.5 * (_mm(y.T.contiguous(), alpha) + 2 * logdet + n * np.log(2 * np.pi))
    if Kinv_method == 'cholesky2':
        return -0.5 * (_mm(alpha.T.contiguous(), alpha) + 2 * logdet + n * np.log(2 * np.pi))
    if Kinv_method == 'cholesky3':
        if D > 1:
            return -0.5 * ((alpha ** 2).sum() + logdet * D + n * D * np.log(2 * np.pi))
        return -0.5 * ((alpha ** 2).sum() + logdet + n * np.log(2 * np.pi)).reshape(1, 1)
    raise ValueError('Kinv_method should be either direct or cholesky')


def conditional_Gaussian(y, Sigma, K_s, K_ss, Kinv_method='cholesky3'):
    """reference gp_computation_pack.py:93-118: mu = K_s^T Sigma^-1 y, cov = K_ss - (L^-1 K_s)^T (L^-1 K_s).
    Without gradient tracking: one fused prediction call.  With it (any input requires grad under grad mode - the
    reference's posterior is an ordinary autograd expression): ONE factorisation with the right-hand sides [y, K_s],
    mu and cov formed from Sigma^-1 [y, K_s] with our own products, differentiable in all four inputs."""
    if Kinv_method not in ('cholesky1', 'cholesky3', 'direct'):
        raise ValueError('Kinv_method should be either direct or cholesky')
    if torch.is_grad_enabled() and any(t.requires_grad for t in (y, Sigma, K_s, K_ss)):
        D = y.shape[1]
        sol, _ = _SolveLogdet.apply(torch.cat([y, K_s], 1), Sigma)
        KsT = K_s.T.contiguous()
        return _mm(KsT, sol[:, :D].contiguous()), K_ss - _mm(KsT, sol[:, D:].contiguous())
    with torch.no_grad():
        mu, cov = ops.dense_predict(None, y, None, None, None, sigma_add=Sigma, Ks=K_s, Kss=K_ss, full_cov=True)
    return mu, cov


def negative_log_likelihood(kernel, log_beta, x_train, y_train):
    """reference gp_computation_pack.py:120-136: jitter relative to mean(K)."""
    n, D = y_train.shape
    fp = fused_or_none(kernel)
    kmean = kernel(x_train, x_train).mean()
    diag = (log_beta.exp().pow(-1) + JITTER * kmean).expand(n)
    if fp is not None:
        inv_ls, amp, clamp = fp
        core = ops.dense_nll(x_train, y_train, inv_ls, amp, diag_add=diag, clamp=clamp)
    else:
        core = ops.dense_nll(None, y_train, None, None, diag_add=diag, sigma_add=kernel(x_train, x_train))
    return -(core + 0.5 * n * D * math.log(2 * PI))


class Tensor_linear(torch.nn.Module):
    """reference gp_computation_pack.py:138-158, including the quirk that forward() applies only the LAST
    mode's matrix (each loop iteration restarts from x)."""

    def __init__(self, l_shape, h_shape):
        super().__init__()
        self.l_shape = l_shape
        self.h_shape = h_shape
        vectors = []
        for i in range(len(self.l_shape)):
            if self.l_shape[i] < self.h_shape[i]:
                init_tensor = torch.eye(self.l_shape[i])
                init_tensor = torch.nn.functional.interpolate(init_tensor.reshape(1, 1, *init_tensor.shape),
                                                              (self.l_shape[i], self.h_shape[i]), mode='bilinear')
                init_tensor = init_tensor.squeeze().T
            elif self.l_shape[i] == self.h_shape[i]:
                init_tensor = torch.eye(self.l_shape[i])
            vectors.append(torch.nn.Parameter(init_tensor))
        self.vectors = torch.nn.ParameterList(vectors)

    def forward(self, x):
        i = len(self.l_shape) - 1
        return mode_dot(x, self.vectors[i], i + 1)
