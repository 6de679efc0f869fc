"""CPU oracle for the FidelityFusion GP hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
(fidelityfusion_b200/) never imports anything under oracle/ and has no CPU fallback.

It restates, function by function, what the reference computes on the hot path
 with the same torch CPU/LAPACK primitives the reference
itself calls (the reference IS PyTorch, so torch fp64 on CPU is the natural
restatement; gradients come from autograd exactly like `loss.backward()` in the
reference's training loops).  Every function cites the reference file:line it
follows.  Parity pinning: tests/test_oracle_golden.py checks each function against
tests/golden/*.npz, which oracle/gen_golden.py produced by importing the real
reference from /root/reference (under the two stubs of oracle/_ref_stubs.py), and
against the known-answer tests of SURVEY.md appendix B.  tensorly (un-vendored,
unpinned third-party) is restated from the published n-mode product definition.

Everything is functional: parameters are plain tensors, so the same functions serve
the autograd oracle, the CPU timing baseline and the batched configs.
"""
from __future__ import annotations

import math

import numpy as np
import torch

JITTER = 1e-6          # cigp_v10.py:13, cigp.py:4, gp_computation_pack.py:16
PI_REF = 3.1415        # cigp_v10.py:15, cigp.py:6, gp_computation_pack.py:17 (sic)


# ----------------------------------------------------------------------------
# kernel-matrix builders
# ----------------------------------------------------------------------------
def ard_kernel(x1, x2, length_scales, signal_variance, eps=1e-9):
    """GaussianProcess/kernel.py:88-105 (ARDKernel.forward)."""
    ls = length_scales.abs() + eps
    sq = torch.cdist(x1 / ls, x2 / ls, p=2) ** 2
    return signal_variance.abs() * torch.exp(-0.5 * sq)


def sqexp_kernel(x1, x2, length_scale, signal_variance):
    """GaussianProcess/kernel.py:258-272 (SquaredExponentialKernel.forward): scalar
    log-parameters, explicit norm expansion, no clamp."""
    sq = (x1 ** 2).sum(1).reshape(-1, 1) + (x2 ** 2).sum(1) - 2 * x1 @ x2.T
    return signal_variance.exp().pow(2) * torch.exp(-0.5 * sq / length_scale.exp().pow(2))


def se_kernel(X, X2, length_scale, scale, exp_format):
    """MFGP_ver2023May/kernel/SE_kernel.py:20-44.  `exp_format` is the result of the
    reference's `noise_exp_format is True` test (False when a config dict is passed,
    as kernel_utils.py:12,24 does)."""
    if exp_format:
        ls = torch.exp(length_scale).view(1, -1)
        sc = torch.exp(scale).view(1, -1)
    else:
        ls = length_scale.view(1, -1)
        sc = scale.view(1, -1)
    if X.ndim > 2:
        X = X.reshape(X.size(0), -1)
        X2 = X2.reshape(X2.size(0), -1)
    X = X / ls
    X2 = X2 / ls
    n1 = (X * X).sum(1).view(-1, 1)
    n2 = (X2 * X2).sum(1).view(-1, 1)
    K = -2.0 * X @ X2.t() + n1 + n2.t()
    return sc * torch.exp(-0.5 * K)


# ----------------------------------------------------------------------------
# dense GP: gen-2024 cigp (GaussianProcess/cigp_v10.py)
# ----------------------------------------------------------------------------
def cigp_log_likelihood(K, log_beta, y, y_var=None):
    """cigp_v10.py:50-69.  Returns what `cigp.negative_log_likelihood` returns, i.e.
    the LOG-LIKELIHOOD (-nll) as a 1-element tensor."""
    n, D = y.shape
    eye = torch.eye(n, dtype=K.dtype)
    Sigma = K + log_beta.exp().pow(-1) * eye + JITTER * eye
    if y_var is not None:
        Sigma = Sigma + y_var.diag() * eye
    L = torch.linalg.cholesky(Sigma)
    Gamma = torch.linalg.solve_triangular(L, y, upper=False)
    nll = 0.5 * (Gamma ** 2).sum() + L.diag().log().sum() * D \
        + 0.5 * n * torch.log(2 * torch.tensor(PI_REF, dtype=K.dtype)) * D
    return -nll


def cigp_predict(K, Kx, Kxx, log_beta, y):
    """cigp_v10.py:24-48: posterior mean and FULL covariance; the noise e^{-log_beta}
    is added to every entry of the covariance (sic, :44)."""
    n = K.shape[0]
    eye = torch.eye(n, dtype=K.dtype)
    Sigma = K + log_beta.exp().pow(-1) * eye + JITTER * eye
    L = torch.linalg.cholesky(Sigma)
    V = torch.linalg.solve_triangular(L, Kx, upper=False)
    mean = Kx.t() @ torch.cholesky_solve(y, L)
    cov = Kxx - V.t() @ V + log_beta.exp().pow(-1)
    return mean, cov


def pack_negative_log_likelihood(kernel_fn, log_beta, x, y):
    """gp_computation_pack.py:120-136: as cigp_log_likelihood but the jitter is
    JITTER * mean(K)."""
    n, D = y.shape
    K = kernel_fn(x, x)
    eye = torch.eye(n, dtype=K.dtype)
    Sigma = K + log_beta.exp().pow(-1) * eye + JITTER * K.mean() * eye
    L = torch.linalg.cholesky(Sigma)
    Gamma = torch.linalg.solve_triangular(L, y, upper=False)
    nll = 0.5 * (Gamma ** 2).sum() + L.diag().log().sum() * D \
        + 0.5 * n * torch.log(2 * torch.tensor(PI_REF, dtype=K.dtype)) * D
    return -nll


# ----------------------------------------------------------------------------
# dense GP: gen-2023 CIGP (MFGP_ver2023May/base_gp/cigp.py)
# ----------------------------------------------------------------------------
def CIGP_loss(K, noise, y, y_var=0.):
    """cigp.py:99-136: +NLL; `noise` is the precision returned by GP_noise_box.get()
    (utils/gp_noise.py:20-24); y_var is added as a FULL matrix / scalar (:127)."""
    n, D = y.shape
    eye = torch.eye(n, dtype=K.dtype)
    Sigma = K + JITTER * eye + noise.pow(-1) * eye + y_var
    L = torch.linalg.cholesky(Sigma)
    gamma = L.inverse() @ y
    return 0.5 * (gamma ** 2).sum() + L.diag().log().sum() * D \
        + 0.5 * n * torch.log(2 * torch.tensor(PI_REF, dtype=K.dtype)) * D


def CIGP_predict(K, Kx, kxx_diag, noise, y, x_var=0.):
    """cigp.py:61-97: mean and DIAGONAL variance expanded to the mean's shape."""
    n = K.shape[0]
    eye = torch.eye(n, dtype=K.dtype)
    Sigma = K + JITTER * eye + noise.pow(-1) * eye
    L = torch.linalg.cholesky(Sigma)
    V = torch.linalg.solve_triangular(L, Kx, upper=False)
    u = Kx.t() @ torch.cholesky_solve(y, L)
    var = kxx_diag.view(-1, 1) - (V ** 2).sum(0).view(-1, 1) + noise.pow(-1)
    return u, var.expand_as(u) + x_var


# ----------------------------------------------------------------------------
# functional pack + GP_basic (gp_computation_pack.py, gp_basic.py)
# ----------------------------------------------------------------------------
def gaussian_log_likelihood(y, cov, method='cholesky3'):
    """gp_computation_pack.py:34-91 (= gp_basic.py:120-153).  Reproduces the
    reference as written: 'cholesky2'/'cholesky3' put Sigma^{-1}y (cholesky_solve),
    not L^{-1}y, inside the quadratic form."""
    n = len(y)
    if method == 'cholesky1':
        L = torch.linalg.cholesky(cov)
        Li = torch.inverse(L)
        return -0.5 * (y.T @ (Li.T @ Li) @ y + 2 * torch.logdet(cov) + n * np.log(2 * np.pi))
    if method == 'cholesky2':
        L = torch.linalg.cholesky(cov)
        g = torch.cholesky_solve(y, L)
        return -0.5 * (g.T @ g + 2 * torch.logdet(cov) + n * np.log(2 * np.pi))
    if method == 'cholesky3':
        L = torch.linalg.cholesky(cov)
        g = torch.cholesky_solve(y, L, upper=False)
        if y.shape[1] > 1:
            D = y.shape[1]
            return -0.5 * ((g ** 2).sum() + 2 * L.diag().log().sum() * D + n * D * np.log(2 * np.pi))
        return -0.5 * (g.T @ g + 2 * L.diag().log().sum() + n * np.log(2 * np.pi))
    if method == 'direct':
        Ki = torch.inverse(cov)
        return -0.5 * (y.T @ Ki @ y + 2 * torch.logdet(cov) + n * np.log(2 * np.pi))
    raise ValueError('Kinv_method should be either direct or cholesky')


def conditional_gaussian(y, Sigma, K_s, K_ss, method='cholesky3'):
    """gp_computation_pack.py:93-118."""
    if method in ('cholesky1', 'cholesky3'):
        L = torch.linalg.cholesky(Sigma)
        alpha = torch.cholesky_solve(y, L)
        mu = K_s.T @ alpha
        v = L.inverse() @ K_s
        return mu, K_ss - v.T @ v
    if method == 'direct':
        Ki = torch.inverse(Sigma)
        return K_s.T @ Ki @ y, K_ss - K_s.T @ Ki @ K_s
    raise ValueError('Kinv_method should be either direct or cholesky')


def gp_basic_cov(K, noise_variance, y_var=None):
    """gp_basic.py:63-65 / 117-119: K + noise_variance^2 I (+ full y_var), no jitter."""
    S = K + noise_variance.pow(2) * torch.eye(K.shape[0], dtype=K.dtype)
    return S if y_var is None else S + y_var


# ----------------------------------------------------------------------------
# n-mode products (tensorly call sites)
# ----------------------------------------------------------------------------
def mode_dot(tensor, m, mode):
    """tensorly.tenalg.mode_dot: fold(M @ unfold(T, mode)); a vector contracts the mode."""
    if m.ndim == 1:
        return torch.tensordot(tensor, m, dims=([mode], [0]))
    return torch.movedim(torch.tensordot(m, tensor, dims=([1], [mode])), 0, mode)


def multi_mode_dot(tensor, mats, modes=None):
    """tensorly.tenalg.multi_mode_dot: sequential mode_dot over `modes` (default 0..)."""
    if modes is None:
        modes = list(range(len(mats)))
    dec = 0
    for m, mode in zip(mats, modes):
        tensor = mode_dot(tensor, m, mode - dec)
        if m.ndim == 1:
            dec += 1
    return tensor


def kron_outer(vectors):
    """tucker_to_tensor((ones[1..1], [v_k[n_k,1]])) as used at hogp.py:173-177: the
    outer product of the vectors."""
    out = vectors[0].reshape(-1)
    for v in vectors[1:]:
        out = out.unsqueeze(-1) * v.reshape(-1)
    return out


# ----------------------------------------------------------------------------
# Kronecker / Tucker GP (hogp.py, hogp_simple.py)
# ----------------------------------------------------------------------------
def hogp_loss(Ks, noise_inv, Y, y_var=0.):
    """MFGP_ver2023May/base_gp/hogp.py:140-198 (= both HOGP_simple.log_likelihood
    copies, hogp_simple.py:79-126): per-mode eigh, A = kron(lambda) + 1/beta,
    T1 = Y x_k U_k^T, b = vec((T1 A^-1/2) x_k U_k), g = (T1 A^-1) x_k U_k,
    loss = (nd/2 log 2pi + 1/2 sum log A + 1/2 b^T b)/nd.  `noise_inv` = 1/beta.
    Returns (loss, A, g)."""
    eig = [torch.linalg.eigh(K, UPLO='U') for K in Ks]
    A = kron_outer([e[0] for e in eig]) + noise_inv + y_var
    T1 = multi_mode_dot(Y, [e[1].T for e in eig])
    T3 = multi_mode_dot(T1 * A.pow(-1 / 2), [e[1] for e in eig])
    b = T3.reshape(-1)
    g = multi_mode_dot(T1 * A.pow(-1), [e[1] for e in eig])
    nd = A.numel()
    loss = -0.5 * nd * math.log(2 * math.pi) - 0.5 * torch.log(A).sum() - 0.5 * (b @ b)
    return -loss / nd, A, g


def hogp_predict(Ks, K_star, kss_diag, A, g, variant='hogp2023', K0=None):
    """Predictive mean/variance of the Kronecker GP.
    mean: hogp.py:216-217.  variance, three reference variants:
      'hogp2023'  hogp.py:219-238: x-mode factor K* K0 + (1e-6 eye)^2 (as written),
      'ffm'       two_fidelity_models/hogp_simple.py:54-75: (K* K0^-1 U0)^2,
      'gp'        GaussianProcess/hogp_simple.py:54-69: K* K0.
    """
    eig = [torch.linalg.eigh(K, UPLO='U') for K in Ks]
    mean = multi_mode_dot(g, [K_star] + list(Ks[1:]))
    diag_dims = kron_outer([K.diag() for K in Ks[1:]]).unsqueeze(0)
    dx = kss_diag
    for _ in range(len(Ks) - 1):
        dx = dx.unsqueeze(-1)
    diag_K = dx * diag_dims
    S2 = (A * A.pow(-1 / 2)).pow(2)
    if variant == 'hogp2023':
        fx = K_star @ Ks[0] + JITTER * torch.eye(K_star.shape[0], Ks[0].shape[0], dtype=A.dtype).pow(2)
    elif variant == 'ffm':
        fx = (K_star @ Ks[0].inverse() @ eig[0][1]).pow(2)
    elif variant == 'gp':
        fx = K_star @ Ks[0]
    else:
        raise ValueError(variant)
    fac = [fx] + [e[1].pow(2) for e in eig[1:]]
    return mean, diag_K + multi_mode_dot(S2, fac)


# ----------------------------------------------------------------------------
# fidelity couplings
# ----------------------------------------------------------------------------
def tensor_linear_init(l, h):
    """gp_computation_pack.py:144-151: eye(l) bilinearly interpolated to (l,h), transposed."""
    if l < h:
        t = torch.eye(l)
        t = torch.nn.functional.interpolate(t.reshape(1, 1, l, l), (l, h), mode='bilinear')
        return t.squeeze().T
    return torch.eye(l)


def tensor_linear_forward(x, vectors):
    """gp_computation_pack.py:155-158: every iteration restarts from x, so only the
    LAST mode's matrix is applied (sic)."""
    y = None
    for i, v in enumerate(vectors):
        y = mode_dot(x, v, i + 1)
    return y


def smooth_mapping_matrix(l, h):
    """multiscale_coupling/matrix.py:8-26."""
    if l < h:
        i = torch.arange(l, dtype=torch.get_default_dtype()).view(-1, 1)
        j = torch.arange(h, dtype=torch.get_default_dtype()).view(1, -1)
        t = 1.0 / ((i * (h / l) - j) ** 2 + 1)
        t = t / t.sum(0, keepdim=True)
        return t.transpose(1, 0)
    assert l == h
    return torch.eye(l)


def matrix_mapping_forward(low, high, vectors, rho):
    """matrix.py:71-76: res = high - rho * (low x_1 W_1 ... x_M W_M)."""
    for i, v in enumerate(vectors):
        low = mode_dot(low, v, i + 1)
    return high - low * rho


def matrix_mapping_backward(low, res, vectors, rho):
    """matrix.py:79-84."""
    for i, v in enumerate(vectors):
        low = mode_dot(low, v, i + 1)
    return low * rho + res


# ----------------------------------------------------------------------------
# whole-unit helpers used by tests and by bench.py's CPU baseline
# ----------------------------------------------------------------------------
def cigp_ard_nll_and_grads(x, y, length_scales, signal_variance, log_beta, y_var=None, want_y_grad=False):
    """One 'NLL+grad eval' exactly as the reference performs it in its training loops
    (CIGAR.py:100-105): loss = -cigp.negative_log_likelihood(x, y); loss.backward().
    Returns (loss float, dict of grads)."""
    ls = length_scales.detach().clone().requires_grad_(True)
    sv = signal_variance.detach().clone().requires_grad_(True)
    lb = log_beta.detach().clone().requires_grad_(True)
    yy = y.detach().clone().requires_grad_(want_y_grad)
    K = ard_kernel(x, x, ls, sv)
    loss = -cigp_log_likelihood(K, lb, yy, y_var)
    loss.backward()
    out = {'length_scales': ls.grad, 'signal_variance': sv.grad, 'log_beta': lb.grad}
    if want_y_grad:
        out['y'] = yy.grad
    return float(loss.detach()), out


def cigp_ard_predict(x, y, xs, length_scales, signal_variance, log_beta):
    """cigp.forward with an ARDKernel (cigp_v10.py:24-48)."""
    with torch.no_grad():
        K = ard_kernel(x, x, length_scales, signal_variance)
        Kx = ard_kernel(x, xs, length_scales, signal_variance)
        Kxx = ard_kernel(xs, xs, length_scales, signal_variance)
        return cigp_predict(K, Kx, Kxx, log_beta, y)


def cigp_ard_predict_dx(x, y, xs, length_scales, signal_variance, log_beta, w_mean, w_cov):
    """d( <w_mean, mean> + <w_cov, cov> ) / d xs of cigp.forward by autograd, as the acquisition optimisers obtain
    it (DMF_acq.py:247-254 differentiate UCB/EI of the posterior w.r.t. the candidate).  Returns (mean, cov, g_xs)."""
    xs = xs.detach().clone().requires_grad_(True)
    K = ard_kernel(x, x, length_scales, signal_variance)
    Kx = ard_kernel(x, xs, length_scales, signal_variance)
    Kxx = ard_kernel(xs, xs, length_scales, signal_variance)
    mean, cov = cigp_predict(K, Kx, Kxx, log_beta, y)
    s = (mean * w_mean).sum() + (cov * w_cov).sum()
    s.backward()
    return mean.detach(), cov.detach(), xs.grad


def acq_score(mean, var, kind, f_best=0.0, x_dimension=5):
    """DiscreteAcquisitionFunction.UCB_MF / EI_MF / PI_MF (MF_BayesianOptimization/Discrete/DMF_acq.py:47-128), with the
    reference's host round trip through scipy.stats.norm and its float32 rounding of cdf/pdf (DMF_acq.py:104)."""
    from scipy.stats import norm
    PI_ACQ = 3.1415926                                         # DMF_acq.py:7
    if kind == 'UCB':
        return mean + 0.2 * int(x_dimension) * var
    std = torch.clamp(torch.sqrt(var), min=1e-9)
    Z = (mean - f_best - 0.01) / std
    if kind == 'EI':
        cdf = torch.tensor(norm.cdf(Z.detach().numpy()), dtype=torch.float32)
        pdf = torch.tensor(norm.pdf(Z.detach().numpy()), dtype=torch.float32)
        return (mean - f_best - 0.01) * cdf + std * pdf
    return -torch.pow(Z, 2) * 0.5 - torch.log(torch.ones(1, 1)) - torch.log(torch.sqrt(2 * PI_ACQ * torch.ones(1, 1)))


def acq_sf_score(kind, mean, var, f_best=0.0, kappa=2.0, xi=0.01, thresholds=None):
    """Single-fidelity acquisition classes UCB / EI / PI / PF (Bayesian_optimization/acq.py:135-149, 166-181, 211-231,
    279-294), with the reference's scipy.stats.norm host round trip and float32 cdf / pdf tensors."""
    from scipy.stats import norm
    if kind == 'UCB':
        return mean + kappa * torch.sqrt(var)
    if kind == 'PF':
        sigma = torch.sqrt(var)
        pf = np.ones(mean.shape[0])
        for i in range(len(thresholds)):
            pf *= norm.cdf(((thresholds[i] - mean[:, i]) / sigma[:, i]).detach().numpy())
        return pf
    std = torch.clamp(torch.sqrt(var), min=1e-9)
    Z = (mean - f_best - xi) / std
    if kind == 'EI':
        return (mean - f_best - xi) * torch.tensor(norm.cdf(Z.detach().numpy()), dtype=torch.float32) \
            + std * torch.tensor(norm.pdf(Z.detach().numpy()), dtype=torch.float32)
    return torch.tensor(norm.cdf(Z.detach().numpy()), dtype=torch.float32)          # PI


def dense_nll_grads_analytic_numpy(x, y, ls_raw, sv_raw, log_beta, eps=1e-9, pi=PI_REF):
    """Independent numpy cross-check (no autograd): closed-form gradient of the cigp NLL
    through Sigma^{-1}.  Used by tests to show the analytic route the CUDA path takes
    agrees with the reference's autograd route."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    ls_raw = np.asarray(ls_raw, dtype=np.float64)
    ell = np.abs(ls_raw) + eps
    amp = abs(float(sv_raw))
    xs = x / ell
    d2 = np.maximum((xs ** 2).sum(1)[:, None] + (xs ** 2).sum(1)[None, :] - 2 * xs @ xs.T, 0.0)
    K = amp * np.exp(-0.5 * d2)
    n, D = y.shape
    noise = math.exp(-float(log_beta))
    Sigma = K + (noise + JITTER) * np.eye(n)
    L = np.linalg.cholesky(Sigma)
    Si = np.linalg.inv(Sigma)
    alpha = Si @ y
    nll = 0.5 * float((y * alpha).sum()) + D * float(np.log(np.diag(L)).sum()) + 0.5 * n * D * math.log(2 * pi)
    G = 0.5 * (D * Si - alpha @ alpha.T)          # dNLL/dSigma
    W = G * K
    g_ls = np.empty_like(ell)
    for k in range(x.shape[1]):
        diff2 = (x[:, k][:, None] - x[:, k][None, :]) ** 2
        g_ls[k] = (W * diff2).sum() / ell[k] ** 3 * np.sign(ls_raw[k])
    g_sv = W.sum() / amp * np.sign(float(sv_raw))
    g_lb = -noise * np.trace(G)
    return nll, {'length_scales': g_ls, 'signal_variance': g_sv, 'log_beta': g_lb, 'y': alpha}


# ---------------------------------------------------------------------------------------------
# Data-side matching
# ---------------------------------------------------------------------------------------------
def overlap_masks(x1, x2):
    """FidelityFusion_Models/MF_data.py:199-202 (and 235-238 negated): rows of x1 present in x2 and vice versa, by the
    reference's own [n1, n2, d] broadcast compare."""
    m1 = torch.all(x1.unsqueeze(1) == x2.unsqueeze(0), dim=-1).any(dim=-1)
    m2 = torch.all(x2.unsqueeze(1) == x1.unsqueeze(0), dim=-1).any(dim=-1)
    return m1, m2


def get_subset_index(data_a, data_b):
    """MFGP_ver2023May/utils/subset_tools.py:58-90 with subset_type='index', restated: for every sample shared by a and
    b (in the order torch.unique(dim=0) lists them) its index in a and its index in b."""
    na = data_a.shape[0]
    allr = torch.cat([data_a, data_b], 0)
    _, inv, cnt = allr.unique(sorted=False, return_inverse=True, return_counts=True, dim=0)
    rep = torch.arange(len(cnt))[cnt > 1]
    mask = (inv.reshape(-1, 1) - rep.reshape(1, -1)) == 0
    idx = torch.arange(mask.shape[0]).reshape(-1, 1) * mask
    return idx[:na].sum(0), idx[na:].sum(0) - na


# ---------------------------------------------------------------------------------------------
# FIDES residual kernel (a12)
# ---------------------------------------------------------------------------------------------
def kernel_res(X1, X2, length_scale, scale, length_scale_z, b, l1, h1, l2, h2, exp_format=False, seed=1024):
    """MFGP_ver2023May/kernel/MCMC_res_kernel.py:33-69: SE kernel in x times the scalar Monte-Carlo integral over the
    fidelity variable; reseeds the global RNG on every call (:47).  Parameters are the RAW nn.Parameter values."""
    if exp_format:
        length_scale, scale, length_scale_z = torch.exp(length_scale), torch.exp(scale), torch.exp(length_scale_z)
    length_scale, scale, length_scale_z = length_scale.view(1, -1), scale.view(1, -1), length_scale_z.view(1, -1)
    N = 100
    torch.manual_seed(seed)
    z1 = torch.rand(N) * (h1 - l1) + l1
    z2 = torch.rand(N) * (h2 - l2) + l2
    X1 = X1 / length_scale
    X2 = X2 / length_scale
    n1 = torch.sum(X1 * X1, dim=1).view(-1, 1)
    n2 = torch.sum(X2 * X2, dim=1).view(-1, 1)
    K = -2.0 * X1 @ X2.t() + n1.expand(X1.size(0), X2.size(0)) + n2.t().expand(X1.size(0), X2.size(0))
    K = scale * torch.exp(-0.5 * K)
    dist_z = (z1 / length_scale_z - z2 / length_scale_z) ** 2
    z_part = (-b * (z1 - h1) - b * (z2 - h2) - 0.5 * dist_z).exp()
    return z_part.mean() * (h1 - l1) * (h2 - l2) * K


def FIDES_predict(K, Kx, kxx_diag, noise, y):
    """base_gp/fides.py:74-109: like CIGP.forward but the variance stays a column [N*, 1]."""
    u, var = CIGP_predict(K, Kx, kxx_diag, noise, y)
    return u, var[:, :1]
