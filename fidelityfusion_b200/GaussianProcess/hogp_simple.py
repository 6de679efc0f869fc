"""HOGP_simple, the Kronecker/Tucker GP for tensor outputs, in the reference's two variants:
  - GaussianProcess/hogp_simple.py:21-113            (`variant='gp'`: y_var added to K_0, variance factor K* K_0, no_grad)
  - FidelityFusion_Models/two_fidelity_models/hogp_simple.py:21-126 (`variant='ffm'`: variance factor (K* K_0^-1 U_0)^2)
Both share ONE kernel object across all modes (hogp_simple.py:29-31) and return the positive, nd-normalised loss
from `log_likelihood` despite its name."""
import math

import torch
import torch.nn as nn

from .. import tensorly_compat as tl


class eigen_pairs:
    def __init__(self, matrix) -> None:
        self.value, self.vector = tl.eigh(matrix)


def _kron_outer(vectors):
    out = vectors[0].reshape(-1)
    for v in vectors[1:]:
        out = out.unsqueeze(-1) * v.reshape(-1)
    return out


class HOGP_simple(nn.Module):
    variant = 'gp'

    def __init__(self, kernel, noise_variance, output_shape, learnable_grid=False, learnable_map=False):
        super().__init__()
        self.noise_variance = nn.Parameter(torch.tensor([noise_variance]))
        self.K = []
        self.K_eigen = []
        self.kernel_list = nn.ModuleList()
        for _ in range(len(output_shape) + 1):
            self.kernel_list.append(kernel)
        self.grid = nn.ParameterList()
        for _value in output_shape:
            self.grid.append(nn.Parameter(torch.tensor(range(_value)).reshape(-1, 1).float()))
        if learnable_grid is False:
            for i in range(len(self.grid)):
                self.grid[i].requires_grad = False
        self.mapping_vector = nn.ParameterList()
        for _value in output_shape:
            self.mapping_vector.append(nn.Parameter(torch.eye(_value)))
        if learnable_map is False:
            for i in range(len(self.mapping_vector)):
                self.mapping_vector[i].requires_grad = False

    def log_likelihood(self, x_train, y_train):
        if isinstance(y_train, list):
            y_train_var = y_train[1]
            y_train = y_train[0]
        else:
            y_train_var = None
        self.K.clear()
        self.K_eigen.clear()
        K0 = self.kernel_list[0](x_train, x_train)
        if y_train_var is not None and self.variant == 'gp':
            K0 = K0 + y_train_var                      # GaussianProcess/hogp_simple.py:83-84; the ffm copy ignores it
        self.K.append(K0)
        for i in range(0, len(self.kernel_list) - 1):
            _in = tl.mode_dot(self.grid[i], self.mapping_vector[i], 0)
            self.K.append(self.kernel_list[i + 1](_in, _in))
        val, A, g, eig = tl.kron_nll(y_train, self.K, self.noise_variance.pow(-1))
        for lam, U in eig:
            ep = eigen_pairs.__new__(eigen_pairs)
            ep.value, ep.vector = lam, U
            self.K_eigen.append(ep)
        self.A = A
        self.g = g
        nd = A.numel()
        return (val + 0.5 * nd * math.log(2 * math.pi)) / nd

    def forward(self, x_train, x_test):
        with torch.no_grad():
            K_star = self.kernel_list[0](x_test, x_train)
            predict_u = tl.multi_mode_dot(self.g, [K_star] + [k.detach() for k in self.K[1:]])
            n_dim = len(self.K_eigen) - 1
            diag_K_dims = _kron_outer([K.detach().diag() for K in self.K[1:]]).unsqueeze(0)
            # K(x*,x*)_ii of a stationary kernel: evaluate the diagonal through the kernel itself
            diag_K_x = self.kernel_list[0](x_test, x_test).diag()
            for _ in range(n_dim):
                diag_K_x = diag_K_x.unsqueeze(-1)
            diag_K = diag_K_x * diag_K_dims
            S_2 = self.A                                   # (A * A^-1/2)^2, hogp_simple.py:61-62
            K0 = self.K[0].detach()
            if self.variant == 'ffm':
                # (K* K0^-1 U0)^2 with K0^-1 U0 = U0 diag(1/lambda0)
                e0 = self.K_eigen[0]
                fx = (tl.mode_dot(e0.vector.contiguous(), K_star, 0) / e0.value.unsqueeze(0)).pow(2)
            else:
                fx = tl.mode_dot(K0, K_star, 0)            # K* @ K0
            facs = [fx] + [self.K_eigen[i + 1].vector.pow(2) for i in range(n_dim)]
            var_diag = diag_K + tl.multi_mode_dot(S_2, facs)
        return predict_u, var_diag


class HOGP_simple_ffm(HOGP_simple):
    """The copy that FidelityFusion_Models/GAR.py:6 imports (two_fidelity_models/hogp_simple.py)."""
    variant = 'ffm'
