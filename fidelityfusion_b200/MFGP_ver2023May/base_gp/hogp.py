"""gen-2023 HOGP (reference MFGP_ver2023May/base_gp/hogp.py:34-240): Kronecker/Tucker GP over tensor outputs.
compute_loss = per-mode kernels -> per-mode eigh -> mode products -> fused core stage, all on libffgp kernels;
the gradient is the closed form w.r.t. each mode's kernel matrix (no differentiation through eigh)."""
import math

import torch

from ... import tensorly_compat as tl
from ..kernel.kernel_utils import create_kernels
from ..utils.dict_tools import update_dict_with_default
from ..utils.gp_noise import GP_noise_box

JITTER = 1e-6
EPS = 1e-10
PI = 3.1415


class eigen_pairs:
    def __init__(self, matrix=None, value=None, vector=None) -> None:
        if matrix is not None:
            value, vector = tl.eigh(matrix)
        self.value = value
        self.vector = vector


default_config = {
    'noise': {'init_value': 1., 'format': 'linear'},
    'kernel': [{'SE': {'noise_exp_format': True, 'length_scale': 1., 'scale': 1.}}],
    'learnable_grid': False,
    'learnable_mapping': False,
    'fidelity_shapes': None,
}


def _kron_outer(vectors):
    out = vectors[0].reshape(-1)
    for v in vectors[1:]:
        out = out.unsqueeze(-1) * v.reshape(-1)
    return out


class HOGP(torch.nn.Module):
    def __init__(self, gp_model_config=None) -> None:
        super().__init__()
        _final_config = update_dict_with_default(default_config, gp_model_config)
        self.gp_model_config = _final_config
        y_shape = self.gp_model_config['fidelity_shapes']
        if y_shape is None:
            raise ValueError('y_shape must be set as list')
        if isinstance(y_shape[0], list) or isinstance(y_shape[0], torch.Size):
            y_shape = y_shape[0]
        self.noise_box = GP_noise_box(self.gp_model_config['noise'])
        self.train_x = None
        self.train_y = None
        self.n_dim = len(y_shape)
        repeat_k_config = self.gp_model_config['kernel'] * (self.n_dim + 1)
        self.kernel_list = create_kernels(repeat_k_config)
        grid = []
        for _value in y_shape:
            grid.append(torch.nn.Parameter(torch.tensor(range(_value)).reshape(-1, 1).float()))
        if self.gp_model_config['learnable_grid'] is False:
            for g in grid:
                g.requires_grad = False
        self.grid = torch.nn.ParameterList(grid)
        mapping = []
        for _value in y_shape:
            mapping.append(torch.nn.Parameter(torch.eye(_value)))
        if self.gp_model_config['learnable_mapping'] is False:
            for m in mapping:
                m.requires_grad = False
        self.mapping_vector = torch.nn.ParameterList(mapping)

    def check_single_tensor(self, t):
        if isinstance(t, list):
            assert len(t) == 1, "HOGP model only support one input"
            t = t[0]
        return t

    def compute_kernel_cache(self):
        """hogp.py:120-137: per-mode kernel matrices (no jitter).  The eigenpairs are produced by the fused
        Kronecker op in compute_loss, which then fills self.eigen_cache."""
        kernel_result = [self.kernel_list[0](self.train_x, self.train_x)]
        for i in range(self.n_dim):
            _in = tl.mode_dot(self.grid[i], self.mapping_vector[i], 0)
            kernel_result.append(self.kernel_list[i + 1](_in, _in))
        self.k_result_cache = kernel_result

    def compute_loss(self, x, y, x_var=0., y_var=0., update_data=False):
        x = self.check_single_tensor(x)
        y = self.check_single_tensor(y)
        if self.train_x is None or update_data:
            self.train_x = x
            self.train_y = y
        self.compute_kernel_cache()
        # hogp.py:176 `A = A + y_var`: a number, or a tensor broadcast against A element by element
        add = y_var if isinstance(y_var, torch.Tensor) and (y_var.numel() > 1 or y_var.requires_grad) else float(y_var)
        val, A, g, eig = tl.kron_nll(self.train_y, self.k_result_cache, self.noise_box.get().pow(-1), add)
        self.eigen_cache = [eigen_pairs(value=lam, vector=U) for lam, U in eig]
        self.A = A
        self.g = g
        nd = A.numel()
        return (val + 0.5 * nd * math.log(2 * math.pi)) / nd

    def forward(self, x, x_vars=0.):
        x = self.check_single_tensor(x)
        with torch.no_grad():
            K_star = self.kernel_list[0](x, self.train_x)
            Ks = [k.detach() for k in self.k_result_cache]
            predict_u = tl.multi_mode_dot(self.g, [K_star] + Ks[1:])
            diag_K_dims = _kron_outer([_K.diag() for _K in Ks[1:]]).unsqueeze(0)
            diag_K_x = self.kernel_list[0](x, x).diag()
            for _ in range(self.n_dim):
                diag_K_x = diag_K_x.unsqueeze(-1)
            diag_K = diag_K_x * diag_K_dims
            S_2 = self.A                                     # (A * A^-1/2)^2, hogp.py:224-225
            # hogp.py:229: K* @ K0 + JITTER * eye(N*, N).pow(2)  (precedence as written)
            eye = torch.eye(K_star.shape[0], Ks[0].shape[0], dtype=K_star.dtype, device=K_star.device)
            fx = tl.mode_dot(Ks[0], K_star, 0) + JITTER * eye.pow(2)
            facs = [fx] + [self.eigen_cache[i + 1].vector.pow(2) for i in range(self.n_dim)]
            var_diag = diag_K + tl.multi_mode_dot(S_2, facs)
        return predict_u, var_diag
