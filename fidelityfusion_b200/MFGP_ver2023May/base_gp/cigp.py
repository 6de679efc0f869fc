"""gen-2023 CIGP (reference MFGP_ver2023May/base_gp/cigp.py:19-136): stateful (caches train_x/train_y),
compute_loss returns +NLL, forward returns mean and DIAGONAL variance expanded to the mean's shape."""
import math

import torch

from ... import ops
from ..kernel.kernel_utils import create_kernel
from ..utils.dict_tools import update_dict_with_default
from ..utils.gp_noise import GP_noise_box

JITTER = 1e-6
EPS = 1e-10
PI = 3.1415

default_config = {
    'noise': {'init_value': 1., 'format': 'exp'},
    'kernel': {'SE': {'noise_exp_format': True, 'length_scale': 1., 'scale': 1.}},
}


class CIGP(torch.nn.Module):
    def __init__(self, gp_model_config=None) -> None:
        super().__init__()
        _final_config = update_dict_with_default(default_config, gp_model_config)
        self.gp_model_config = _final_config
        self.noise_box = GP_noise_box(self.gp_model_config['noise'])
        self.kernel = create_kernel(self.gp_model_config['kernel'])
        self.train_x = None
        self.train_y = None
        self.factor_cache = ops.FactorCache()

    def check_single_tensor(self, t):
        if isinstance(t, list):
            assert len(t) == 1, "CIGP model only support one input"
            t = t[0]
        return t

    def forward(self, x, x_var=0.):
        x = self.check_single_tensor(x)
        if self.train_x is None:
            print("gp model model hasn't been trained. predict failed")
            return None
        with torch.no_grad():
            noise_inv = self.noise_box.get().pow(-1)
            inv_ls, amp, clamp = self.kernel.fused_params()
            u, var = ops.dense_predict(self.train_x, self.train_y, x, inv_ls, amp,
                                       diag_add=(noise_inv + JITTER).reshape(1), cov_offset=noise_inv,
                                       full_cov=False, clamp=clamp, cache=self.factor_cache,
                                       cache_token=ops.state_token(self, self.train_x, self.train_y))
            var_diag = var.view(-1, 1).expand_as(u) + x_var
        return u, var_diag

    def compute_loss(self, x, y, x_var=0., y_var=0., update_data=False):
        x = self.check_single_tensor(x)
        y = self.check_single_tensor(y)
        assert y.ndim == 2, "y should be 2d tensor"
        if self.train_x is None or update_data:
            self.train_x = x
            self.train_y = y
        self.factor_cache.invalidate()
        n, D = y.shape
        diag = (self.noise_box.get().pow(-1) + JITTER).reshape(1)
        sigma_add = None
        if isinstance(y_var, torch.Tensor):
            if y_var.numel() == 1:
                sigma_add = y_var.reshape(1, 1).expand(n, n)       # a scalar is added to EVERY entry (cigp.py:127)
            else:
                sigma_add = y_var
        elif y_var != 0.:
            sigma_add = torch.full((n, n), float(y_var), dtype=y.dtype, device=y.device)
        inv_ls, amp, clamp = self.kernel.fused_params()
        core = ops.dense_nll(x, y, inv_ls, amp, diag_add=diag, sigma_add=sigma_add, clamp=clamp)
        return core + 0.5 * n * D * math.log(2 * PI)
