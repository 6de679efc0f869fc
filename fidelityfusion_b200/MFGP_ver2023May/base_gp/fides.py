"""gen-2023 FIDES (reference MFGP_ver2023May/base_gp/fides.py:23-151): CIGP with the residual kernel
`Kernel_res(X1, X2, l1, h1, l2, h2)` whose fidelity bounds are set by `set_fidelity`.  compute_loss returns +NLL with
the reference's PI = 3.1415; forward returns the mean [N*, D] and the DIAGONAL variance as a column [N*, 1]
(not expanded, unlike CIGP)."""
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
    'kernel': {'kernel_res': {'noise_exp_format': True, 'length_scale': 1., 'scale': 1., 'length_scale_z': 1.}},
}


class FIDES(torch.nn.Module):
    def __init__(self, config) -> None:
        super().__init__()
        _final_config = update_dict_with_default(default_config, config)
        self.config = _final_config
        self.noise_box = GP_noise_box(self.config['noise'])
        self.train_x = None
        self.train_y = None
        self.kernel = create_kernel(self.config['kernel'])
        self.fi_define = False
        self.factor_cache = ops.FactorCache()

    def check_single_tensor(self, t):
        if isinstance(t, list):
            assert len(t) == 1, "CIGP model only support one input"
            t = t[0]
        return t

    def set_fidelity(self, l1, h1, l2, h2):
        self.l1, self.h1, self.l2, self.h2 = l1, h1, l2, h2
        self.fi_define = True
        self.factor_cache.invalidate()

    def forward(self, x, x_var=0.):
        x = self.check_single_tensor(x)
        if self.train_x is None:
            print("gp model model hasn't been trained. predict failed")
            return None
        with torch.no_grad():       # parameters are constants of the posterior (the reference's callers do not train through it)
            noise_inv = self.noise_box.get().pow(-1)
            inv_ls, amp, clamp = self.kernel.fused_params(self.l1, self.h1, self.l2, self.h2)
            mean, var = ops.dense_predict(self.train_x, self.train_y, x, inv_ls, amp,
                                          diag_add=(noise_inv + JITTER).reshape(1), cov_offset=noise_inv,
                                          full_cov=False, clamp=clamp, cache=self.factor_cache,
                                          cache_token=(ops.state_token(self, self.train_x, self.train_y),
                                                       self.l1, self.h1, self.l2, self.h2))
        return mean, var.view(-1, 1)

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
        inv_ls, amp, clamp = self.kernel.fused_params(self.l1, self.h1, self.l2, self.h2)
        core = ops.dense_nll(x, y, inv_ls, amp, diag_add=diag, clamp=clamp)     # y_var is ignored, fides.py:140-149
        return core + 0.5 * n * D * math.log(2 * PI)
