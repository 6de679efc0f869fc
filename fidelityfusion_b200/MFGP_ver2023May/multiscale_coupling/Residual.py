"""Residual coupling (reference MFGP_ver2023May/multiscale_coupling/Residual.py:9-33): res = high - rho * low."""
import torch

from ..utils.dict_tools import update_dict_with_default

default_config = {
    'rho_value_init': 1.,
    'trainable': True,
}


class Residual(torch.nn.Module):
    def __init__(self, config=None) -> None:
        super().__init__()
        self.config = update_dict_with_default(default_config, config)
        self.rho = torch.nn.Parameter(torch.tensor(self.config['rho_value_init'], dtype=torch.float32))
        self.rho.requires_grad = bool(self.config['trainable'])

    def forward(self, low_fidelity, high_fidelity):
        return high_fidelity - low_fidelity * self.rho

    def backward(self, low_fidelity, res):
        return low_fidelity * self.rho + res

    def var_forward(self, low_fidelity_var, high_fidelity_var):
        return high_fidelity_var - low_fidelity_var * self.rho

    def var_backward(self, low_fidelity_var, res_var):
        return low_fidelity_var * self.rho + res_var
