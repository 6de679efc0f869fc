"""Residual coupling (reference MFGP_ver2023May/multiscale_coupling/Residual.py:9-33): res = high - rho * low.
Config keys `rho_value_init` (1.) and `trainable` (True); state_dict key `rho`."""
from ..utils.dict_tools import update_dict_with_default
from ._coupling import RhoCoupling

default_config = {'rho_value_init': 1., 'trainable': True}


class Residual(RhoCoupling):
    def __init__(self, config=None) -> None:
        super().__init__()
        self.config = update_dict_with_default(default_config, config)
        self._register_rho(self.config['rho_value_init'], self.config['trainable'])
