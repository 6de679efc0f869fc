"""Noise-precision parameter box of the gen-2023 models (interface of reference MFGP_ver2023May/utils/gp_noise.py:9-24).

The state_dict key (`value`), the fp32 storage and the two parameterisations are the contract the GP mirrors and the
reference checkpoints rely on; the parameterisations live in one table so `get()` has no branches of its own.
"""
import torch

# format -> (init_value -> stored parameter, stored parameter -> precision)
def _identity(t):
    return t


_PARAMETERISATIONS = {
    'exp': (torch.log, torch.exp),
    'linear': (_identity, _identity),
}


class GP_noise_box(torch.nn.Module):
    def __init__(self, noise_config):
        super().__init__()
        fmt = noise_config['format']
        if fmt not in _PARAMETERISATIONS:
            raise AssertionError("noise format should be 'exp' or 'linear'")
        self.config, self.format = noise_config, fmt
        to_stored, self._to_precision = _PARAMETERISATIONS[fmt]
        init = torch.tensor(noise_config['init_value'], dtype=torch.float32)      # fp32-pinned, as the reference stores it
        self.value = torch.nn.Parameter(to_stored(init))

    def get(self):
        return self._to_precision(self.value)
