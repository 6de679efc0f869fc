"""Kernel factories of the gen-2023 configs (reference MFGP_ver2023May/kernel/kernel_utils.py:5-28).  A config is
{kernel_name: kernel_config}; the kernel_config DICT is handed to the kernel class positionally, exactly like the
reference does - which is why `noise_exp_format is True` is false for config-built kernels."""
import torch

from .SE_kernel import SE_kernel


def _res_kernel(cfg):
    from .MCMC_res_kernel import Kernel_res
    return Kernel_res(cfg)


_SINGLE = {'SE': SE_kernel, 'kernel_res': _res_kernel}      # create_kernel (kernel_utils.py:18-28)
_LISTED = {'SE': SE_kernel}                                  # create_kernels only knows SE (kernel_utils.py:5-16)


def _build(table, name, cfg):
    if name not in table:
        raise NotImplementedError
    return table[name](cfg)


def create_kernels(kernel_configs):
    return torch.nn.ModuleList([_build(_LISTED, name, cfg) for entry in kernel_configs for name, cfg in entry.items()])


def create_kernel(kernel_config):
    if isinstance(kernel_config, list) and len(kernel_config) == 1:
        kernel_config = kernel_config[0]
    for name, cfg in kernel_config.items():
        return _build(_SINGLE, name, cfg)          # the first entry decides, like the reference's early return
