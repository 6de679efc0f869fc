import torch

from .SE_kernel import SE_kernel


def create_kernels(kernel_configs):
    """reference MFGP_ver2023May/kernel/kernel_utils.py:5-16 (passes the config dict positionally)."""
    kernel_list = []
    for _k_config in kernel_configs:
        for kernel_name, kernel_config in _k_config.items():
            if kernel_name == 'SE':
                kernel_list.append(SE_kernel(kernel_config))
            else:
                raise NotImplementedError
    return torch.nn.ModuleList(kernel_list)


def create_kernel(kernel_config):
    """reference kernel_utils.py:18-28."""
    if isinstance(kernel_config, list) and len(kernel_config) == 1:
        kernel_config = kernel_config[0]
    for kernel_name, kernel_config in kernel_config.items():
        if kernel_name == 'SE':
            return SE_kernel(kernel_config)
        elif kernel_name == 'kernel_res':
            from .MCMC_res_kernel import Kernel_res
            return Kernel_res(kernel_config)
        raise NotImplementedError
