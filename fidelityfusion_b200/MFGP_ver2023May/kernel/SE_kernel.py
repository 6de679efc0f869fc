import torch

from ... import ops


class SE_kernel(torch.nn.Module):
    """reference MFGP_ver2023May/kernel/SE_kernel.py:4-44.  `noise_exp_format is True` is tested by identity, so a
    config DICT passed by kernel_utils.create_kernel (kernel_utils.py:12,24) selects the linear format, and the
    Parameters then wrap that dict's tensors... the reference fails there unless length_scale/scale are given;
    with the dict as first positional argument length_scale=1., scale=1. (raw) are used - reproduced."""

    def __init__(self, noise_exp_format, length_scale=1., scale=1.) -> None:
        super().__init__()
        self.noise_exp_format = noise_exp_format
        length_scale = torch.tensor(length_scale)
        scale = torch.tensor(scale)
        if noise_exp_format is True:
            self.length_scale = torch.nn.Parameter(torch.log(length_scale))
            self.scale = torch.nn.Parameter(torch.log(scale))
        else:
            self.length_scale = torch.nn.Parameter(length_scale)
            self.scale = torch.nn.Parameter(scale)

    def fused_params(self):
        if self.noise_exp_format is True:
            return torch.exp(-self.length_scale).reshape(-1), torch.exp(self.scale).reshape(-1), False
        return (1.0 / self.length_scale).reshape(-1), self.scale.reshape(-1), False

    def forward(self, X, X2):
        inv_ls, amp, clamp = self.fused_params()
        return ops.kernel_matrix(X, X2, inv_ls, amp, clamp)
