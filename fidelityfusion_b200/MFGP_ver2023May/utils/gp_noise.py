import torch


class GP_noise_box(torch.nn.Module):
    """reference MFGP_ver2023May/utils/gp_noise.py:9-24: precision parameter, 'exp' or 'linear' format, fp32-pinned."""

    def __init__(self, noise_config):
        super().__init__()
        assert noise_config['format'] in ['exp', 'linear'], "noise format should be 'exp' or 'linear'"
        self.config = noise_config
        self.format = noise_config['format']
        if self.format == 'exp':
            self.value = torch.nn.Parameter(torch.log(torch.tensor(noise_config['init_value'], dtype=torch.float32)))
        else:
            self.value = torch.nn.Parameter(torch.tensor(noise_config['init_value'], dtype=torch.float32))

    def get(self):
        if self.format == 'exp':
            return torch.exp(self.value)
        return self.value
