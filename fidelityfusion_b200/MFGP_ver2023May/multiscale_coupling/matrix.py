"""Matrix_Mapping coupling (reference MFGP_ver2023May/multiscale_coupling/matrix.py:8-90):
res = high - rho * (low x_1 W_1 ... x_M W_M).  The mode products run on the libffgp kernels (tensorly_compat.mode_dot ->
ffgp_mode_dot_f64) with gradients for the trainable W_k; rho is fp32-pinned and frozen unless `trainable_rho`.
state_dict keys: `vectors.k`, `rho`."""
import torch

from ... import tensorly_compat as tl
from ..utils.dict_tools import update_dict_with_default
from ._coupling import RhoCoupling

default_config = {
    'low_fidelity_shape': None,
    'high_fidelity_shape': None,
    'matrix_init_method': "smooth",
    'rho_value_init': 1.,
    'trainable_rho': False,
}


def _smooth_mapping_matrix(shape):
    """matrix.py:8-24: [h, l] interpolation weights 1 / ((i h/l - j)^2 + 1), columns normalised; identity when l == h."""
    l, h = shape
    if l == h:
        return torch.eye(l)
    assert l < h, NotImplemented
    dt = torch.get_default_dtype()
    src = torch.arange(l, dtype=dt).view(-1, 1) * (h / l)
    dst = torch.arange(h, dtype=dt).view(1, -1)
    weight = torch.ones(l, h) / ((src - dst) ** 2 + 1)
    return (weight / weight.sum(0, keepdim=True)).transpose(1, 0)


def _eye_distribution(shape):
    """matrix.py:26-34: the identity stretched to [l, h] by bilinear interpolation, transposed."""
    l, h = shape
    eye = torch.eye(l)
    if l == h:
        return eye
    if l < h:
        return torch.nn.functional.interpolate(eye.reshape(1, 1, l, l), (l, h), mode='bilinear').squeeze().T
    return None


_INIT = {'smooth': _smooth_mapping_matrix, 'eye': _eye_distribution}


class Matrix_Mapping(RhoCoupling):
    def __init__(self, config=None) -> None:
        super().__init__()
        self.config = update_dict_with_default(default_config, config)
        self.l_shape, self.h_shape = self.config['low_fidelity_shape'], self.config['high_fidelity_shape']
        assert self.l_shape is not None and self.h_shape is not None, \
            "low_fidelity_shape and high_fidelity_shape should be set"
        init = _INIT[self.config['matrix_init_method']]
        self.vectors = torch.nn.ParameterList(
            [torch.nn.Parameter(init((lo, hi)).contiguous()) for lo, hi in zip(self.l_shape, self.h_shape)])
        self._register_rho(self.config['rho_value_init'], self.config['trainable_rho'])

    def _map(self, low_fidelity):
        for mode, w in enumerate(self.vectors, start=1):
            low_fidelity = tl.mode_dot(low_fidelity, w, mode)
        return low_fidelity
