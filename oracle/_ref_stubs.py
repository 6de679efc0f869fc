"""Stub modules that let the UNMODIFIED reference import in this container.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py, never by the product).

The reference imports `matplotlib` (plots only) and `tensorly` (n-mode products)
at module top; neither is installed here and there is no network.  `matplotlib`
is registered as an empty module.  `tensorly` is replaced by the textbook n-mode
product definition (Kolda & Bader 2009, sec. 2.5), which is what tensorly's
pytorch backend computes: fold(M @ unfold(T, mode)).  SURVEY.md section 8(c)
records that tensorly is un-vendored and unpinned, so parity at that boundary is
anchored on the reference call sites and the KATs of SURVEY.md appendix B.
"""
import sys
import types

import torch


def _mode_dot(tensor, matrix_or_vector, mode, transpose=False):
    m = matrix_or_vector
    if m.ndim == 1:
        return torch.tensordot(tensor, m, dims=([mode], [0]))
    if transpose:
        m = m.T
    out = torch.tensordot(m, tensor, dims=([1], [mode]))
    return torch.movedim(out, 0, mode)


def _multi_mode_dot(tensor, matrix_or_vec_list, modes=None, skip=None, transpose=False):
    if modes is None:
        modes = list(range(len(matrix_or_vec_list)))
    res = tensor
    decrement = 0
    for i, (m, mode) in enumerate(zip(matrix_or_vec_list, modes)):
        if skip is not None and i == skip:
            continue
        res = _mode_dot(res, m, mode - decrement, transpose=transpose)
        if m.ndim == 1:
            decrement += 1
    return res


def install():
    if 'tensorly' in sys.modules and getattr(sys.modules['tensorly'], '_ffgp_stub', False):
        return
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', plt)

    tl = types.ModuleType('tensorly')
    tl._ffgp_stub = True
    tenalg = types.ModuleType('tensorly.tenalg')
    tenalg.mode_dot = _mode_dot
    tenalg.multi_mode_dot = _multi_mode_dot
    tl.tenalg = tenalg
    tl.set_backend = lambda *_a, **_k: None
    tl.tucker_to_tensor = lambda ct, **_k: _multi_mode_dot(ct[0], ct[1])
    tl.tensor_to_vec = lambda t: t.reshape(-1)
    tl.ones = lambda shape, **kw: torch.ones(shape, **{k: v for k, v in kw.items() if k in ('device', 'dtype')})
    sys.modules['tensorly'] = tl
    sys.modules['tensorly.tenalg'] = tenalg
