import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _default_dtype_f64():
    """Harness convention of SURVEY.md 8(c): default dtype fp64 (tests that exercise fp32 switch explicitly)."""
    import torch
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """norm-wise relative error |a-b|_inf / max(|b|_inf, tiny)."""
    def arr(v):
        if hasattr(v, 'detach'):
            v = v.detach().cpu().numpy()
        return np.asarray(v, dtype=np.float64)
    a, b = arr(a), arr(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b))) / denom if b.size else 0.0
