"""The reference-side binding: route FidelityFusion's GP operator modules to the B200 implementation.

The reference (IceLab-X/FidelityFusion) has no FFI; its plug-in boundary for the GP hot path is the module namespace
that its fusion models import (`FidelityFusion_Models/*.py:6-8`, `MFGP_ver2023May/*.py:5-8`).  `install()` registers
our operator modules under those names in `sys.modules`, so the reference's own L4 code (`train_CIGAR`, `train_AR`,
`train_GAR`, gen-2023 `AR/CIGAR/GAR.compute_loss`, ...) runs UNMODIFIED on libffgp's CUDA kernels.

    import fidelityfusion_b200.binding as binding
    binding.install()                                    # before the first `import FidelityFusion_Models...`
    torch.set_default_device('cuda')                     # the reference creates torch.eye()/zeros() on the default device
    from FidelityFusion_Models.CIGAR import CIGAR, train_CIGAR

tests/test_binding.py runs exactly this against the unmodified reference tree."""
from __future__ import annotations

import importlib
import sys
import types

_P = 'fidelityfusion_b200.'

# reference module name -> (our module, reference file it stands in for)
ALIASES = {
    # gen-2024 operators (GaussianProcess/)
    'GaussianProcess.kernel':              (_P + 'GaussianProcess.kernel',              'GaussianProcess/kernel.py'),
    'GaussianProcess.cigp_v10':            (_P + 'GaussianProcess.cigp_v10',            'GaussianProcess/cigp_v10.py'),
    'GaussianProcess.gp_computation_pack': (_P + 'GaussianProcess.gp_computation_pack', 'GaussianProcess/gp_computation_pack.py'),
    'GaussianProcess.gp_basic':            (_P + 'GaussianProcess.gp_basic',            'GaussianProcess/gp_basic.py'),
    'GaussianProcess.hogp_simple':         (_P + 'GaussianProcess.hogp_simple',         'GaussianProcess/hogp_simple.py'),
    # the HOGP_simple copy that FidelityFusion_Models/GAR.py:6 imports
    'FidelityFusion_Models.two_fidelity_models.hogp_simple':
        (_P + 'FidelityFusion_Models.two_fidelity_models.hogp_simple', 'FidelityFusion_Models/two_fidelity_models/hogp_simple.py'),
    # gen-2023 operators (MFGP_ver2023May/)
    'MFGP_ver2023May.base_gp.cigp':                 (_P + 'MFGP_ver2023May.base_gp.cigp',  'MFGP_ver2023May/base_gp/cigp.py'),
    'MFGP_ver2023May.base_gp.hogp':                 (_P + 'MFGP_ver2023May.base_gp.hogp',  'MFGP_ver2023May/base_gp/hogp.py'),
    'MFGP_ver2023May.base_gp.fides':                (_P + 'MFGP_ver2023May.base_gp.fides', 'MFGP_ver2023May/base_gp/fides.py'),
    'MFGP_ver2023May.kernel.SE_kernel':             (_P + 'MFGP_ver2023May.kernel.SE_kernel',       'MFGP_ver2023May/kernel/SE_kernel.py'),
    'MFGP_ver2023May.kernel.kernel_utils':          (_P + 'MFGP_ver2023May.kernel.kernel_utils',    'MFGP_ver2023May/kernel/kernel_utils.py'),
    'MFGP_ver2023May.kernel.MCMC_res_kernel':       (_P + 'MFGP_ver2023May.kernel.MCMC_res_kernel', 'MFGP_ver2023May/kernel/MCMC_res_kernel.py'),
    'MFGP_ver2023May.multiscale_coupling.matrix':   (_P + 'MFGP_ver2023May.multiscale_coupling.matrix',   'MFGP_ver2023May/multiscale_coupling/matrix.py'),
    'MFGP_ver2023May.multiscale_coupling.Residual': (_P + 'MFGP_ver2023May.multiscale_coupling.Residual', 'MFGP_ver2023May/multiscale_coupling/Residual.py'),
    # the un-vendored third-party n-mode products (call sites hogp.py:132.., matrix.py:73,81, gp_computation_pack.py:157)
    'tensorly':        (_P + 'tensorly_compat', 'tensorly (third party, not in the tree)'),
    'tensorly.tenalg': (_P + 'tensorly_compat', 'tensorly.tenalg'),
}

# optional: the data side and the acquisition functions (8f-2)
DATA_ALIASES = {
    'FidelityFusion_Models.MF_data': (_P + 'FidelityFusion_Models.MF_data', 'FidelityFusion_Models/MF_data.py'),
}
ACQ_ALIASES = {
    'MF_BayesianOptimization.Discrete.DMF_acq':
        (_P + 'MF_BayesianOptimization.Discrete.DMF_acq', 'MF_BayesianOptimization/Discrete/DMF_acq.py'),
    'Bayesian_optimization.acq': (_P + 'Bayesian_optimization.acq', 'Bayesian_optimization/acq.py'),
}

_installed = {}


def _stub_plotting():
    """matplotlib is imported at module top by every reference model file for its __main__ demo only.  When it is not
    installed, an empty module lets those files import; nothing on the GP path touches it."""
    try:
        import matplotlib.pyplot  # noqa: F401
        return
    except Exception:
        pass
    mpl, plt = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    sys.modules['matplotlib'], sys.modules['matplotlib.pyplot'] = mpl, plt


def install(data_manager=True, acquisition=True, stub_missing_plotting=False):
    """Register the drop-in modules under the reference's module names.  Call before the reference's model modules
    are imported (a module the reference already imported keeps the classes it bound at import time: that is an
    error here, not a silent half-binding).  Returns {reference name: our module}."""
    table = dict(ALIASES)
    if data_manager:
        table.update(DATA_ALIASES)
    if acquisition:
        table.update(ACQ_ALIASES)
    ours = {ref: importlib.import_module(mod) for ref, (mod, _) in table.items()}
    clash = [ref for ref, m in ours.items() if ref in sys.modules and sys.modules[ref] is not m]
    if clash:
        raise RuntimeError('binding.install(): already imported from the reference: ' + ', '.join(sorted(clash)) +
                           ' - call install() before importing FidelityFusion_Models / MFGP_ver2023May / GaussianProcess')
    if stub_missing_plotting:
        _stub_plotting()
    sys.modules.update(ours)
    _installed.update(ours)
    return ours


def uninstall():
    """Remove the aliases (tests)."""
    for ref, m in list(_installed.items()):
        if sys.modules.get(ref) is m:
            del sys.modules[ref]
    _installed.clear()


def is_ours(obj):
    """True if obj's class is defined under fidelityfusion_b200 (used by tests to prove nothing fell through to the
    reference's torch implementation)."""
    return type(obj).__module__.startswith('fidelityfusion_b200.')
