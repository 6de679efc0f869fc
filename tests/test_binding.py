"""The drop-in boundary, proven on the reference's OWN L4 code.

fidelityfusion_b200.binding.install() aliases our operator modules under the reference's module names; after that
the UNMODIFIED files FidelityFusion_Models/*.py and MFGP_ver2023May/*.py import, construct and - on a GPU - train.
These tests need a FidelityFusion checkout: /root/reference in the build container, or baseline/_ref (a git-ignored
scratch copy made by tools/stage_reference.sh, which travels to the GPU box).  Without one they skip, and
tests/test_gpu_l4_replay.py replays the same compositions against the same golden trajectories instead.

Each test runs in a fresh interpreter: install() edits sys.modules, which must not leak into the other tests (they
import the reference under oracle/_ref_stubs.py)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'tools'))
from run_l4_on_gpu import find_reference  # noqa: E402

REF = find_reference()
needs_ref = pytest.mark.skipif(REF is None, reason='no FidelityFusion tree (/root/reference or baseline/_ref)')

_CONSTRUCT = r'''
import sys, json, warnings
warnings.filterwarnings('ignore')
sys.path.insert(0, %(root)r)
import torch
torch.set_default_dtype(torch.float64)
import fidelityfusion_b200.binding as binding
ours = binding.install(stub_missing_plotting=True)
sys.path.insert(0, %(ref)r)
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    import FidelityFusion_Models as FFM                       # __init__ pulls AR, CIGAR, CAR, GAR, NAR, ResGP (:1-8)
    import FidelityFusion_Models.CIGAR, FidelityFusion_Models.GAR, FidelityFusion_Models.AR_autoRegression
    import FidelityFusion_Models.ResGP, FidelityFusion_Models.NAR, FidelityFusion_Models.CAR_ContinuousAutoRegression
    import MFGP_ver2023May as G23                             # __init__ pulls AR, CIGAR, GAR, CAR, NAR, ResGP, CIGP, HOGP
    import GaussianProcess.kernel as kernel
    import tensorly
    tensorly.set_backend('pytorch')                           # what 6 reference files do at import
res = {'aliases': sorted(ours), 'l4_from_reference': all(
    m.__file__.startswith(%(ref)r) for m in (sys.modules['FidelityFusion_Models.CIGAR'], sys.modules['FidelityFusion_Models.GAR'],
                                              sys.modules['MFGP_ver2023May.AR_AutoRegression'], sys.modules['MFGP_ver2023May.GAR_GeneralizedAutoAR']))}
leaves = lambda m: sorted({type(s).__module__ for s in m.modules()
                           if s is not m and not isinstance(s, (torch.nn.ModuleList, torch.nn.ParameterList))})
K = kernel.SquaredExponentialKernel
mods = {
    'AR':    FFM.AR(3, [K() for _ in range(3)], if_nonsubset=True),
    'ResGP': FFM.ResGP(2, [K() for _ in range(2)], if_nonsubset=True),
    'NAR':   FFM.NAR(2, [kernel.ARDKernel(2), kernel.ARDKernel(3)], if_nonsubset=True),
    'CIGAR': FFM.CIGAR(3, [kernel.ARDKernel(5) for _ in range(3)], [(16,), (32,), (64,)], if_nonsubset=True),
    'GAR':   FFM.GAR(2, [K() for _ in range(2)], [(4, 4), (4, 4)], if_nonsubset=True),
    'AR23':    G23.AR({'fidelity_shapes': [(4,), (4,)]}),
    'ResGP23': G23.ResGP({'fidelity_shapes': [(4,), (4,)]}),
    'NAR23':   G23.NAR({'fidelity_shapes': [(4,), (4,)]}),
    'CAR23':   G23.CAR({'fidelity_shapes': [(4,), (4,)]}),
    'CIGAR23': G23.CIGAR({'fidelity_shapes': [torch.Size([4]), torch.Size([4])]}),
    'GAR23':   G23.GAR({'fidelity_shapes': [torch.Size([4, 3]), torch.Size([4, 3])]}),
    'CIGP23':  G23.CIGP(None), 'HOGP23': G23.HOGP({'fidelity_shapes': [torch.Size([3, 2])]}),
}
res['submodules'] = {k: leaves(m) for k, m in mods.items()}
res['state_dict_keys'] = {k: sorted(m.state_dict()) for k, m in mods.items() if k in ('CIGAR', 'AR23', 'GAR23')}
print('RESULT ' + json.dumps(res))
'''


def _run(code, timeout=1200):
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=timeout, cwd='/tmp')
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')][-1]
    return json.loads(line[len('RESULT '):])


@needs_ref
def test_reference_l4_imports_and_constructs_on_the_drop_ins():
    """INTEGRATION.md section 1 in a fresh interpreter: every L4 module of both generations imports from the
    reference tree, every GP / kernel / coupling sub-module of the constructed models is OURS."""
    res = _run(_CONSTRUCT % {'root': ROOT, 'ref': REF})
    assert res['l4_from_reference'], 'the L4 model files must be the reference\'s own'
    for name, leaves in res['submodules'].items():
        assert leaves, name
        foreign = [m for m in leaves if not m.startswith('fidelityfusion_b200.')]
        assert not foreign, f'{name}: sub-modules not routed to the drop-in: {foreign}'
    # GAR.py:6 imports the two_fidelity_models copy of HOGP_simple: it must be ours too
    assert 'fidelityfusion_b200.GaussianProcess.hogp_simple' in res['submodules']['GAR']
    assert 'gpr_list.0.kernel.length_scales' in res['state_dict_keys']['CIGAR']
    assert 'Tensor_linear_list.1.vectors.0' in res['state_dict_keys']['CIGAR']
    assert 'cigp_list.0.noise_box.value' in res['state_dict_keys']['AR23']
    assert 'hogp_list.1.kernel_list.2.length_scale' in res['state_dict_keys']['GAR23']


def test_install_refuses_a_half_binding():
    """If the reference's operator modules were imported first, their classes are already bound: install() must say so."""
    code = ("import sys, types; sys.path.insert(0, %r); sys.modules['GaussianProcess'] = types.ModuleType('GaussianProcess');"
            "sys.modules['GaussianProcess.kernel'] = types.ModuleType('GaussianProcess.kernel');"
            "import fidelityfusion_b200.binding as b\n"
            "try:\n    b.install()\n    print('RESULT \"no error\"')\n"
            "except RuntimeError as e:\n    print('RESULT \"refused\"')\n") % ROOT
    assert _run(code) == 'refused'


def test_alias_table_covers_every_operator_module_the_l4_files_import():
    """Static check against the reference tree when present, else against the recorded import list of SURVEY 8b."""
    import fidelityfusion_b200.binding as binding
    table = {**binding.ALIASES, **binding.DATA_ALIASES, **binding.ACQ_ALIASES}
    need = ['GaussianProcess.kernel', 'GaussianProcess.cigp_v10', 'GaussianProcess.gp_computation_pack',
            'GaussianProcess.gp_basic', 'FidelityFusion_Models.two_fidelity_models.hogp_simple',
            'MFGP_ver2023May.base_gp.cigp', 'MFGP_ver2023May.base_gp.hogp', 'MFGP_ver2023May.base_gp.fides',
            'MFGP_ver2023May.multiscale_coupling.matrix', 'MFGP_ver2023May.multiscale_coupling.Residual',
            'tensorly', 'tensorly.tenalg']
    assert not [n for n in need if n not in table]
    import importlib.util
    for ref_name, (mod, _) in table.items():
        assert importlib.util.find_spec(mod) is not None, mod
    if REF is not None:                                   # every target names a file that exists in the reference
        for ref_name, (_, path) in table.items():
            if not path.startswith('tensorly'):
                assert os.path.exists(os.path.join(REF, path)), path


# north-star tolerance (1e-9 relative) on every recorded quantity: per-iteration losses, final parameters, first-step
# gradients, predictions.  Measured on B200 (profiles/r02_l4_binding_on_b200.txt): dense paths 1e-14, Kronecker paths
# <= 2.2e-10 after 8 Adam steps although the reference differentiates THROUGH eigh and we use the closed form.
# EI / PI carry the float32-rounded normal cdf / pdf of the reference (acq.py:178, 230): one float32 ulp where erfc and
# scipy's ndtr round differently
_LOOSE = {'l4_bo_cigp_acq': {'ei': 2e-7, 'pi': 2e-7}}


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('case', ['l4_cigar3_c3', 'l4_ar3_nonsubset', 'l4_resgp2_nonsubset', 'l4_nar2_nonsubset',
                                  'l4_gar2_c4', 'l4_ar2023_c1', 'l4_gar2023_c4', 'l4_cigar2023', 'l4_bo_cigp_acq', 'l4_family2023'])
def test_unmodified_reference_l4_trains_on_cuda_drop_ins(case):
    """The reference's own train_* / compute_loss / forward on .cuda() models after binding.install(): per-iteration
    losses, final parameters and predictions against the SAME code on the CPU reference (tests/golden/l4_*.npz)."""
    if case == 'l4_bo_cigp_acq' and not os.path.exists(os.path.join(REF, 'Bayesian_optimization', 'cigp.py')):
        pytest.skip('the staged reference tree predates tools/stage_reference.sh staging Bayesian_optimization/')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_l4_on_gpu.py'), '--ref', REF, '--json', case],
                       capture_output=True, text=True, timeout=1800, cwd='/tmp')
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])['cases'][case]
    assert rep['ffgp_launches'] > 0, 'no libffgp kernel ran: the binding fell through to torch'
    loose = _LOOSE.get(case, {})
    bad = {k: v for k, v in rep['rel_err'].items() if not v < loose.get(k, 1e-9)}
    assert not bad, f'{case}: {bad}'
