"""Run the UNMODIFIED reference's L4 training / prediction entry points (oracle/l4_cases.py) on torch-CPU and write
tests/golden/l4_*.npz.  TEST INFRASTRUCTURE; run once in the build container:  python oracle/gen_golden_l4.py [case ...]

The same case functions are run by tests/test_binding.py on the CUDA drop-ins (fidelityfusion_b200.binding.install())
and compared with these fixtures."""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '..', 'tests', 'golden')
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import _ref_stubs  # noqa: E402

_ref_stubs.install()
import torch  # noqa: E402

warnings.filterwarnings('ignore')
torch.set_default_dtype(torch.float64)
import l4_cases  # noqa: E402

if __name__ == '__main__':
    names = sys.argv[1:] or list(l4_cases.CASES)
    cwd = os.getcwd()
    os.chdir('/tmp')          # Experiments.log_debugger (if ever constructed) writes under the cwd; keep the repo clean
    for name in names:
        t0 = time.time()
        out = l4_cases.to_numpy(l4_cases.CASES[name]('cpu'))
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print(f'wrote {name} in {time.time() - t0:.1f}s', {k: v.shape for k, v in out.items()})
    os.chdir(cwd)
