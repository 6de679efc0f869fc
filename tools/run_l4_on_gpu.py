"""Run the reference's UNMODIFIED L4 code (train_CIGAR / train_AR / train_GAR / ..., gen-2023 AR / GAR / CIGAR) on the
CUDA drop-ins and compare with the CPU trajectories of the unmodified reference (tests/golden/l4_*.npz).

    python tools/run_l4_on_gpu.py [--ref DIR] [--json] [case ...]

DIR = a FidelityFusion checkout (default: $FF_REFERENCE, /root/reference, or baseline/_ref - a git-ignored scratch copy
made by tools/stage_reference.sh so that the tree travels to the GPU box).  This is INTEGRATION.md section 1 executed:
binding.install(), torch.set_default_device('cuda'), then the reference's own modules."""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def find_reference(explicit=None):
    for c in (explicit, os.environ.get('FF_REFERENCE'), '/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if c and os.path.isdir(os.path.join(c, 'FidelityFusion_Models')):
            return c
    return None


def main():
    args = sys.argv[1:]
    ref = None
    if '--ref' in args:
        i = args.index('--ref')
        ref = args[i + 1]
        del args[i:i + 2]
    as_json = '--json' in args
    args = [a for a in args if a != '--json']
    ref = find_reference(ref)
    if ref is None:
        print(json.dumps({'unavailable': 'no FidelityFusion tree found'}))
        return 0
    import numpy as np
    import torch
    warnings.filterwarnings('ignore')
    torch.set_default_dtype(torch.float64)                 # harness convention, SURVEY 8(c)
    import fidelityfusion_b200.binding as binding
    ours = binding.install(stub_missing_plotting=True)     # BEFORE the reference's modules are imported
    sys.path.insert(0, ref)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import l4_cases
    torch.set_default_device('cuda')                       # the reference builds torch.eye()/zeros() on the default device
    from fidelityfusion_b200 import _lib
    report = {'reference': ref, 'aliases': len(ours), 'cases': {}}
    os.chdir('/tmp')
    for name in (args or list(l4_cases.CASES)):
        n0 = _lib.lib().ffgp_launch_count()
        t0 = time.time()
        out = l4_cases.to_numpy(l4_cases.CASES[name]('cuda'))
        torch.cuda.synchronize()
        dt = time.time() - t0
        with np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')) as z:
            gold = {k: z[k] for k in z.files}
        errs = {}
        for k, g in gold.items():
            a = np.asarray(out[k], dtype=np.float64).reshape(g.shape)
            if not np.isfinite(g).all():
                errs[k] = 0.0 if np.array_equal(np.isfinite(a), np.isfinite(g)) else float('inf')
                continue
            errs[k] = float(np.max(np.abs(a - g)) / max(np.max(np.abs(g)), 1e-300))
        report['cases'][name] = {'seconds': round(dt, 3), 'ffgp_launches': int(_lib.lib().ffgp_launch_count() - n0),
                                 'max_rel_err': max(errs.values()), 'rel_err': errs}
        if not as_json:
            worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
            print(f'{name}: {dt:.2f}s, {report["cases"][name]["ffgp_launches"]} ffgp launches, worst: ' +
                  ', '.join(f'{k}={v:.1e}' for k, v in worst), flush=True)
    if as_json:
        print(json.dumps(report))
    return 0


if __name__ == '__main__':
    sys.exit(main())
