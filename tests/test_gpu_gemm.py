"""ffgp_gemm_f64 (include/ffgp.h) against torch.matmul: every operand layout, K-range mode, lower_only and batch level,
once with one CTA per tile and once with the persistent grid the batched path uses (tools/check_gemm.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('persist,n64', [('0', '1'), ('2', '1'), ('0', '2'), ('2', '2'), ('2', '0')])
def test_gemm_all_modes(persist, n64):
    # FFGP_GEMM_N64=2 forces the 128 x 64-tile variant (two CTAs per SM) wherever it is eligible, 0 disables it
    env = dict(os.environ, FFGP_PERSIST=persist, FFGP_GEMM_N64=n64)
    # the persistent grid is run three times: the stage-release race it once exposed hit ~1 tile in 500 (gemm_tma.cuh)
    for _ in range(3 if persist == '2' else 1):
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'check_gemm.py')], env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_factor256_kernel_all_modes():
    """factor256_kernel (csrc/factor256.cuh: potrf + trtri of a 256-block in one launch) against torch.linalg.cholesky
    and against the six-launch path it replaces, with and without the structural-zero skipping, single and batched
    problems, sizes on both sides of the 256 / 512 block boundaries, and a non-PD block reported through info
    (tools/f256_check.py runs FFGP_F256 = 0, 3, 2 in sub-processes: the switch is read once per process)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'f256_check.py')], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert r.stdout.count('non-PD detected') == 3, r.stdout[-4000:]
