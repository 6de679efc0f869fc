"""The C-ABI library builds, loads and exports every symbol include/ffgp.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from fidelityfusion_b200.csrc import build
    build.build()
    from fidelityfusion_b200 import _lib
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'ffgp.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ffgp_\w+)\s*\(', src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ('ffgp_dense_nll_f64', 'ffgp_dense_predict_f64', 'ffgp_kernel_matrix_f64', 'ffgp_potrf_trtri_f64',
              'ffgp_mode_dot_f64', 'ffgp_mode_gram_f64', 'ffgp_syevj_f64', 'ffgp_kron_core_f64'):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f'{s} declared in include/ffgp.h but not exported by libffgp.so'


def test_host_only_queries(lib):
    assert lib.ffgp_version() == 100
    # N=8192 single problem: three N^2 buffers + small vectors
    b = lib.ffgp_dense_workspace_bytes(8192, 16, 1, 0, 1)
    assert 3 * 8192 * 8192 * 8 <= b < 3.1 * 8192 * 8192 * 8
    # big batches are chunked: the workspace stays bounded (FFGP_WS_GB, default 32 GiB of the B200's 180 GB - BASELINE
    # config 5 runs as ONE 26 GiB chunk, four times as many problems still fit the same bound)
    assert 20 * 2**30 < lib.ffgp_dense_workspace_bytes(512, 8, 1, 64, 4096) < 33 * 2**30
    assert lib.ffgp_dense_workspace_bytes(512, 8, 1, 64, 16384) < 33 * 2**30
    assert lib.ffgp_dense_workspace_bytes(0, 8, 1, 0, 1) == 0
    assert lib.ffgp_syevj_workspace_bytes(128, 4) >= 4 * 128 * 128 * 8
    assert lib.ffgp_mode_gram_scratch_bytes(128, 512, 32, 32) > 0


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    rc = lib.ffgp_dense_nll_f64(None, None, None, None, None, None, 8, 2, 1, 1, 0, 0, 0, None, 0,
                                None, None, None, None, None, None, None, None, None)
    assert rc < 0 and b'null pointer' in lib.ffgp_last_error_string()
    rc = lib.ffgp_syevj_f64(None, 4, 1, None, None, None, 0, None, None)
    assert rc < 0


def _header_prototypes():
    """name -> (return type, [argument C types]) parsed from include/ffgp.h."""
    src = open(os.path.join(ROOT, 'include', 'ffgp.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r'\b(int|size_t|const char\s*\*|unsigned long long)\s+(ffgp_\w+)\s*\(([^)]*)\)\s*;', src):
        args = [a.strip() for a in args.split(',')] if args.strip() not in ('', 'void') else []
        protos[name] = (ret.replace(' ', ''), args)
    return protos


def test_ctypes_signatures_match_the_header(lib):
    """Every prototype of include/ffgp.h is bound in _lib.py with the same arity and the same pointer / integer / double
    kind per argument (an ABI drift between the header and the loader would otherwise only show up on the GPU box)."""
    protos = _header_prototypes()
    assert len(protos) >= 25
    for name, (ret, args) in protos.items():
        fn = getattr(lib, name)
        assert fn.argtypes is not None or not args, f'{name}: no argtypes declared in _lib.py'
        at = list(fn.argtypes or [])
        assert len(at) == len(args), f'{name}: header has {len(args)} arguments, _lib.py binds {len(at)}'
        for k, (c_arg, ct) in enumerate(zip(args, at)):
            if '*' in c_arg:
                want = (ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p)
            elif re.match(r'(const\s+)?double\b', c_arg):
                want = (ctypes.c_double,)
            elif re.match(r'(const\s+)?size_t\b', c_arg):
                want = (ctypes.c_size_t,)
            elif re.match(r'(const\s+)?long long\b', c_arg):
                want = (ctypes.c_longlong,)
            else:
                want = (ctypes.c_int,)
            assert ct in want, f'{name} argument {k} ({c_arg!r}) is bound as {ct}'
        want_ret = {'int': ctypes.c_int, 'size_t': ctypes.c_size_t, 'constchar*': ctypes.c_char_p,
                    'unsignedlonglong': ctypes.c_ulonglong}[ret]
        assert fn.restype is want_ret, f'{name}: return type {fn.restype} != {want_ret}'
