"""ctypes binding of libffgp.so (the C ABI declared in include/ffgp.h).

The product path has NO CPU fallback: if the shared library is missing or a call is made
with non-CUDA tensors this module raises.  Build the library with `python __graft_entry__.py`
(or fidelityfusion_b200/csrc/build.py)."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFGP_LIB_PATH: A/B runs of two builds of the SAME C ABI on one box (tools/); production loads the in-tree library
LIB_PATH = os.environ.get('FFGP_LIB_PATH') or os.path.join(_HERE, 'libffgp.so')
_lib = None

c_dp = ctypes.c_void_p


class FFGPError(RuntimeError):
    pass


def _sig(fn, restype, argtypes):
    fn.restype = restype
    fn.argtypes = argtypes


def lib():
    """Load libffgp.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FFGPError(f'{LIB_PATH} not found: build the CUDA extension first (python -c "import __graft_entry__ as g; g.build()"). '
                        'fidelityfusion_b200 has no CPU fallback.')
    L = ctypes.CDLL(LIB_PATH)
    i, sz, vp = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p
    ll = ctypes.c_longlong
    _sig(L.ffgp_version, i, [])
    _sig(L.ffgp_last_error_string, ctypes.c_char_p, [])
    _sig(L.ffgp_launch_count, ctypes.c_ulonglong, [])
    _sig(L.ffgp_trace_dump, i, [])
    _sig(L.ffgp_debug_last_eigh_sweeps, i, [])
    _sig(L.ffgp_gemm_f64, i, [i, i, vp, i, ll, vp, i, ll, vp, i, ll, i, i, i, ctypes.c_double, ctypes.c_double, i, i, i, vp])
    _sig(L.ffgp_kernel_matrix_f64, i, [vp, vp, vp, vp, i, i, i, i, i, i, vp, vp])
    _sig(L.ffgp_kernel_matrix_bwd_scratch_bytes, sz, [i, i, i, i])
    _sig(L.ffgp_kernel_matrix_bwd_f64, i, [vp, vp, vp, vp, vp, i, i, i, i, i, vp, vp, vp, sz, vp])
    _sig(L.ffgp_kernel_matrix_bwd_x_scratch_bytes, sz, [i, i, i, i])
    _sig(L.ffgp_kernel_matrix_bwd_x_f64, i, [vp, vp, vp, vp, vp, i, i, i, i, i, vp, vp, vp, sz, vp])
    _sig(L.ffgp_dense_predict_bwd_scratch_bytes, sz, [i, i, i, i])
    _sig(L.ffgp_dense_predict_bwd_f64, i, [vp] * 6 + [i] * 7 + [vp, sz, vp, vp, sz, vp])
    _sig(L.ffgp_acquisition_f64, i, [vp, vp, i, i, ctypes.c_double, ctypes.c_double, ctypes.c_double, i, vp, vp, vp, vp])
    _sig(L.ffgp_dense_workspace_bytes, sz, [i, i, i, i, i])
    _sig(L.ffgp_dense_nll_f64, i, [vp] * 6 + [i] * 7 + [vp, sz] + [vp] * 7 + [vp, vp])
    _sig(L.ffgp_dense_predict_f64, i, [vp] * 10 + [i] * 9 + [vp, sz] + [vp, vp, vp, vp])
    _sig(L.ffgp_dense_fit_f64, i, [vp] * 10 + [i] * 11 + [vp, sz] + [vp] * 9 + [vp, vp])
    _sig(L.ffgp_adam_step_f64, i, [vp, vp, i] + [ctypes.c_double] * 4 + [i, vp, vp, i, vp])
    _sig(L.ffgp_row_match_f64, i, [vp, vp, i, i, i, vp, vp])
    _sig(L.ffgp_batched_pack_f64, i, [vp] * 10 + [i] * 7 + [ctypes.c_double, ctypes.c_double, vp, i, vp])
    _sig(L.ffgp_batched_pack_acq_f64, i, [vp] * 10 + [i] * 7 + [ctypes.c_double, ctypes.c_double, i] + [ctypes.c_double] * 3 +
         [i, vp, i, vp])
    _sig(L.ffgp_potrf_trtri_f64, i, [vp, i, i, vp, sz, vp, vp, vp, vp, vp])
    _sig(L.ffgp_mode_dot_f64, i, [vp, vp, vp, ll, i, ll, i, i, vp])
    _sig(L.ffgp_mode_gram_scratch_bytes, sz, [ll, ll, i, i])
    _sig(L.ffgp_mode_gram_f64, i, [vp, vp, vp, ll, ll, i, i, vp, sz, vp])
    _sig(L.ffgp_kron_scale_f64, i, [vp, vp, ctypes.POINTER(ctypes.c_int), i, i, i, vp, ctypes.c_double, vp, vp])
    _sig(L.ffgp_kron_ck_scratch_bytes, sz, [i])
    _sig(L.ffgp_kron_ck_f64, i, [vp, ctypes.POINTER(ctypes.c_int), i, i, vp, ctypes.c_double, vp, vp, sz, vp])
    _sig(L.ffgp_syevj_workspace_bytes, sz, [i, i])
    _sig(L.ffgp_syevj_f64, i, [vp, i, i, vp, vp, vp, sz, vp, vp])
    _sig(L.ffgp_kron_core_scratch_bytes, sz, [ll])
    _sig(L.ffgp_kron_core_f64, i, [vp, vp, ctypes.POINTER(ctypes.c_int), i, vp, ctypes.c_double, vp, vp, vp, vp, sz, vp])
    _lib = L
    return L


def ptr(t):
    """Device pointer of a contiguous CUDA fp64/int32 tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise FFGPError('fidelityfusion_b200 operates on CUDA tensors only (no CPU fallback)')
    if not t.is_contiguous():
        raise FFGPError('internal error: non-contiguous tensor passed to the C ABI')
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what):
    if rc != 0:
        msg = lib().ffgp_last_error_string().decode()
        raise FFGPError(f'{what} failed with status {rc}: {msg}')
