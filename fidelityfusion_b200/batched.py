"""Batched independent GPs (hyper-parameter restarts, per-candidate re-fits of the MF-BO acquisition sweeps,
reference MF_BayesianOptimization/.../v1/CFKG.py:124-129, per-fidelity models CIGAR.py:28-31) and their
sharding across the GPUs of one box.

Each problem b is the reference's `cigp(ARDKernel(d), log_beta_b)` on (x_b, y_b): the unit of work is one
NLL + gradient evaluation (and optionally the posterior at xs_b).  The reference runs these sequentially in
Python; here one C call factorises the whole batch.

Multi-GPU: problems are independent, so rank r of R owns the contiguous block
[r*B/R, (r+1)*B/R) and there is no data-path collective; after the last kernel ONE all-gather of a packed
result buffer [B_local, 1 + (d+2) + ns*(D+1)] returns every problem's NLL / gradients / predictions to all ranks.
"""
from __future__ import annotations

import math

import torch

from . import ops

JITTER = 1e-6
PI = 3.1415
EPS = 1e-9


def batched_cigp_eval(x, y, length_scales, signal_variance, log_beta, xs=None, want_grad=True):
    """x [B,n,d], y [B,n,D], length_scales [B,d], signal_variance [B], log_beta [B] (raw reference parameters),
    xs [B,ns,d] or None.  Returns dict: nll [B] (= -cigp.negative_log_likelihood), g_length_scales [B,d],
    g_signal_variance [B], g_log_beta [B], and mean [B,ns,D], var [B,ns] (diag of cigp.forward's covariance)."""
    Bn, n, d = x.shape
    D = y.shape[2]
    ls = length_scales.detach().clone().requires_grad_(want_grad)
    sv = signal_variance.detach().clone().requires_grad_(want_grad)
    lb = log_beta.detach().clone().requires_grad_(want_grad)
    inv_ls = 1.0 / (ls.abs() + EPS)
    amp = sv.abs()
    noise = torch.exp(-lb)
    diag = (noise + JITTER).unsqueeze(1).expand(Bn, n)
    core = ops.dense_nll(x, y, inv_ls, amp, diag_add=diag, clamp=True)
    nll = core + 0.5 * n * D * math.log(2 * PI)
    out = {'nll': nll.detach()}
    if want_grad:
        nll.sum().backward()
        out.update(g_length_scales=ls.grad, g_signal_variance=sv.grad, g_log_beta=lb.grad)
    if xs is not None:
        with torch.no_grad():
            mean, var = ops.dense_predict(x, y, xs, inv_ls.detach(), amp.detach(), diag_add=diag.detach(),
                                          cov_offset=noise.detach(), full_cov=False, clamp=True)
        out.update(mean=mean, var=var)
    return out


def shard_range(total, rank, world):
    """Contiguous block partition [lo, hi) of `total` problems for `rank` of `world` (remainder to the low ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(res, keys):
    """[B_local, F] buffer: one row per problem, fields in `keys` order, each flattened."""
    return torch.cat([res[k].reshape(res[k].shape[0], -1) for k in keys], dim=1).contiguous()


def unpack_results(buf, shapes, keys):
    out, o = {}, 0
    for k in keys:
        w = math.prod(shapes[k])
        out[k] = buf[:, o:o + w].reshape((buf.shape[0],) + tuple(shapes[k]))
        o += w
    return out


def sharded_cigp_eval(x, y, length_scales, signal_variance, log_beta, xs=None, want_grad=True, group=None,
                      compute_fn=batched_cigp_eval):
    """Every rank passes the FULL problem set (or at least its own block - only [lo,hi) is read) and receives the
    full result set.  Falls back to a single-rank call when torch.distributed is not initialised.
    `compute_fn` exists so the CPU (gloo) tests can exercise the partition/gather logic without a GPU."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return compute_fn(x, y, length_scales, signal_variance, log_beta, xs, want_grad)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Bn = x.shape[0]
    lo, hi = shard_range(Bn, rank, world)
    sl = slice(lo, hi)
    res = compute_fn(x[sl], y[sl], length_scales[sl], signal_variance[sl], log_beta[sl],
                     None if xs is None else xs[sl], want_grad)
    keys = [k for k in ('nll', 'g_length_scales', 'g_signal_variance', 'g_log_beta', 'mean', 'var') if k in res]
    shapes = {k: tuple(res[k].shape[1:]) for k in keys}
    local = pack_results(res, keys)
    counts = [shard_range(Bn, r, world) for r in range(world)]
    cmax = max(c[1] - c[0] for c in counts)
    if local.shape[0] < cmax:            # ragged split: pad to the largest block so ONE fixed-size collective suffices
        pad = torch.zeros((cmax - local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    gathered = torch.empty((world * cmax, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)
    if all(c[1] - c[0] == cmax for c in counts):
        full = gathered
    else:
        full = torch.cat([gathered[r * cmax: r * cmax + (c[1] - c[0])] for r, c in enumerate(counts)], 0)
    return unpack_results(full, shapes, keys)
