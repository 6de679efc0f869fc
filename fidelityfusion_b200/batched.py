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


# score kinds of ffgp_acquisition_f64 / ffgp_batched_pack_acq_f64 (one vocabulary for the whole package):
#   'UCB', 'EI', 'PI'      DiscreteAcquisitionFunction.UCB_MF / EI_MF / PI_MF (MF_BayesianOptimization/Discrete/DMF_acq.py:49-128)
#   'UCB_STD', 'PI_CDF'    UCB / PI of the single-fidelity module (Bayesian_optimization/acq.py:135-231); its EI is 'EI'
ACQ_KINDS = {'UCB': 0, 'EI': 1, 'PI': 2, 'UCB_STD': 3, 'PI_CDF': 4}


def batched_cigp_eval(x, y, length_scales, signal_variance, log_beta, xs=None, want_grad=True, check=True, acq=None):
    """x [B,n,d], y [B,n,D], length_scales [B,d], signal_variance [B], log_beta [B] (raw reference parameters),
    xs [B,ns,d] or None.  Returns dict: nll [B] (= -cigp.negative_log_likelihood), g_length_scales [B,d],
    g_signal_variance [B], g_log_beta [B], and mean [B,ns,D], var [B,ns] (diag of cigp.forward's covariance).

    ONE C call (ffgp_dense_fit_f64): every problem is factorised once and the NLL, its analytic gradient and the
    posterior all come from that factor; the chain rule to the reference's raw parameters is a few vector ops.

    check=True raises torch.linalg.LinAlgError at the call when a covariance is not positive definite, like the
    reference's torch.linalg.cholesky - which costs a host synchronisation per call.  check=False keeps the call
    asynchronous (sweeps can be enqueued back to back) and returns the LAPACK-style status as out['info'] (fp64 [B],
    0 = ok, k = leading minor k not PD) for the caller to inspect, e.g. once per optimisation step: `check_batch_info`.

    acq = dict(kind=<a key of ACQ_KINDS>, f_best=, beta=, xi=, round_f32=True) (needs xs, D = 1) adds
    out['score'] [B,ns]: the acquisition score of every test point, computed in the epilogue that writes the result rows
    (ffgp_batched_pack_acq_f64) - no extra launch, and it travels through sharded_cigp_eval's single all-gather."""
    from . import _lib as B
    L = B.lib()
    Bn, n, d = x.shape
    D = y.shape[2]
    dev = x.device
    if not x.is_cuda:
        raise B.FFGPError('fidelityfusion_b200 operates on CUDA tensors only (no CPU fallback)')
    if Bn == 0:
        return empty_result(d, D, 0 if xs is None else xs.shape[1], want_grad, check, dev, acq is not None)
    f64 = lambda t: t.detach().to(torch.float64).contiguous()
    xc, yc = f64(x), f64(y)
    ls, sv, lb = f64(length_scales), f64(signal_variance).reshape(Bn), f64(log_beta).reshape(Bn)
    ell = ls.abs() + EPS
    inv_ls = (1.0 / ell).contiguous()
    amp = sv.abs().contiguous()
    noise = torch.exp(-lb)
    diag = (noise + JITTER).unsqueeze(1).expand(Bn, n).contiguous()
    ns = 0 if xs is None else xs.shape[1]
    xsc = f64(xs) if xs is not None else None
    nll = torch.empty(Bn, dtype=torch.float64, device=dev)
    alpha = torch.empty(Bn, n, D, dtype=torch.float64, device=dev)
    g_il = torch.empty(Bn, d, dtype=torch.float64, device=dev) if want_grad else None
    g_amp = torch.empty(Bn, dtype=torch.float64, device=dev) if want_grad else None
    g_diag = torch.empty(Bn, n, dtype=torch.float64, device=dev) if want_grad else None
    mean = torch.empty(Bn, ns, D, dtype=torch.float64, device=dev) if ns else None
    var = torch.empty(Bn, ns, dtype=torch.float64, device=dev) if ns else None
    info = torch.empty(Bn, dtype=torch.int32, device=dev)
    wsb = L.ffgp_dense_workspace_bytes(n, d, D, ns, Bn)
    ws = ops._ws_cache.get(wsb, dev)
    noise_c = noise.contiguous()
    rc = L.ffgp_dense_fit_f64(B.ptr(xc), B.ptr(yc), B.ptr(xsc), B.ptr(inv_ls), B.ptr(amp), B.ptr(diag), None, None, None,
                              B.ptr(noise_c), n, d, D, ns, Bn, 1, 1, 1, int(want_grad), 0, 0, B.ptr(ws), wsb,
                              B.ptr(nll), None, B.ptr(alpha), B.ptr(g_il), B.ptr(g_amp), B.ptr(g_diag), None,
                              B.ptr(mean), B.ptr(var), B.ptr(info), B.stream_ptr())
    B.check(rc, 'ffgp_dense_fit_f64')
    if check:
        ops.check_info(info)
    # ONE launch writes the packed result rows: NLL constant, chain rule to the raw parameters, predictions and (for the
    # asynchronous mode) the status column.  The dict entries below are views of that buffer; sharded_cigp_eval ships
    # the buffer itself through its single all-gather.
    if acq is not None and (ns == 0 or D != 1):
        raise ValueError('acq needs test points xs and a single output column (D = 1)')
    keys, widths = result_layout(d, D, ns, want_grad, check, acq is not None)
    ld = sum(widths)
    packed = torch.empty(Bn, ld, dtype=torch.float64, device=dev)
    a = acq or {}
    kind = a.get('kind', 'EI')
    rc = L.ffgp_batched_pack_acq_f64(B.ptr(nll), B.ptr(g_il), B.ptr(g_amp), B.ptr(g_diag), B.ptr(mean), B.ptr(var), B.ptr(info),
                                     B.ptr(ls), B.ptr(sv), B.ptr(lb), Bn, n, d, D, ns, int(want_grad), int(not check),
                                     0.5 * n * D * math.log(2 * PI), EPS,
                                     -1 if acq is None else (ACQ_KINDS[kind] if isinstance(kind, str) else int(kind)),
                                     float(a.get('f_best', 0.0)), float(a.get('beta', 1.0)), float(a.get('xi', 0.01)),
                                     int(bool(a.get('round_f32', True))), B.ptr(packed), ld, B.stream_ptr())
    B.check(rc, 'ffgp_batched_pack_acq_f64')
    out = unpack_results(packed, result_shapes(d, D, ns), keys)
    out['_packed'] = packed
    return out


def result_layout(d, D, ns, want_grad, check, with_score=False):
    """(keys, column widths) of a packed result row, in the order ffgp_batched_pack_acq_f64 writes them."""
    keys, widths = ['nll'], [1]
    if want_grad:
        keys += ['g_length_scales', 'g_signal_variance', 'g_log_beta']
        widths += [d, 1, 1]
    if ns:
        keys += ['mean', 'var']
        widths += [ns * D, ns]
        if with_score:
            keys.append('score')
            widths.append(ns)
    if not check:
        keys.append('info')
        widths.append(1)
    return keys, widths


def result_shapes(d, D, ns):
    return {'nll': (), 'g_length_scales': (d,), 'g_signal_variance': (), 'g_log_beta': (), 'mean': (ns, D), 'var': (ns,),
            'score': (ns,), 'info': ()}


def empty_result(d, D, ns, want_grad, check, device, with_score=False):
    """The result dict of a batch of ZERO problems (a rank whose block is empty when there are fewer problems than
    ranks): nothing to compute, but the rank still joins the collective with correctly shaped fields."""
    z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=device)
    out = {'nll': z(0)}
    if not check:
        out['info'] = z(0)
    if want_grad:
        out.update(g_length_scales=z(0, d), g_signal_variance=z(0), g_log_beta=z(0))
    if ns:
        out.update(mean=z(0, ns, D), var=z(0, ns))
        if with_score:
            out['score'] = z(0, ns)
    return out


def check_batch_info(info):
    """Raise what the reference raises for the first problem whose covariance was not positive definite
    (`info` = out['info'] of a check=False evaluation, possibly all-gathered over the ranks).  One host sync."""
    ops.check_info(info.round().to(torch.int32))


def shard_range(total, rank, world):
    """Contiguous block partition [lo, hi) of `total` problems for `rank` of `world` (remainder to the low ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(res, keys):
    """[B_local, F] buffer: one row per problem, fields in `keys` order, each flattened."""
    return torch.cat([res[k].reshape(res[k].shape[0], math.prod(res[k].shape[1:])) for k in keys], dim=1).contiguous()


def unpack_results(buf, shapes, keys):
    out, o = {}, 0
    for k in keys:
        w = math.prod(shapes[k])
        out[k] = buf[:, o:o + w].reshape((buf.shape[0],) + tuple(shapes[k]))
        o += w
    return out


def sharded_cigp_eval(x, y, length_scales, signal_variance, log_beta, xs=None, want_grad=True, group=None,
                      compute_fn=batched_cigp_eval, check=True, acq=None):
    """Every rank passes the FULL problem set (or at least its own block - only [lo,hi) is read) and receives the
    full result set.  Falls back to a single-rank call when torch.distributed is not initialised.
    `compute_fn` exists so the CPU (gloo) tests can exercise the partition/gather logic without a GPU."""
    import torch.distributed as dist
    kw = {} if check else {'check': False}
    if acq is not None:
        kw['acq'] = acq
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return compute_fn(x, y, length_scales, signal_variance, log_beta, xs, want_grad, **kw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Bn = x.shape[0]
    lo, hi = shard_range(Bn, rank, world)
    sl = slice(lo, hi)
    if hi > lo:
        res = compute_fn(x[sl], y[sl], length_scales[sl], signal_variance[sl], log_beta[sl],
                         None if xs is None else xs[sl], want_grad, **kw)
    else:           # fewer problems than ranks: skip the compute, still join the collective (the others would hang)
        res = empty_result(x.shape[2], y.shape[2], 0 if xs is None else xs.shape[1], want_grad, check, x.device, acq is not None)
    keys = [k for k in ('nll', 'g_length_scales', 'g_signal_variance', 'g_log_beta', 'mean', 'var', 'score', 'info') if k in res]
    shapes = {k: tuple(res[k].shape[1:]) for k in keys}
    local = res['_packed'] if '_packed' in res else pack_results(res, keys)      # the CUDA path packs in its own kernel
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
