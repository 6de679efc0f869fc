"""World-size-2 gloo run (CPU) of the batched-GP sharding: contiguous block partition, one packed all-gather,
identical full result on every rank.  The GPU compute is replaced by the CPU oracle through `compute_fn`
(test-only injection); on the B200 box the same code path runs with NCCL and the CUDA compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _oracle_compute(x, y, ls, sv, lb, xs, want_grad, check=True, acq=None):
    """Stands in for batched_cigp_eval.  check=False adds the status column the CUDA path returns (problem b of the
    GLOBAL batch is flagged when its log_beta is NaN, like the device-side factorisation would)."""
    from oracle import ff_oracle as O
    if not check:
        bad = torch.isnan(lb)
        res = _oracle_compute(x, y, ls, sv, torch.where(bad, torch.zeros_like(lb), lb), xs, want_grad, acq=acq)
        res['info'] = bad.double() * 7.0
        return res
    out = {k: [] for k in ('nll', 'g_length_scales', 'g_signal_variance', 'g_log_beta', 'mean', 'var')}
    for b in range(x.shape[0]):
        loss, gr = O.cigp_ard_nll_and_grads(x[b], y[b], ls[b], sv[b:b + 1], lb[b:b + 1])
        m, c = O.cigp_ard_predict(x[b], y[b], xs[b], ls[b], sv[b:b + 1], lb[b:b + 1])
        out['nll'].append(torch.tensor(loss)); out['g_length_scales'].append(gr['length_scales'])
        out['g_signal_variance'].append(gr['signal_variance'][0]); out['g_log_beta'].append(gr['log_beta'][0])
        out['mean'].append(m); out['var'].append(c.diag())
    res = {k: torch.stack(v) for k, v in out.items()}
    if acq is not None:                                # the score block of the packed row (ffgp_batched_pack_acq_f64)
        res['score'] = O.acq_sf_score('EI', res['mean'][..., 0], res['var'], f_best=acq['f_best'], xi=acq['xi'])
    return res


def _problems(Bn):
    gen = torch.Generator().manual_seed(77)
    return (torch.rand(Bn, 12, 3, generator=gen, dtype=torch.float64), torch.randn(Bn, 12, 1, generator=gen, dtype=torch.float64),
            torch.rand(Bn, 3, generator=gen, dtype=torch.float64) + 0.5, torch.ones(Bn, dtype=torch.float64),
            torch.rand(Bn, generator=gen, dtype=torch.float64), torch.rand(Bn, 4, 3, generator=gen, dtype=torch.float64))


def _worker(rank, world, port, Bn, q, check=True, acq=None):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from fidelityfusion_b200.batched import sharded_cigp_eval
        torch.set_default_dtype(torch.float64)
        x, y, ls, sv, lb, xs = _problems(Bn)
        if not check:
            lb[Bn - 2] = float('nan')                      # lives in the LAST rank's block
        res = sharded_cigp_eval(x, y, ls, sv, lb, xs, compute_fn=_oracle_compute, check=check, acq=acq)
        q.put((rank, {k: v.numpy().copy() for k, v in res.items()}))     # plain arrays: no shared-memory handles
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('Bn', [6, 7, 1])          # 1: rank 1's block is empty and must still join the all-gather
def test_sharded_eval_two_ranks_gloo(Bn):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, Bn, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.set_default_dtype(torch.float64)
    full = _oracle_compute(*_problems(Bn), True)
    for r in (0, 1):
        for k, v in full.items():
            assert tuple(got[r][k].shape) == tuple(v.shape), k
            assert (got[r][k] == v.numpy()).all(), (r, k)                   # every rank holds the identical full result


def test_sharded_eval_asynchronous_status_is_gathered():
    """check=False: the per-problem status word travels in the packed all-gather, so every rank sees a failure that
    happened in another rank's block and check_batch_info names the global batch index."""
    from fidelityfusion_b200.batched import check_batch_info
    Bn = 7
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, Bn, q, False)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        info = torch.from_numpy(got[r]['info'])
        assert info.shape == (Bn,) and float(info[Bn - 2]) == 7.0 and float(info.sum()) == 7.0
        with pytest.raises(torch.linalg.LinAlgError, match=f'Batch element {Bn - 2}'):
            check_batch_info(info)
        assert got[r]['nll'].shape == (Bn,)


@pytest.mark.parametrize('Bn', [5, 1])
def test_sharded_acquisition_scores_travel_in_the_single_all_gather(Bn):
    """acq=...: the score block is one more field of the packed row, gathered with the predictions it is computed from
    (also when a rank's block is empty)."""
    acq = dict(kind='EI', f_best=0.1, xi=0.01)
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, Bn, q, True, acq)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.set_default_dtype(torch.float64)
    full = _oracle_compute(*_problems(Bn), True, acq=acq)
    for r in (0, 1):
        assert tuple(got[r]['score'].shape) == (Bn, 4)
        for k, v in full.items():
            assert (got[r][k] == v.numpy()).all(), (r, k)
