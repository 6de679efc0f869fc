"""Subset / overlap matching.  CPU: the oracle restatement against the golden vectors produced by the
unmodified reference (oracle/gen_golden_data.py).  GPU: ffgp_row_match_f64 through fidelityfusion_b200.data_match
against the same vectors, against the oracle on seeded inputs, and size-independent properties at N = 20000."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'data_match.npz')


def gold():
    return {k: torch.from_numpy(v) for k, v in np.load(GOLD).items()}


def same(a, b):
    """Bitwise equality that treats NaN == NaN (rows are copied, never computed)."""
    return a.shape == b.shape and bool(((a == b) | (a.isnan() & b.isnan())).all())


def test_oracle_masks_reproduce_the_reference_selection():
    from oracle import ff_oracle as O
    g = gold()
    m1, m2 = O.overlap_masks(g['x1'], g['x2'])
    assert same(g['x1'][m1], g['ov_x1']) and same(g['y1'][m1], g['ov_y1'])
    assert same(g['x2'][m2], g['ov_x2']) and same(g['y2'][m2], g['ov_y2'])
    assert same(g['x1'][~m1], g['un_x1']) and same(g['x2'][~m2], g['un_x2'])
    # IEEE corner cases stored in the fixture: the -0.0 / 0.0 pair matches, the NaN rows match nothing
    assert bool(m1[-2]) and bool(m2[-2]) and not bool(m1[-1]) and not bool(m2[-1])
    ia, ib = O.get_subset_index(g['a'], g['b'])
    assert torch.equal(ia, g['ia']) and torch.equal(ib, g['ib'])
    assert torch.equal(g['a'][ia], g['b'][ib])


@pytest.mark.gpu
def test_gpu_overlap_and_unique_match_the_reference():
    from fidelityfusion_b200 import data_match as M
    g = gold()
    x1, y1, x2, y2 = (g[k].cuda() for k in ('x1', 'y1', 'x2', 'y2'))
    ov = M.get_overlap_input_data(x1, y1, x2, y2)
    un = M.get_unique_input_data(x1, y1, x2, y2)
    for got, key in zip(ov, ('ov_x1', 'ov_y1', 'ov_x2', 'ov_y2')):
        assert same(got.cpu(), g[key]), key
    for got, key in zip(un, ('un_x1', 'un_y1', 'un_x2', 'un_y2')):
        assert same(got.cpu(), g[key]), key


@pytest.mark.gpu
def test_gpu_get_subset_matches_the_gen2023_checker():
    from fidelityfusion_b200 import data_match as M
    g = gold()
    a, b = g['a'].cuda(), g['b'].cuda()
    ia, ib = M.get_subset(a, b, 'index')
    assert torch.equal(ia.cpu(), g['ia']) and torch.equal(ib.cpu(), g['ib'])
    ma, mb = M.get_subset(a, b, 'mask')
    assert torch.equal(ma.cpu(), g['ma']) and torch.equal(mb.cpu(), g['mb'])
    with pytest.raises(AssertionError):
        M.get_subset(torch.cat([a, a[:1]]), b)                     # duplicate samples are an error (subset_tools.py:49-53)


@pytest.mark.gpu
@pytest.mark.parametrize('n1,n2,d', [(1, 1, 1), (129, 257, 5), (300, 7, 64), (1000, 999, 2)])
def test_gpu_row_match_against_oracle_on_ragged_sizes(n1, n2, d):
    from fidelityfusion_b200 import data_match as M
    from oracle import ff_oracle as O
    g = torch.Generator().manual_seed(n1 * 7 + n2)
    pool = torch.randint(0, 3, (n1 + n2, d), generator=g).double() if d <= 2 else torch.rand(n1 + n2, d, generator=g, dtype=torch.float64)
    x1 = pool[torch.randint(0, n1 + n2, (n1,), generator=g)]
    x2 = pool[torch.randint(0, n1 + n2, (n2,), generator=g)]
    m1, m2 = O.overlap_masks(x1, x2)
    g1, g2 = M.overlap_masks(x1.cuda(), x2.cuda())
    assert torch.equal(g1.cpu(), m1) and torch.equal(g2.cpu(), m2)
    first = M.row_match(x1.cuda(), x2.cuda()).cpu().long()
    eq = torch.all(x1.unsqueeze(1) == x2.unsqueeze(0), dim=-1)
    want = torch.where(eq.any(1), eq.float().argmax(1), torch.full((n1,), -1))
    assert torch.equal(first, want)                                # the FIRST equal row
    assert M.row_match(x1[:0].cuda(), x2.cuda()).numel() == 0 and bool((M.row_match(x1.cuda(), x2[:0].cuda()) == -1).all())


@pytest.mark.gpu
def test_gpu_row_match_full_size_properties():
    """N = 20000 per fidelity (where the reference's [n1, n2, d] boolean tensor would take 2 GB): a permuted subset is
    found exactly, the match is an involution on the shared rows, and nothing outside the subset matches."""
    from fidelityfusion_b200 import data_match as M
    g = torch.Generator().manual_seed(9)
    n, d, k = 20000, 5, 12345
    x1 = torch.rand(n, d, generator=g, dtype=torch.float64).cuda()
    perm = torch.randperm(n, generator=g)[:k].cuda()
    x2 = torch.cat([x1[perm], torch.rand(n - k, d, generator=g, dtype=torch.float64).cuda() + 2.0])
    m12, m21 = M.row_match(x1, x2).long(), M.row_match(x2, x1).long()
    assert int((m12 >= 0).sum()) == k and int((m21 >= 0).sum()) == k
    assert torch.equal(m21[:k], perm)
    shared = (m12 >= 0).nonzero().reshape(-1)
    assert torch.equal(m21[m12[shared]], shared)
    assert torch.equal(x1[shared], x2[m12[shared]])
