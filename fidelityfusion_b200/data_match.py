"""Subset / overlap matching between fidelities, the step on the data side of the GP hot path.

gen-2024: `MultiFidelityDataManager.get_overlap_input_data / get_unique_input_data`
(FidelityFusion_Models/MF_data.py:177-251) select the rows of x1 that also occur in x2 (and vice versa) with a
`[n1, n2, d]` broadcast compare; gen-2023: `Subset_checker.get_subset` (MFGP_ver2023May/utils/subset_tools.py:58-90)
does the same through `torch.unique`.  Both are exact floating-point row equality.  Here one CUDA kernel
(`ffgp_row_match_f64`) returns, for every row of a, the index of its first equal row in b; masks, ordered index pairs
and the selected data follow from that with O(n) work.  Same return values and ordering as the reference functions."""
from __future__ import annotations

import torch

from . import _lib as B


def row_match(a, b):
    """int32 [na]: index of the first row of b equal to a[i] (all columns, IEEE `==`), -1 if none.  a, b: [n, ...] CUDA
    tensors of the same trailing shape (trailing dims are flattened, like comparing whole samples)."""
    if a.shape[1:] != b.shape[1:]:
        raise ValueError('row_match: samples of different shape')
    import math
    d = math.prod(a.shape[1:])
    a2 = a.detach().reshape(a.shape[0], d).to(torch.float64).contiguous()
    b2 = b.detach().reshape(b.shape[0], d).to(torch.float64).contiguous()
    out = torch.empty(a2.shape[0], dtype=torch.int32, device=a.device)
    if a2.shape[0] == 0:
        return out
    if b2.shape[0] == 0:
        return out.fill_(-1)
    rc = B.lib().ffgp_row_match_f64(B.ptr(a2), B.ptr(b2), a2.shape[0], b2.shape[0], a2.shape[1], B.ptr(out), B.stream_ptr())
    B.check(rc, 'ffgp_row_match_f64')
    return out


def overlap_masks(x1, x2):
    """(mask1 [n1], mask2 [n2]): rows of x1 that occur in x2 and rows of x2 that occur in x1 (MF_data.py:199-202)."""
    return row_match(x1, x2) >= 0, row_match(x2, x1) >= 0


def get_overlap_input_data(x1, y1, x2, y2):
    """MF_data.py:177-213 without the normalisation layer: (common_x1, y1[mask1], common_x2, y2[mask2]), each side in
    its own original order."""
    m1, m2 = overlap_masks(x1, x2)
    return x1[m1], y1[m1], x2[m2], y2[m2]


def get_unique_input_data(x1, y1, x2, y2):
    """MF_data.py:215-251: the rows that do NOT occur on the other side."""
    m1, m2 = overlap_masks(x1, x2)
    return x1[~m1], y1[~m1], x2[~m2], y2[~m2]


def unique_check(t):
    """subset_tools.py:49-53: duplicate samples are an error."""
    if t.shape[0] > 1:
        first = row_match(t, t)
        if bool((first != torch.arange(t.shape[0], device=t.device, dtype=torch.int32)).any()):
            raise AssertionError('|error|: tensor has duplicate samples')


def get_subset(data_a, data_b, subset_type='index'):
    """subset_tools.py:58-90.  'mask': 0/1 masks over the samples of a and b; 'index': index pairs (ia, ib) with
    data_a[ia] == data_b[ib], ordered like the reference (by the lexicographic order of the shared samples, which is
    the order torch.unique(dim=0) lists them in)."""
    assert subset_type in ['index', 'mask'], "|error|: subset_type should be 'index' or 'mask'"
    unique_check(data_a)
    unique_check(data_b)
    ma = row_match(data_a, data_b)
    if subset_type == 'mask':
        mb = row_match(data_b, data_a)
        return (ma >= 0).long(), (mb >= 0).long()
    ia = (ma >= 0).nonzero().reshape(-1)
    if ia.numel() > 1:
        rows = data_a.detach().reshape(data_a.shape[0], -1)[ia]
        order = torch.arange(ia.numel(), device=ia.device)
        for k in range(rows.shape[1] - 1, -1, -1):             # stable sorts, last column first = lexicographic
            order = order[torch.sort(rows[order, k], stable=True).indices]
        ia = ia[order]
    return ia, ma[ia].long()
