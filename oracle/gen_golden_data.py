"""Golden vectors for the data-side matching, produced by the UNMODIFIED reference:
FidelityFusion_Models/MF_data.py (MultiFidelityDataManager.get_overlap_input_data / get_unique_input_data) and
MFGP_ver2023May/utils/subset_tools.py (Subset_checker.get_subset).  TEST INFRASTRUCTURE; run once in the build
container:   python oracle/gen_golden_data.py   ->  tests/golden/data_match.npz"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'data_match.npz')


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


mfd = load('FidelityFusion_Models/MF_data.py', 'ref_mf_data')
sst = load('MFGP_ver2023May/utils/subset_tools.py', 'ref_subset_tools')
torch.set_default_dtype(torch.float64)
g = torch.Generator().manual_seed(84)
pool = torch.rand(300, 3, generator=g)
pool[7, 1] = 0.0
i1 = torch.randperm(300, generator=g)[:170]
i2 = torch.cat([i1[torch.randperm(170, generator=g)[:60]], torch.randperm(300, generator=g)[:90]]).unique()
i2 = i2[torch.randperm(i2.numel(), generator=g)]
x1, x2 = pool[i1].clone(), pool[i2].clone()
# IEEE corner cases of `==`: -0.0 equals 0.0; a NaN row equals nothing (not even its copy on the other side)
x1 = torch.cat([x1, torch.tensor([[0.25, -0.0, 0.5], [float('nan'), 1.0, 2.0]])])
x2 = torch.cat([x2, torch.tensor([[0.25, 0.0, 0.5], [float('nan'), 1.0, 2.0]])])
y1 = torch.randn(x1.shape[0], 2, generator=g)
y2 = torch.randn(x2.shape[0], 2, generator=g)
dm = mfd.MultiFidelityDataManager([
    {'raw_fidelity_name': '0', 'fidelity_indicator': 0, 'X': x1, 'Y': y1},
    {'raw_fidelity_name': '1', 'fidelity_indicator': 1, 'X': x2, 'Y': y2}])
ov = dm.get_overlap_input_data(0, 1, normal=False)
un = dm.get_unique_input_data(0, 1, normal=False)
# gen-2023 checker on finite, duplicate-free samples (its unique() pass asserts on duplicates)
a, b = pool[i1][:120].clone(), pool[i2][:100].clone()
ia, ib = sst.Subset_checker.get_subset(a, b, subset_type='index')
ma, mb = sst.Subset_checker.get_subset(a, b, subset_type='mask')
np.savez_compressed(OUT, x1=x1.numpy(), y1=y1.numpy(), x2=x2.numpy(), y2=y2.numpy(),
                    ov_x1=ov[0].numpy(), ov_y1=ov[1].numpy(), ov_x2=ov[2].numpy(), ov_y2=ov[3].numpy(),
                    un_x1=un[0].numpy(), un_y1=un[1].numpy(), un_x2=un[2].numpy(), un_y2=un[3].numpy(),
                    a=a.numpy(), b=b.numpy(), ia=ia.numpy(), ib=ib.numpy(), ma=ma.numpy(), mb=mb.numpy())
print('wrote', OUT, 'overlap', tuple(ov[0].shape), 'unique', tuple(un[0].shape), tuple(un[2].shape), 'subset pairs', ia.numel())
