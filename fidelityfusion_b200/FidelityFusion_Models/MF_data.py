"""Data manager of the gen-2024 fusion models (reference FidelityFusion_Models/MF_data.py), the step either side of
the GP hot path: per-fidelity storage, z-normalisation, overlap / unique row matching and the
non-subset fill that `train_AR/CIGAR/GAR/...` call before each fidelity's optimisation loop.

Same class and method names, argument meaning, return values and ordering as the reference.  What changed is where the
work runs: the `[n1, n2, d]` broadcast compares of MF_data.py:199-202 / 236-239 are ONE launch of `ffgp_row_match_f64`
per direction (traffic = the two inputs instead of n1*n2*d bytes), and every tensor the manager creates lives on the
device of the data it was given (the reference creates `torch.zeros(...)` on the default device)."""
from __future__ import annotations

import torch

from .. import data_match

EPS = 1e-10


class Normalizer:
    """reference MF_data.py:9-73: z-normalisation of x per column (`normal_x_dim`) and of y globally
    (`normal_y_mode=0`) or per column (1); unbiased std, EPS added to the denominator only."""

    def __init__(self, x, y, normal_x_dim=0, normal_y_mode=0) -> None:
        self.x_mean = x.mean(dim=normal_x_dim)
        self.x_std = x.std(dim=normal_x_dim)
        if normal_y_mode == 0:
            self.y_mean = y.mean()
            self.y_std = y.std()
        elif normal_y_mode == 1:
            self.y_mean = y.mean(0)
            self.y_std = y.std(0)

    def normalize_x(self, x):
        return (x - self.x_mean.expand_as(x)) / (self.x_std.expand_as(x) + EPS)

    def normalize(self, x, y):
        return self.normalize_x(x), (y - self.y_mean.expand_as(y)) / (self.y_std.expand_as(y) + EPS)

    def denormalize(self, mean, var):
        return mean * self.y_std.expand_as(mean) + self.y_mean.expand_as(mean), var * (self.y_std ** 2).expand_as(var)


class MultiFidelityDataManager:
    """reference MF_data.py:76-325.  `data_dict[name] = {'fidelity_index', 'X', 'Y'}`; one Normalizer per fidelity
    index, fitted on the FIRST batch added for it (:139-140)."""

    def __init__(self, initial_data=None):
        self.data_dict = {}
        self.normalizelayer = {}
        for item in (initial_data or []):
            self.add_data(item['raw_fidelity_name'], item['fidelity_indicator'], item['X'], item['Y'])

    def add_data(self, raw_fidelity_name, fidelity_index, x, y):
        entry = self.data_dict.get(raw_fidelity_name)
        if entry is None:
            self.data_dict[raw_fidelity_name] = {'fidelity_index': fidelity_index, 'X': x, 'Y': y}
        else:
            entry['X'] = torch.cat([entry['X'], x])
            entry['Y'] = torch.cat([entry['Y'], y])
        if fidelity_index is not None and fidelity_index not in self.normalizelayer:
            self.normalizelayer[fidelity_index] = Normalizer(x, y)

    def _entry_of(self, fidelity_index):
        for entry in self.data_dict.values():
            if entry['fidelity_index'] == fidelity_index:
                return entry
        return None

    def get_data(self, fidelity_index, normal=True):
        entry = self._entry_of(fidelity_index)
        if entry is None:
            return None, None
        if normal and fidelity_index in self.normalizelayer:
            return self.normalizelayer[fidelity_index].normalize(entry['X'], entry['Y'])
        return entry['X'], entry['Y']

    def get_data_by_name(self, raw_fidelity_name, normal=True):
        entry = self.data_dict.get(raw_fidelity_name)
        if entry is None:
            return None, None
        if normal and entry['fidelity_index'] in self.normalizelayer:
            return self.normalizelayer[entry['fidelity_index']].normalize(entry['X'], entry['Y'])
        return entry['X'], entry['Y']

    def _matched(self, fidelity_index1, fidelity_index2, normal, select):
        x1, y1 = self.get_data(fidelity_index1, normal=False)
        x2, y2 = self.get_data(fidelity_index2, normal=False)
        if x1 is None or x2 is None:
            return None
        sx1, sy1, sx2, sy2 = select(x1, y1, x2, y2)
        if normal and fidelity_index1 in self.normalizelayer and fidelity_index2 in self.normalizelayer:
            sx1, sy1 = self.normalizelayer[fidelity_index1].normalize(sx1, sy1)
            sx2, sy2 = self.normalizelayer[fidelity_index2].normalize(sx2, sy2)
        return sx1, sy1, sx2, sy2

    def get_overlap_input_data(self, fidelity_index1, fidelity_index2, normal=False):
        """MF_data.py:177-213: rows of each fidelity whose input also occurs in the other one, in their own order."""
        out = self._matched(fidelity_index1, fidelity_index2, normal, data_match.get_overlap_input_data)
        if out is None:
            print("No overlap data found")
            return None, None, None, None
        return out

    def get_unique_input_data(self, fidelity_index1, fidelity_index2, normal=False):
        """MF_data.py:215-251: the complement."""
        out = self._matched(fidelity_index1, fidelity_index2, normal, data_match.get_unique_input_data)
        if out is None:
            print("No unique data found")
            return None, None, None, None
        return out

    def get_nonsubset_fill_data(self, model, fidelity_index1, fidelity_index2):
        """MF_data.py:253-303: training set of fidelity 2's residual GP when its inputs are not a subset of fidelity
        1's.  Returns (x, [y_low_mean, y_low_var], [y_high_mean, y_high_var]): shared inputs first (observed low
        outputs, zero variance), then fidelity 2's own inputs with the low output filled in by the model's posterior
        (`model.forward(self, x, to_fidelity=fidelity_index1)`, variance on the trailing diagonal block)."""
        subset_x1, subset_y1, subset_x2, subset_y2 = self.get_overlap_input_data(fidelity_index1, fidelity_index2)
        _, _, unique_x2, unique_y2 = self.get_unique_input_data(fidelity_index1, fidelity_index2)
        n1, n2 = self.normalizelayer[fidelity_index1], self.normalizelayer[fidelity_index2]
        _, subset_y1 = n1.normalize(subset_x1, subset_y1)
        subset_x2, subset_y2 = n2.normalize(subset_x2, subset_y2)
        unique_x2, unique_y2 = n2.normalize(unique_x2, unique_y2)
        ns, nu = subset_x2.shape[0], unique_x2.shape[0]
        zeros = lambda k, like: torch.zeros((k, k), dtype=like.dtype, device=like.device)

        if nu == 0:                                              # fidelity 2's inputs are a subset
            return subset_x2, [subset_y1, zeros(ns, subset_y1)], [subset_y2, zeros(ns, subset_y2)]

        fill_mean, fill_var = model.forward(self, unique_x2, to_fidelity=fidelity_index1)
        if fill_var.shape[0] != fill_var.shape[1]:               # HOGP returns the diagonal only (:277, :292)
            fill_var = torch.diag_embed(fill_var.squeeze())
        if ns == 0:                                              # nothing shared
            return unique_x2, [fill_mean.reshape(-1, 1), fill_var], [unique_y2, zeros(nu, unique_y2)]

        y_low_mean = torch.cat([subset_y1, fill_mean.reshape(-1, 1)], dim=0)
        if fill_mean.dim() == 0:
            fill_mean = fill_mean.reshape(1)
        k = ns + fill_mean.shape[0]
        y_low_var = zeros(k, y_low_mean)
        y_low_var[-fill_var.shape[0]:, -fill_var.shape[1]:] = fill_var
        y_high_mean = torch.cat([subset_y2, unique_y2], dim=0)
        x = torch.cat([subset_x2, unique_x2], dim=0)
        return x, [y_low_mean, y_low_var], [y_high_mean, zeros(ns + nu, y_high_mean)]

    def display_fidelity_data_info(self, fidelity_index):
        for name, entry in self.data_dict.items():
            if entry['fidelity_index'] == fidelity_index:
                print("<---------Fidelity data information:--------->")
                print("Fidelity index: {}".format(fidelity_index))
                print("Fidelity name: {}".format(name))
                print("data_num: {}".format(entry['X'].shape[0]))
                print("X_shape: {}".format(entry['X'].shape))
                print("Y_shape: {}".format(entry['Y'].shape))
        else:
            print("No fidelity data found")
