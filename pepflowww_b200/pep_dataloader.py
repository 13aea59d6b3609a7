"""Batch schema of the reference's data path, fed from synthetic complexes.

The reference's PepDataset (models_con/pep_dataloader.py:87-196) reads an LMDB of pickled
dicts built from PDB files with Biopython; neither the data nor those libraries exist offline,
so `SyntheticPepDataset` emits items with the SAME schema (SURVEY.md App. C /
pep_dataloader.py:41-66: pocket residues first, peptide last, peptide-CA centroid at the origin,
chain_nb 1 for the pocket / 0 for the peptide) and `PaddingCollate` restates
pepflow/utils/data.py:19-78 (pad to the batch max length, `eight` rounds L up to a multiple of 8,
adds `res_mask`).
"""
import math

import numpy as np
import torch

from .constants import PAD_RESIDUE_INDEX, restype_to_heavyatom_masks, torsions_mask


def _unit(rng, n):
    v = rng.standard_normal((n, 3))
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_synthetic_complex(index, len_pocket=128, len_peptide=12, seed=0):
    """One peptide-pocket complex (SURVEY.md section 8d generator). Returns a dict of tensors."""
    rng = np.random.Generator(np.random.PCG64([seed, index]))
    lr, lp = len_pocket, len_peptide
    L = lr + lp
    aa = rng.integers(0, 20, size=L)
    # peptide CA trace: 3.8 A random walk; pocket CA: 5-12 A shell around random peptide residues
    steps = _unit(rng, lp) * 3.8
    ca_pep = np.cumsum(steps, axis=0)
    anchor = rng.integers(0, lp, size=lr)
    ca_rec = ca_pep[anchor] + _unit(rng, lr) * rng.uniform(5.0, 12.0, size=(lr, 1))
    ca = np.concatenate([ca_rec, ca_pep], axis=0)
    ca = ca - ca_pep.mean(axis=0, keepdims=True)
    pos = ca[:, None, :] + rng.normal(0.0, 1.5, size=(L, 15, 3))
    u1 = _unit(rng, L)
    u2 = _unit(rng, L)
    u2 = u2 - 0.5 * (u1 * u2).sum(-1, keepdims=True) * u1   # keep N, C non-collinear
    u2 = u2 / np.linalg.norm(u2, axis=-1, keepdims=True)
    pos[:, 0] = ca + 1.46 * u1   # N
    pos[:, 1] = ca               # CA
    pos[:, 2] = ca + 1.52 * u2   # C
    aa_t = torch.from_numpy(aa).long()
    mask = restype_to_heavyatom_masks[aa_t]
    pos_t = torch.from_numpy(pos).float() * mask[..., None]
    tmask = torsions_mask[aa_t]
    tors = torch.from_numpy(rng.uniform(0.0, 2 * math.pi, size=(L, 5))).float() * tmask
    return {
        "id": f"synth_{seed}_{index}",
        "aa": aa_t,
        "pos_heavyatom": pos_t,
        "mask_heavyatom": mask.clone(),
        "res_nb": torch.cat([torch.arange(1, lr + 1), torch.arange(1, lp + 1)]).long(),
        "chain_nb": torch.cat([torch.ones(lr), torch.zeros(lp)]).long(),
        "generate_mask": torch.cat([torch.zeros(lr), torch.ones(lp)]).bool(),
        "torsion_angle": tors,
        "torsion_angle_mask": tmask.bool(),
    }


class SyntheticPepDataset(torch.utils.data.Dataset):
    """Same item schema as PepDataset.__getitem__ (pep_dataloader.py:190-196)."""

    def __init__(self, num_complexes=64, len_pocket=128, len_peptide=12, seed=0, transform=None):
        self.n, self.lr, self.lp, self.seed, self.transform = num_complexes, len_pocket, len_peptide, seed, transform

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        if not 0 <= index < self.n:
            raise IndexError(index)
        data = make_synthetic_complex(index, self.lr, self.lp, self.seed)
        return self.transform(data) if self.transform is not None else data


PepDataset = SyntheticPepDataset

DEFAULT_PAD_VALUES = {"aa": PAD_RESIDUE_INDEX, "chain_id": " ", "icode": " "}


class PaddingCollate:
    def __init__(self, length_ref_key="aa", pad_values=DEFAULT_PAD_VALUES, no_padding=(), eight=True):
        self.length_ref_key, self.pad_values, self.eight = length_ref_key, pad_values, eight
        self.no_padding = set(no_padding)

    def _pad(self, key, x, n):
        value = self.pad_values.get(key, 0)
        if isinstance(x, torch.Tensor):
            if x.size(0) == n:
                return x
            pad = torch.full([n - x.size(0)] + list(x.shape[1:]), fill_value=value).to(x)
            return torch.cat([x, pad], dim=0)
        if isinstance(x, list):
            return x + [value] * (n - len(x))
        return x

    def __call__(self, data_list):
        n = max(d[self.length_ref_key].size(0) for d in data_list)
        if self.eight:
            n = math.ceil(n / 8) * 8
        keys = set(data_list[0].keys())
        for d in data_list[1:]:
            keys &= set(d.keys())
        padded = []
        for d in data_list:
            item = {k: (v if k in self.no_padding else self._pad(k, v, n)) for k, v in d.items() if k in keys}
            l = d[self.length_ref_key].size(0)
            item["res_mask"] = torch.cat([torch.ones(l, dtype=torch.bool), torch.zeros(n - l, dtype=torch.bool)])
            padded.append(item)
        return torch.utils.data.default_collate(padded)


def synthetic_batch(num_complexes, len_pocket, len_peptide, seed=0, first_index=0, eight=False):
    """Collated batch of independent synthetic complexes (indices first_index ... )."""
    items = [make_synthetic_complex(first_index + i, len_pocket, len_peptide, seed) for i in range(num_complexes)]
    return PaddingCollate(eight=eight)(items)
