"""Synthetic peptide-pocket complexes (SURVEY.md section 8d generator), deterministic weights and residue tables.

Facts restated from the reference: AA enum order pepflow/modules/protein/constants.py:53-58, heavy-atom slots :95-117,
number of chi angles per residue type :402-424, torsions_mask models_con/torsion.py:230-232, the batch schema of
models_con/pep_dataloader.py:41-66 and the collate of pepflow/utils/data.py:63-78.
"""
import math
import os
import zlib

import numpy as np
import torch

_AA = "ALA CYS ASP GLU PHE GLY HIS ILE LYS LEU MET ASN PRO GLN ARG SER THR VAL TRP TYR".split()
_N_SIDE = dict(ALA=1, ARG=7, ASN=4, ASP=4, CYS=2, GLN=5, GLU=5, GLY=0, HIS=6, ILE=4, LEU=4, LYS=5, MET=4, PHE=7, PRO=3,
               SER=2, THR=3, TRP=10, TYR=8, VAL=3)
_N_CHI = dict(ALA=0, ARG=4, ASN=2, ASP=2, CYS=1, GLN=3, GLU=3, GLY=0, HIS=2, ILE=2, LEU=2, LYS=4, MET=3, PHE=2, PRO=2,
              SER=1, THR=1, TRP=2, TYR=2, VAL=1)


def _tables():
    hm = torch.zeros(22, 15, dtype=torch.bool)      # N CA C O + side chain; slot 14 (OXT) never set; UNK / PAD empty
    tm = torch.zeros(22, 5, dtype=torch.float32)    # [psi, chi1..chi4]; UNK row = psi only; PAD row = 0
    for i, name in enumerate(_AA):
        hm[i, :4 + _N_SIDE[name]] = True
        tm[i, :1 + _N_CHI[name]] = 1.0
    tm[20, 0] = 1.0
    return hm, tm


restype_to_heavyatom_masks, torsions_mask = _tables()


def _unit(rng, n):
    v = rng.standard_normal((n, 3))
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_synthetic_complex(index, len_pocket=128, len_peptide=12, seed=0):
    rng = np.random.Generator(np.random.PCG64([seed, index]))
    lr, lp = len_pocket, len_peptide
    L = lr + lp
    aa = rng.integers(0, 20, size=L)
    steps = _unit(rng, lp) * 3.8
    ca_pep = np.cumsum(steps, axis=0)
    anchor = rng.integers(0, lp, size=lr)
    ca_rec = ca_pep[anchor] + _unit(rng, lr) * rng.uniform(5.0, 12.0, size=(lr, 1))
    ca = np.concatenate([ca_rec, ca_pep], axis=0)
    ca = ca - ca_pep.mean(axis=0, keepdims=True)
    pos = ca[:, None, :] + rng.normal(0.0, 1.5, size=(L, 15, 3))
    u1 = _unit(rng, L)
    u2 = _unit(rng, L)
    u2 = u2 - 0.5 * (u1 * u2).sum(-1, keepdims=True) * u1
    u2 = u2 / np.linalg.norm(u2, axis=-1, keepdims=True)
    pos[:, 0] = ca + 1.46 * u1
    pos[:, 1] = ca
    pos[:, 2] = ca + 1.52 * u2
    aa_t = torch.from_numpy(aa).long()
    mask = restype_to_heavyatom_masks[aa_t]
    pos_t = torch.from_numpy(pos).float() * mask[..., None]
    tmask = torsions_mask[aa_t]
    tors = torch.from_numpy(rng.uniform(0.0, 2 * math.pi, size=(L, 5))).float() * tmask
    return {"aa": aa_t, "pos_heavyatom": pos_t, "mask_heavyatom": mask.clone(),
            "res_nb": torch.cat([torch.arange(1, lr + 1), torch.arange(1, lp + 1)]).long(),
            "chain_nb": torch.cat([torch.ones(lr), torch.zeros(lp)]).long(),
            "generate_mask": torch.cat([torch.zeros(lr), torch.ones(lp)]).bool(),
            "torsion_angle": tors, "torsion_angle_mask": tmask.bool()}


def synthetic_batch(num_complexes, len_pocket, len_peptide, seed=0, first_index=0):
    """Collated batch of equally sized complexes (no padding needed): every field stacked, res_mask all True."""
    items = [make_synthetic_complex(first_index + i, len_pocket, len_peptide, seed) for i in range(num_complexes)]
    out = {k: torch.stack([it[k] for it in items], 0) for k in items[0]}
    out["res_mask"] = torch.ones(num_complexes, len_pocket + len_peptide, dtype=torch.bool)
    return out


def state_dict_spec():
    """{key: shape} of the reference FlowModel's state_dict (6,880,353 parameters)."""
    spec = {}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_dict_keys.txt")) as fh:
        for line in fh:
            parts = line.split()
            spec[parts[0]] = tuple(int(p) for p in parts[1:])
    return spec


_FREQ_BANDS = {"node_embedder.dihed_embed.freq_bands": 3, "edge_embedder.dihedral_embed.freq_bands": 3,
               "ga_encoder.angles_embedder.freq_bands": 12}


def deterministic_state_dict(reference_sd, seed=114514):
    """Same filler as pepflowww_b200.utils.deterministic_state_dict (numpy PCG64 keyed by (seed, crc32(key)))."""
    out = {}
    for key, ref in reference_sd.items():
        if key.endswith("freq_bands"):
            out[key] = ref.clone()
            continue
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))
        u = torch.from_numpy(rng.random(tuple(ref.shape), dtype=np.float32) * 2.0 - 1.0)
        leaf = key.split(".")[-1]
        if key.endswith("head_weights"):
            val = 0.541324854612918 + 0.3 * u
        elif "embed.weight" in key or key.endswith("current_seq_embedder.weight"):
            val = u
        elif key.endswith("aapair_to_distcoef.weight"):
            val = 0.5 * u
        elif ref.dim() >= 2:
            val = u * float(np.sqrt(3.0 / ref.shape[-1]))
        elif leaf == "bias" or leaf == "in_proj_bias":
            val = 0.1 * u
        elif leaf == "weight":
            val = 1.0 + 0.1 * u
        else:
            val = 0.1 * u
        out[key] = val.to(ref.dtype)
    return out


def reference_state_dict(seed=114514):
    """Deterministic non-degenerate weights with the reference's keys / shapes, built without instantiating a model."""
    proto = {k: torch.zeros(s) for k, s in state_dict_spec().items()}
    for k, n in _FREQ_BANDS.items():
        proto[k] = torch.tensor([float(i + 1) for i in range(n)] + [1.0 / (i + 1) for i in range(n)])
    return deterministic_state_dict(proto, seed)
