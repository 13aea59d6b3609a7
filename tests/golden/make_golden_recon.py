"""Golden fixture for the post-sampling reconstruction: runs the UNMODIFIED reference's
full_atom_reconstruction / get_heavyatom_mask (models_con/torsion.py:126-226) and reconstruct_backbone
(pepflow/modules/common/geometry.py:446-489) on seeded inputs and writes tests/golden/reconstruction.npz.
Build-container only:  python tests/golden/make_golden_recon.py
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_shim  # noqa: E402
from make_golden import random_rotations, rng_for, save  # noqa: E402


def main():
    ref_shim.load_reference()
    from models_con import torsion
    from pepflow.modules.common import geometry

    rng = rng_for("reconstruction")
    B, L = 3, 47
    aa = torch.from_numpy(rng.integers(0, 20, size=(B, L))).long()
    aa[0, :20] = torch.arange(20)          # every residue type at least once
    aa[1, 5] = 20                          # UNK rows: zero rigid groups
    aa[2, -1] = 20
    R = random_rotations(rng, (B, L))
    t = torch.from_numpy(rng.standard_normal((B, L, 3)).astype(np.float32)) * 8.0
    angles = torch.from_numpy(rng.uniform(0, 2 * np.pi, size=(B, L, 5)).astype(np.float32))
    angles[0, 0] = 0.0
    # two chains, a numbering gap, a descending pair, trailing padding
    res_nb = torch.arange(1, L + 1).repeat(B, 1)
    chain_nb = torch.zeros(B, L, dtype=torch.long)
    chain_nb[:, 30:] = 1
    res_nb[:, 30:] = torch.arange(1, L - 30 + 1)
    res_nb[1, 10:20] += 3
    res_nb[2, 7] = res_nb[2, 6] - 1
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[2, L - 6:] = False
    mask[1, 12] = False

    pos14, R_ret, t_ret = torsion.full_atom_reconstruction(R, t, angles, aa)
    mask15 = torsion.get_heavyatom_mask(aa)
    pos_bb = geometry.reconstruct_backbone(R, t, aa, chain_nb, res_nb, mask)
    # the inverse map, one complex at a time as the dataset builder calls it (models_con/torsion.py:49-66); a second
    # input with degenerate side chains (all atoms of some residues collapsed onto CA) exercises the NaN -> mask path
    tor = [torsion.get_torsion_angle(pos14[b], aa[b]) for b in range(B)]
    pos_deg = pos14.clone()
    pos_deg[0, 3:9, 4:] = pos_deg[0, 3:9, 1:2]
    tor_deg = torsion.get_torsion_angle(pos_deg[0], aa[0])
    save("reconstruction", torsion=torch.stack([x[0] for x in tor]), torsion_mask=torch.stack([x[1] for x in tor]),
         pos_deg=pos_deg[0], torsion_deg=tor_deg[0], torsion_mask_deg=tor_deg[1], R=R, t=t, angles=angles, aa=aa, chain_nb=chain_nb, res_nb=res_nb, mask=mask,
         pos14=pos14, R_ret=R_ret, t_ret=t_ret, mask15=mask15, pos_bb=pos_bb)


if __name__ == "__main__":
    main()
