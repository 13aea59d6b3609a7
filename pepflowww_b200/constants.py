"""Residue / atom tables the hot path indexes.

Restates the facts (not the code) of the reference's
pepflow/modules/protein/constants.py: AA enum order (:53-58), heavy-atom slot names
(:95-117), chi_angles_mask (:402-424); and models_con/torsion.py:122-124
(restype_to_heavyatom_masks) and :230-232 (torsions_mask).
"""
import enum

import torch

PAD_RESIDUE_INDEX = 21
max_num_heavyatoms = 15
MAX_NUM_HEAVYATOMS = max_num_heavyatoms
NUM_AA_SLOTS = 22  # 20 standard + UNK + PAD


class AA(enum.IntEnum):
    ALA = 0; CYS = 1; ASP = 2; GLU = 3; PHE = 4
    GLY = 5; HIS = 6; ILE = 7; LYS = 8; LEU = 9
    MET = 10; ASN = 11; PRO = 12; GLN = 13; ARG = 14
    SER = 15; THR = 16; VAL = 17; TRP = 18; TYR = 19
    UNK = 20


class BBHeavyAtom(enum.IntEnum):
    N = 0; CA = 1; C = 2; O = 3; CB = 4; OXT = 14


# side-chain heavy atoms after the N, CA, C, O backbone slots (PDB naming)
_SIDECHAIN = {
    "ALA": "CB", "ARG": "CB CG CD NE CZ NH1 NH2", "ASN": "CB CG OD1 ND2",
    "ASP": "CB CG OD1 OD2", "CYS": "CB SG", "GLN": "CB CG CD OE1 NE2",
    "GLU": "CB CG CD OE1 OE2", "GLY": "", "HIS": "CB CG ND1 CD2 CE1 NE2",
    "ILE": "CB CG1 CG2 CD1", "LEU": "CB CG CD1 CD2", "LYS": "CB CG CD CE NZ",
    "MET": "CB CG SD CE", "PHE": "CB CG CD1 CD2 CE1 CE2 CZ", "PRO": "CB CG CD",
    "SER": "CB OG", "THR": "CB OG1 CG2",
    "TRP": "CB CG CD1 CD2 NE1 CE2 CE3 CZ2 CZ3 CH2",
    "TYR": "CB CG CD1 CD2 CE1 CE2 CZ OH", "VAL": "CB CG1 CG2",
}
# number of defined chi angles per residue type
_NUM_CHI = {
    "ALA": 0, "ARG": 4, "ASN": 2, "ASP": 2, "CYS": 1, "GLN": 3, "GLU": 3, "GLY": 0,
    "HIS": 2, "ILE": 2, "LEU": 2, "LYS": 4, "MET": 3, "PHE": 2, "PRO": 2, "SER": 1,
    "THR": 1, "TRP": 2, "TYR": 2, "VAL": 1, "UNK": 0,
}


def heavyatom_names(aa: AA):
    """15 slot names for a residue type; '' for an unused slot; slot 14 is OXT."""
    if aa == AA.UNK:
        return [""] * max_num_heavyatoms
    side = _SIDECHAIN[aa.name].split()
    if aa == AA.GLY:
        names = ["N", "CA", "C", "O", ""]
    else:
        names = ["N", "CA", "C", "O"] + side
    names = names + [""] * (max_num_heavyatoms - 1 - len(names)) + ["OXT"]
    return names


def _build_heavyatom_masks():
    m = torch.zeros(NUM_AA_SLOTS, max_num_heavyatoms, dtype=torch.bool)
    for aa in AA:
        m[int(aa)] = torch.tensor([n != "" and n != "OXT" for n in heavyatom_names(aa)])
    return m


def _build_torsions_mask():
    # [psi, chi1..chi4]; rows 0..20 get psi=1 (UNK included), row 21 (PAD) is all zero
    m = torch.zeros(NUM_AA_SLOTS, 5, dtype=torch.float32)
    for aa in AA:
        n = _NUM_CHI[aa.name]
        m[int(aa)] = torch.tensor([1.0] + [1.0] * n + [0.0] * (4 - n))
    return m


restype_to_heavyatom_masks = _build_heavyatom_masks()
torsions_mask = _build_torsions_mask()


# ---- rigid-group tables of the post-sampling reconstruction (reference constants.py:665-749, 878-890) ----------
# Ideal-geometry data, shipped as pepflowww_b200/data/restype_rigid_tables.npz (scripts/make_restype_tables.py).
_RIGID_TABLES = {}


def rigid_tables(device="cpu"):
    """dict of tensors on `device`: rigid_rot [21,8,3,3], rigid_trans [21,8,3], atom_group [21,14] i32,
    atom_pos [21,14,3], bb_coords [21,3,3], bb_oxygen [21,3], chi_atoms [21,4,4] i32, heavyatom_mask [22,15] u8."""
    import os

    import numpy as np
    device = torch.device(device)
    key = (device.type, device.index)
    if key not in _RIGID_TABLES:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "restype_rigid_tables.npz")
        with np.load(path) as z:
            t = {k: torch.from_numpy(z[k]).to(device).contiguous() for k in z.files}
        t["heavyatom_mask"] = restype_to_heavyatom_masks.to(torch.uint8).to(device).contiguous()
        _RIGID_TABLES[key] = t
    return _RIGID_TABLES[key]
