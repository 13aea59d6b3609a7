"""Writes pepflowww_b200/data/restype_rigid_tables.npz: the per-residue-type rigid-group tables that
full_atom_reconstruction / reconstruct_backbone index (ideal-geometry constants of the 20 amino acids).

Build-container only (`python scripts/make_restype_tables.py`): the numbers are read from the tensors the UNMODIFIED
reference builds at import time (pepflow/modules/protein/constants.py:665-668 filled by :670-749, :878-879 filled by
:881-890), through the import shim the golden fixtures use.  The data file is committed; nothing reads
/root/reference at run time.

  rigid_rot   [21, 8, 3, 3] f32  rotation of rigid group g (backbone, omega, phi, psi, chi1..chi4) in its parent frame
  rigid_trans [21, 8, 3]    f32  its translation
  atom_group  [21, 14]      i32  rigid group each atom14 slot belongs to
  atom_pos    [21, 14, 3]   f32  position of the slot in its group's frame
  bb_coords   [21, 3, 3]    f32  N, CA, C in the backbone frame
  bb_oxygen   [21, 3]       f32  O in the psi frame
  chi_atoms   [21, 4, 4]    i32  atom14 slots of the four atoms that define chi1..chi4 (-1: the angle does not exist;
                                 constants.py:372-400 looked up through restype_atom14_name_to_index :150-155)
Row 20 (UNK) is zero everywhere except bb_* which the reference fills for every row it has data for.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import ref_shim  # noqa: E402


def main():
    ref_shim.load_reference()
    import pepflow.modules.protein.constants as C
    out = dict(
        rigid_rot=C.restype_rigid_group_rotation.numpy().astype(np.float32),
        rigid_trans=C.restype_rigid_group_translation.numpy().astype(np.float32),
        atom_group=C.restype_heavyatom_to_rigid_group.numpy().astype(np.int32),
        atom_pos=C.restype_heavyatom_rigid_group_positions.numpy().astype(np.float32),
        bb_coords=C.backbone_atom_coordinates_tensor.numpy().astype(np.float32),
        bb_oxygen=C.bb_oxygen_coordinate_tensor.numpy().astype(np.float32),
    )
    chi = np.full((21, 4, 4), -1, dtype=np.int32)
    for aa in range(21):
        for i, names in enumerate(C.chi_angles_atoms[C.AA(aa)] if C.AA(aa) in C.chi_angles_atoms else []):
            chi[aa, i] = [C.restype_atom14_name_to_index[C.AA(aa)][n] for n in names]
    out["chi_atoms"] = chi
    assert out["rigid_rot"].shape == (21, 8, 3, 3) and out["atom_group"].max() <= 7
    path = os.path.join(ROOT, "pepflowww_b200", "data", "restype_rigid_tables.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(np.abs(v).sum()))


if __name__ == "__main__":
    main()
