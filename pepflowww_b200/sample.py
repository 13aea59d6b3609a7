"""From a sampled trajectory to PDB files - the reference's models_con/sample.py:68-120 (save_samples_bb,
save_samples_sc), which consume what inference.py:105-106 stores: the last entry of FlowModel.sample plus the batch.
The geometry runs in the reconstruction kernels (csrc/pf_recon.cu) on the device the batch lives on (or cuda:0 when the
stored sample is on the host); the PDB text is written by writers.save_pdb."""
import os

import torch
import torch.nn.functional as F

from . import geometry, torsion
from .utils import recursive_to
from .writers import save_pdb


def _device_of(samples):
    dev = samples["batch"]["aa"].device
    return dev if dev.type == "cuda" else torch.device("cuda:0")


def _meta(batch):
    """Per-residue chain ids / residue numbers / insertion codes of the (replicated) complex.  The collate function
    turns the per-item lists into L tuples of B strings (sample.py:71, "fix chain id in collate func"); synthetic
    batches carry none of the three: peptide = chain A, pocket = chain B, residue numbers from res_nb."""
    L = batch["aa"].shape[1]
    if "chain_id" in batch:
        chain_id = [list(item) for item in zip(*batch["chain_id"])][0]
    else:
        chain_id = ["A" if g else "B" for g in batch["generate_mask"][0].tolist()]
    resseq = batch["resseq"][0] if "resseq" in batch else batch["res_nb"][0]
    return {"chain_nb": batch["chain_nb"][0].cpu(), "chain_id": chain_id, "resseq": resseq.cpu(), "icode": [" "] * L}


def side_chain_atoms(samples):
    """(pos [B,L,15,3], mask [B,L,15], aa [B,L]) on the host: generated residues rebuilt from frames + torsions + sampled
    types, context residues kept as they are in the batch (sample.py:105-109)."""
    dev = _device_of(samples)
    batch = recursive_to(samples["batch"], dev)
    s = {k: samples[k].to(dev) for k in ("rotmats", "trans", "angles", "seqs")}
    pos_ha, mask_new = torsion.reconstruct_side_chains(s)
    pos_new = torch.where(batch["generate_mask"][:, :, None, None], pos_ha, batch["pos_heavyatom"])
    return pos_new.cpu(), mask_new.cpu(), s["seqs"].cpu()


def backbone_atoms(samples):
    """(pos [B,L,15,3], mask [B,L,15], aa [B,L]): generated residues as N, CA, C, O only (sample.py:77-83)."""
    dev = _device_of(samples)
    batch = recursive_to(samples["batch"], dev)
    s = {k: samples[k].to(dev) for k in ("rotmats", "trans", "seqs")}
    pos_bb = geometry.reconstruct_backbone(s["rotmats"], s["trans"], s["seqs"], batch["chain_nb"], batch["res_nb"],
                                           batch["res_mask"])
    pos_ha = F.pad(pos_bb, pad=(0, 0, 0, 15 - 4), value=0.0)
    pos_new = torch.where(batch["generate_mask"][:, :, None, None], pos_ha, batch["pos_heavyatom"])
    mask_bb = torch.zeros_like(batch["mask_heavyatom"])
    mask_bb[:, :, :4] = True
    mask_new = torch.where(batch["generate_mask"][:, :, None], mask_bb, batch["mask_heavyatom"])
    return pos_new.cpu(), mask_new.cpu(), s["seqs"].cpu()


def _write_all(samples, save_dir, atoms):
    os.makedirs(save_dir, exist_ok=True)
    batch = recursive_to(samples["batch"], "cpu")
    meta = _meta(batch)
    pos_new, mask_new, aa_new = atoms
    paths = []
    for i in range(aa_new.shape[0]):
        paths.append(os.path.join(save_dir, f"sample_{i}.pdb"))
        save_pdb(dict(meta, aa=aa_new[i], mask_heavyatom=mask_new[i], pos_heavyatom=pos_new[i]), path=paths[-1])
    paths.append(os.path.join(save_dir, "gt.pdb"))
    save_pdb(dict(meta, aa=batch["aa"][0], mask_heavyatom=batch["mask_heavyatom"][0],
                  pos_heavyatom=batch["pos_heavyatom"][0]), path=paths[-1])
    return paths


def save_samples_sc(samples, save_dir):
    """sample_i.pdb (full atoms) for every sampled replica and gt.pdb (sample.py:96-120).  Returns the paths."""
    return _write_all(samples, save_dir, side_chain_atoms(samples))


def save_samples_bb(samples, save_dir):
    """The same with backbone atoms only on the generated residues (sample.py:68-94)."""
    return _write_all(samples, save_dir, backbone_atoms(samples))
