"""Sampling driver with the reference's command line (models_con/inference.py:40-110): for every complex of the dataset
replicate it `num_samples` times, evaluate the losses once, run FlowModel.sample, record CA / rotation RMSD and
amino-acid recovery of the final state against the ground truth, store `<output>/outputs/<id>.pt` (the last trajectory
entry plus the batch - the input of sample.save_samples_sc / _bb) and `<output>/outputs.csv`.

    python -m pepflowww_b200.inference --config configs/learn_angle.yaml --device cuda:0 --ckpt model1.pt --output out

The PepMerge LMDB is not available offline: without --ckpt the weights are the seeded random initialisation, and the
dataset is the synthetic generator of pep_dataloader.py with the item schema of PepDataset."""
import argparse
import csv
import os
from copy import deepcopy

import torch

from .config import load_config
from .flow_model import FlowModel
from .pep_dataloader import PaddingCollate, SyntheticPepDataset
from .utils import load_checkpoint, process_dic, recursive_to, seed_all

collate_fn = PaddingCollate(eight=False)


def sample_metrics(final, batch):
    """CA RMSD, rotation-matrix RMSD and amino-acid recovery over the generated residues (inference.py:76-78).
    `final` = last entry of FlowModel.sample (host tensors), `batch` the collated batch."""
    gm = batch["generate_mask"].cpu()
    n = gm.sum() + 1e-8
    tran = torch.sqrt(torch.sum((final["trans"] - final["trans_1"]) ** 2 * gm[..., None].long()) / n)
    rot = torch.sqrt(torch.sum((final["rotmats"] - final["rotmats_1"]) ** 2 * gm[..., None, None].long()) / n)
    aar = torch.sum((final["seqs"] == final["seqs_1"]) * gm.long()) / n
    return {"tran": float(tran), "rot": float(rot), "aar": float(aar), "len": int(gm.sum())}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default=None)
    ap.add_argument("--device", type=str, default="cuda:0")
    ap.add_argument("--ckpt", type=str, default=None)
    ap.add_argument("--output", type=str, required=True)
    ap.add_argument("--num_steps", type=int, default=200)
    ap.add_argument("--num_samples", type=int, default=64)
    ap.add_argument("--sample_bb", type=bool, default=True)
    ap.add_argument("--sample_ang", type=bool, default=True)
    ap.add_argument("--sample_seq", type=bool, default=True)
    ap.add_argument("--num_complexes", type=int, default=2, help="synthetic dataset size")
    ap.add_argument("--pocket", type=int, default=128)
    ap.add_argument("--peptide", type=int, default=12)
    args = ap.parse_args(argv)
    config, _ = load_config(args.config) if args.config else load_config()
    device = torch.device(args.device)
    if device.type != "cuda":
        raise RuntimeError("inference needs a CUDA device (FlowModel.sample has no CPU fallback)")
    seed_all(114514)
    model = FlowModel(config.model).to(device)
    if args.ckpt:
        model.load_state_dict(process_dic(load_checkpoint(args.ckpt, map_location=device)["model"]))
    model.eval()
    dataset = SyntheticPepDataset(args.num_complexes, args.pocket, args.peptide, seed=0)
    os.makedirs(os.path.join(args.output, "outputs"), exist_ok=True)
    rows = []
    for i in range(len(dataset)):
        item = dataset[i]
        batch = recursive_to(collate_fn([deepcopy(item) for _ in range(args.num_samples)]), device)
        with torch.no_grad():
            loss_dic = model(batch)
        traj = model.sample(batch, num_steps=args.num_steps, sample_bb=args.sample_bb, sample_ang=args.sample_ang,
                            sample_seq=args.sample_seq)
        m = sample_metrics(traj[-1], batch)
        m.update(id=batch["id"][0], trans_loss=float(loss_dic["trans_loss"]), rot_loss=float(loss_dic["rot_loss"]))
        print({k: float(v) for k, v in loss_dic.items()})
        print(f"tran:{m['tran']},rot:{m['rot']},aar:{m['aar']},len:{m['len']}")
        rows.append(m)
        final = dict(traj[-1])
        final["batch"] = recursive_to(batch, "cpu")
        torch.save(final, os.path.join(args.output, "outputs", f"{batch['id'][0]}.pt"))
    with open(os.path.join(args.output, "outputs.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["id", "len", "tran", "aar", "rot", "trans_loss", "rot_loss"])
        w.writeheader()
        w.writerows(rows)
    return rows


if __name__ == "__main__":
    main()
