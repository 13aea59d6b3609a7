"""Statistical fixture of 200-step FlowModel.sample from the UNMODIFIED reference (/root/reference, CPU fp32), with the
reference's OWN random number path (torch.multinomial for the residue types, SciPy rotations / torch.randn for the initial
noise): per sample the three metrics the reference's inference driver records at models_con/inference.py:77-79 - CA RMSD,
rotation-matrix RMSD and amino-acid recovery of the final state over the generated residues.  Build-container only:

    python tests/golden/make_golden_sampling_stats.py

32 synthetic complexes (20-residue pocket, 8-residue peptide, the generator of pepflowww_b200.pep_dataloader) x 16 noise
draws = 512 samples; writes tests/golden/sampling_stats.npz.  Besides the three realised metrics it records `p_gt`, the
mean softmax probability the final denoiser call assigns to the true residue type over the generated residues (the
EXPECTATION of the amino-acid recovery, read from the logits of the last GAEncoder.forward call through a forward hook -
the reference code itself is not modified): the realised recovery of 8 residues is a coarse count (binomial noise
+-0.014 per draw of 256 residues), its expectation is what separates two samplers.  The GPU test (tests/test_gpu_zz_sampling_statistics.py::
test_200_step_sampling_statistics) draws the same number of samples from the CUDA path with its own RNG (Philox) and
compares the distributions (means within 4 standard errors, two-sample Kolmogorov-Smirnov statistic).
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_shim  # noqa: E402
from pepflowww_b200.pep_dataloader import synthetic_batch  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402

WEIGHT_SEED = 114514
N_COMPLEX, POCKET, PEPTIDE, DRAWS, STEPS, DATA_SEED = 32, 20, 8, 16, 200, 31


def metrics(final, gm):
    """inference.py:77-79 per complex (the reference sums over the whole batch; per-sample values give a distribution)."""
    n = gm.sum(-1).float() + 1e-8
    g = gm.long()
    tran = torch.sqrt(((final["trans"] - final["trans_1"]) ** 2 * g[..., None]).sum((-1, -2)) / n)
    rot = torch.sqrt(((final["rotmats"] - final["rotmats_1"]) ** 2 * g[..., None, None]).sum((-1, -2, -3)) / n)
    aar = ((final["seqs"] == final["seqs_1"]) * g).sum(-1) / n
    return tran, rot, aar


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_shim.load_reference()
    cfg, _ = ns.load_config("/root/reference/configs/learn_angle.yaml")
    torch.manual_seed(0)
    model = ns.FlowModel(cfg.model).eval()
    model.load_state_dict(deterministic_state_dict(model.state_dict(), WEIGHT_SEED))
    torch.set_grad_enabled(False)
    batch = synthetic_batch(N_COMPLEX, POCKET, PEPTIDE, seed=DATA_SEED)
    gm = batch["generate_mask"]
    out = {"tran": [], "rot": [], "aar": [], "p_gt": []}
    last = {}
    inner = model.ga_encoder.forward

    def recording_forward(*a, **k):
        res = inner(*a, **k)
        last["logits"] = res[3]
        return res

    model.ga_encoder.forward = recording_forward
    t0 = time.time()
    for d in range(DRAWS):
        np.random.seed(1000 + d)
        torch.manual_seed(1000 + d)
        traj = model.sample(batch, num_steps=STEPS)
        tran, rot, aar = metrics(traj[-1], gm)
        p = torch.softmax(last["logits"], -1).gather(-1, batch["aa"].clamp(0, 19)[..., None])[..., 0]
        out["tran"].append(tran); out["rot"].append(rot); out["aar"].append(aar)
        out["p_gt"].append((p * gm).sum(-1) / (gm.sum(-1).float() + 1e-8))
        print(f"draw {d}: tran {float(tran.mean()):.3f} rot {float(rot.mean()):.3f} aar {float(aar.mean()):.3f} "
              f"({time.time() - t0:.0f} s)", flush=True)
    arrs = {k: torch.stack(v).numpy() for k, v in out.items()}       # [DRAWS, N_COMPLEX]
    np.savez_compressed(os.path.join(HERE, "sampling_stats.npz"), n_complex=N_COMPLEX, pocket=POCKET, peptide=PEPTIDE,
                        draws=DRAWS, steps=STEPS, data_seed=DATA_SEED, weight_seed=WEIGHT_SEED, **arrs)
    print("wrote sampling_stats.npz", {k: (float(v.mean()), float(v.std())) for k, v in arrs.items()})


if __name__ == "__main__":
    main()
