"""Golden outputs of the UNMODIFIED reference at the headline residue count (256 + 15): FlowModel.encode + one
GAEncoder.forward on two synthetic complexes.  Build-container only:   python tests/golden/make_golden_headline.py

Only the OUTPUTS are stored (tests/golden/ga_encoder_headline.npz, ~90 KB): the inputs are a deterministic function of the
seeds below (synthetic_batch -> the reference's own encode -> seeded noise), which tests/test_oracle_golden.py rebuilds
through the oracle's encode.  That test therefore pins, at the benchmark shape, the oracle's embedders AND its denoiser
against the reference in one go (the small-shape fixtures pin them separately)."""
import math
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
from pepflowww_b200.pep_dataloader import synthetic_batch  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402

from make_golden_headline_inputs import DATA_SEED, NOISE_SEED, WEIGHT_SEED, headline_inputs  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_shim.load_reference()
    cfg, _ = ns.load_config("/root/reference/configs/learn_angle.yaml")
    model = ns.FlowModel(cfg.model).eval()
    model.load_state_dict(deterministic_state_dict(model.state_dict(), WEIGHT_SEED))
    torch.set_grad_enabled(False)
    batch = synthetic_batch(2, 256, 15, seed=DATA_SEED)
    r1, x1, a1, s1, node, edge = model.encode(batch)
    enc = dict(rotmats_1=r1, trans_1=x1, angles_1=a1, seqs_1=s1, node_embed=node, edge_embed=edge)
    inp = headline_inputs(enc, batch)
    keys = ("t", "rotmats_t", "trans_t", "angles_t", "seqs_t", "node_embed", "edge_embed", "generate_mask", "res_mask")
    R, x, ang, logits = model.ga_encoder(*[inp[k] for k in keys])
    np.savez_compressed(os.path.join(HERE, "ga_encoder_headline.npz"), out_rotmats=R.numpy(), out_trans=x.numpy(),
                        out_angles=ang.numpy(), out_logits=logits.numpy(),
                        node_embed_checksum=np.array([float(node.double().sum()), float(node.double().abs().sum())]),
                        edge_embed_checksum=np.array([float(edge.double().sum()), float(edge.double().abs().sum())]),
                        seeds=np.array([WEIGHT_SEED, DATA_SEED, NOISE_SEED]))
    print("wrote ga_encoder_headline.npz", os.path.getsize(os.path.join(HERE, "ga_encoder_headline.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
