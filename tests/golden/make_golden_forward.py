"""Golden fixture for the training-loss arithmetic: the UNMODIFIED reference's
FlowModel.forward (models_con/flow_model.py:111-227) on the `encode` fixture's batch, with the corruption noise it drew
recorded next to the six losses so the other side can inject it.  Build-container only:
    python tests/golden/make_golden_forward.py
Rewrites tests/golden/forward_losses.npz (same seeds and losses as the section-5 block of make_golden.py)."""
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
from make_golden import WEIGHT_SEED, save  # noqa: E402
from oracle import pepflow_oracle as orc  # noqa: E402
from pepflowww_b200.pep_dataloader import synthetic_batch  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402


def main():
    ns = ref_shim.load_reference()
    cfg, _ = ns.load_config("/root/reference/configs/learn_angle.yaml")
    torch.manual_seed(0)
    model = ns.FlowModel(cfg.model).eval()
    model.load_state_dict(deterministic_state_dict(model.state_dict(), WEIGHT_SEED))
    torch.set_grad_enabled(False)
    batch = synthetic_batch(2, 18, 5, seed=7)
    B, L = batch["aa"].shape
    ns.flow_model_mod.sample_from = lambda c: orc.categorical_from_uniform(c, torch.full(c.shape[:2], 0.5))
    np.random.seed(321)
    torch.manual_seed(321)
    losses = model(batch)
    # replay the reference's RNG order (flow_model.py:126-150) to recover what it drew
    np.random.seed(321)
    torch.manual_seed(321)
    from pepflow.modules.so3.dist import uniform_so3
    t = torch.rand((B, 1))
    t = t * (1 - 2 * cfg.model.interpolant.min_t) + cfg.model.interpolant.min_t
    trans_0 = torch.randn((B, L, 3))
    rotmats_0 = uniform_so3(B, L)
    angles_0 = torch.rand((B, L, 5)) * 2 * np.pi
    seqs_0_simplex = model.k * torch.randn((B, L, 20))
    old = np.load(os.path.join(HERE, "forward_losses.npz"))
    for k, v in losses.items():
        assert abs(float(v) - float(old[k])) <= 1e-6 * abs(float(old[k])), (k, float(v), float(old[k]))
    save("forward_losses", **{k: v for k, v in losses.items()}, t=t, trans_0=trans_0, rotmats_0=rotmats_0,
         angles_0=angles_0, seqs_0_simplex=seqs_0_simplex, u_t=torch.full((B, L), 0.5), u_pred=torch.full((B, L), 0.5))
    print({k: float(v) for k, v in losses.items()})


if __name__ == "__main__":
    main()
