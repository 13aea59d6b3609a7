"""GPU, statistical: 200 free-running Euler steps of the CUDA path (its own random numbers) against the unmodified reference's
200-step runs (tests/golden/sampling_stats.npz).  Kept in its own module, collected last: it is the only GPU test whose
outcome is a statistical statement (false-alarm probability ~1e-3 by construction)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model(dev, state_dict):
    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    cfg, _ = load_config()
    m = FlowModel(cfg.model).eval()
    m.load_state_dict(state_dict)
    return m.to(dev)


def _ks_statistic(a, b):
    """two-sample Kolmogorov-Smirnov statistic sup |F_a - F_b|"""
    a, b = np.sort(np.asarray(a, dtype=np.float64)), np.sort(np.asarray(b, dtype=np.float64))
    grid = np.concatenate([a, b])
    fa = np.searchsorted(a, grid, side="right") / a.size
    fb = np.searchsorted(b, grid, side="right") / b.size
    return float(np.abs(fa - fb).max())


def test_200_step_sampling_statistics(dev, model):
    """SURVEY section 4 tier 3: 200 free-running Euler steps with the CUDA path's own random numbers (torch's GPU generator
    for the initial noise, Philox inside the kernels for the residue types) against the UNMODIFIED reference's 200-step
    runs with ITS random number path on the same 32 complexes and weights (tests/golden/sampling_stats.npz, written by
    tests/golden/make_golden_sampling_stats.py).  Per sample: CA RMSD / rotation RMSD / amino-acid recovery of
    models_con/inference.py:77-79 and p_gt, the probability the final denoiser call gives the true residue type (the
    expectation of the recovery).  Same distribution: the unit of the mean tests is the complex (draws of one complex are
    correlated) - paired t over the 32 complexes below 4; pooled two-sample KS under the 0.1 % critical value for the two
    continuous metrics; p_gt, which has almost no sampling noise, within 5 % of the reference's."""
    import os
    from pepflowww_b200.pep_dataloader import synthetic_batch
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sampling_stats.npz"))
    n_complex, pocket, peptide = int(g["n_complex"]), int(g["pocket"]), int(g["peptide"])
    draws, steps = int(g["draws"]), int(g["steps"])
    host = synthetic_batch(n_complex, pocket, peptide, seed=int(g["data_seed"]))
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    gm = host["generate_mask"]
    gl = gm.long()
    n = gm.sum(-1).float() + 1e-8
    ours = {"tran": [], "rot": [], "aar": [], "p_gt": []}
    torch.manual_seed(2026)
    for _ in range(draws):
        smp = model.sampler_init(batch, num_steps=steps, stream_to_host=True)
        for i in range(steps):
            smp.step(i)
        final = smp.trajectory_to_host()[-1]
        p = torch.softmax(smp.pred[3].cpu(), -1).gather(-1, host["aa"].clamp(0, 19)[..., None])[..., 0]
        ours["tran"].append(torch.sqrt(((final["trans"] - final["trans_1"]) ** 2 * gl[..., None]).sum((-1, -2)) / n))
        ours["rot"].append(torch.sqrt(((final["rotmats"] - final["rotmats_1"]) ** 2 * gl[..., None, None]).sum((-1, -2, -3)) / n))
        ours["aar"].append(((final["seqs"] == final["seqs_1"]) * gl).sum(-1) / n)
        ours["p_gt"].append((p * gm).sum(-1) / n)
    ks_crit = 1.95 * math.sqrt(2.0 / (n_complex * draws))      # alpha = 0.001
    for k in ("tran", "rot", "aar", "p_gt"):
        a = torch.stack(ours[k]).numpy()                 # [draws, n_complex]
        b = g[k]
        d = a.mean(0) - b.mean(0)
        t_stat = float(d.mean() / (d.std(ddof=1) / math.sqrt(d.size) + 1e-12))
        ks = _ks_statistic(a.reshape(-1), b.reshape(-1))
        print(f"200-step {k}: ours {a.mean():.4f} +- {a.std():.4f}  reference {b.mean():.4f} +- {b.std():.4f}  "
              f"paired t over complexes {t_stat:+.2f}  KS {ks:.3f} (crit {ks_crit:.3f})")
        assert abs(t_stat) < 4.0, k
        if k in ("tran", "rot"):                         # aar takes 9 discrete values: its mean test and p_gt cover it
            assert ks < ks_crit, k
        if k == "p_gt":
            assert abs(a.mean() - b.mean()) < 0.05 * b.mean(), (a.mean(), b.mean())
