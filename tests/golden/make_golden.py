"""Generates the committed golden fixtures by running the UNMODIFIED reference
(/root/reference, CPU fp32) on seeded synthetic inputs.  Build-container only:

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Weights are NOT stored: both sides rebuild them with
pepflowww_b200.utils.deterministic_state_dict(seed) (numpy PCG64, torch-independent).
The categorical draws of FlowModel.sample are made reproducible by replacing
`models_con.flow_model.sample_from` (imported by name at flow_model.py:18) with an inverse-CDF
sampler fed from a recorded uniform stream - the same sampler the oracle and the CUDA path use.
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
from oracle import pepflow_oracle as orc  # noqa: E402
from pepflowww_b200.pep_dataloader import synthetic_batch  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402

WEIGHT_SEED = 114514


def rng_for(tag):
    import zlib
    return np.random.Generator(np.random.PCG64([20260101, zlib.crc32(tag.encode())]))


def random_rotations(rng, shape):
    q = rng.standard_normal(tuple(shape) + (4,))
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return orc.quat_to_rot(torch.from_numpy(q).float())


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def ga_inputs(tag, B, L, n_pad):
    rng = rng_for(tag)
    f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    res_mask = torch.ones(B, L, dtype=torch.long)
    if n_pad:
        res_mask[-1, L - n_pad:] = 0
    gen = torch.zeros(B, L, dtype=torch.long)
    gen[:, L - n_pad - 5: L - n_pad] = 1
    return dict(
        t=torch.from_numpy(rng.uniform(0.01, 1.0, size=(B, 1)).astype(np.float32)),
        rotmats_t=random_rotations(rng, (B, L)),
        trans_t=f(B, L, 3) * 6.0,
        angles_t=torch.from_numpy(rng.uniform(0, 2 * np.pi, size=(B, L, 5)).astype(np.float32)),
        seqs_t=torch.from_numpy(rng.integers(0, 20, size=(B, L))).long(),
        node_embed=f(B, L, 128),
        edge_embed=f(B, L, L, 64),
        generate_mask=gen, res_mask=res_mask)


def main():
    ns = ref_shim.load_reference()
    cfg, _ = ns.load_config("/root/reference/configs/learn_angle.yaml")
    torch.manual_seed(0)
    model = ns.FlowModel(cfg.model).eval()
    sd = deterministic_state_dict(model.state_dict(), WEIGHT_SEED)
    model.load_state_dict(sd)
    torch.set_grad_enabled(False)

    # state-dict surface (keys + shapes), for the load-unchanged test
    with open(os.path.join(HERE, "state_dict_keys.txt"), "w") as fh:
        for k, v in model.state_dict().items():
            fh.write(f"{k} {' '.join(str(d) for d in v.shape)}\n")

    # ---- 1. GAEncoder.forward (+ block-0 IPA and EdgeTransition seams), padded and unpadded
    for tag, (B, L, n_pad) in {"ga_encoder_a": (2, 21, 0), "ga_encoder_b": (2, 24, 3)}.items():
        inp = ga_inputs(tag, B, L, n_pad)
        ga = model.ga_encoder
        R, x, ang, logits = ga(inp["t"], inp["rotmats_t"], inp["trans_t"], inp["angles_t"], inp["seqs_t"],
                               inp["node_embed"], inp["edge_embed"], inp["generate_mask"], inp["res_mask"])
        from data import utils as du
        rig = du.create_rigid(inp["rotmats_t"], inp["trans_t"])
        s_in = inp["node_embed"] * inp["res_mask"][..., None]
        ipa0 = ga.trunk["ipa_0"](s_in, inp["edge_embed"], rig, inp["res_mask"])
        et0 = ga.trunk["edge_transition_0"](s_in, inp["edge_embed"])
        tf0 = ga.trunk["seq_tfmr_0"](s_in, src_key_padding_mask=(1 - inp["res_mask"]).bool())
        nt0 = ga.trunk["node_transition_0"](s_in)
        upd = torch.from_numpy(rng_for(tag + "upd").standard_normal((B, L, 6)).astype(np.float32)) * 0.3
        rig2 = rig.compose_q_update_vec(upd, inp["res_mask"][..., None])
        rig3 = rig2.compose_q_update_vec(upd * 0.5, inp["res_mask"][..., None])
        save(tag, **inp, out_rotmats=R, out_trans=x, out_angles=ang, out_logits=logits,
             ipa0=ipa0, et0=et0, tf0=tf0, nt0=nt0, upd=upd,
             rig2_rot=rig2.get_rots().get_rot_mats(), rig2_trans=rig2.get_trans(),
             rig3_rot=rig3.get_rots().get_rot_mats(), rig3_trans=rig3.get_trans())

    # ---- 2. SO(3) log / exp / geodesic incl. theta ~ 0 and theta ~ pi, torus geodesic
    from data import so3_utils
    import models_con.torus as torus
    rng = rng_for("so3")
    n = 512
    base = random_rotations(rng, (n,))
    axis = rng.standard_normal((n, 3))
    axis /= np.linalg.norm(axis, axis=-1, keepdims=True)
    theta = rng.uniform(0, np.pi, size=n)
    theta[:16] = 0.0
    theta[16:48] = 10.0 ** rng.uniform(-9, -3, size=32)
    theta[48:112] = np.pi - 10.0 ** rng.uniform(-7, -1.5, size=64)
    theta[112:128] = np.pi
    rel = so3_utils.rotvec_to_rotmat(torch.from_numpy(axis * theta[:, None]).float())
    target = base @ rel
    tt = torch.from_numpy(rng.uniform(0, 1, size=(n, 1)).astype(np.float32))
    rotvec = so3_utils.rotmat_to_rotvec(rel)
    geo = so3_utils.geodesic_t(tt, target, base)
    vf = so3_utils.calc_rot_vf(base, target)
    expo = so3_utils.rotvec_to_rotmat(torch.from_numpy(axis * theta[:, None]).float() * 0.37)
    a0 = torch.from_numpy(rng.uniform(0, 2 * np.pi, size=(n, 5)).astype(np.float32))
    a1 = torch.from_numpy(rng.uniform(0, 2 * np.pi, size=(n, 5)).astype(np.float32))
    a1[:8] = a0[:8]
    tor = torus.tor_geodesic_t(tt, a1, a0)
    save("manifold", base=base, target=target, rel=rel, t=tt, rotvec=rotvec, geodesic=geo, rot_vf=vf,
         rotvec_in=torch.from_numpy(axis * theta[:, None]).float() * 0.37, exp_out=expo,
         ang0=a0, ang1=a1, tor_geodesic=tor, tor_log=torus.tor_logmap(a0, a1))

    # ---- 3. embedders + encode on a synthetic batch
    batch = synthetic_batch(2, 18, 5, seed=7)
    r1, x1, a1_, s1, node, edge = model.encode(batch)
    save("encode", **{k: v for k, v in batch.items() if isinstance(v, torch.Tensor)},
         rotmats_1=r1, trans_1=x1, angles_1=a1_, seqs_1=s1, node_embed=node, edge_embed=edge)

    # ---- 4. FlowModel.sample, 4 Euler steps, categorical draws from a recorded uniform stream
    num_steps = 4
    B, L = batch["aa"].shape
    rng = rng_for("sample")
    u_all = torch.from_numpy(rng.random((1 + 2 * num_steps, B, L), dtype=np.float32))
    calls = {"n": 0}

    def sample_from_uniform(c):
        u = u_all[calls["n"]]
        calls["n"] += 1
        return orc.categorical_from_uniform(c, u)

    ns.flow_model_mod.sample_from = sample_from_uniform
    np.random.seed(123)
    torch.manual_seed(123)
    traj = model.sample(batch, num_steps=num_steps)
    assert calls["n"] == 2 * num_steps, calls
    # replay the reference's host RNG order to recover the initial noise (flow_model.py:253-273)
    np.random.seed(123)
    torch.manual_seed(123)
    from pepflow.modules.so3.dist import uniform_so3
    gm = batch["generate_mask"]
    rot0 = torch.where(gm[..., None, None], uniform_so3(B, L), r1)
    tr0 = torch.randn(B, L, 3)
    tr0c, _ = model.zero_center_part(tr0, gm, batch["res_mask"])
    tr0c = torch.where(gm[..., None], tr0c, x1)
    ang0 = torch.where(gm[..., None], torch.rand(B, L, 5) * 2 * np.pi, a1_)
    sx0 = 5.0 * torch.randn(B, L, 20)
    s0 = torch.where(gm, orc.categorical_from_uniform(torch.softmax(sx0, -1), u_all[0]), s1)
    sx0 = torch.where(gm[..., None], sx0, model.seq_to_simplex(s1))
    out = {"uniforms": u_all, "rotmats_0": rot0, "trans_0_raw": tr0, "trans_0": tr0c, "angles_0": ang0,
           "seqs_0": s0, "seqs_0_simplex": sx0}
    for i, d in enumerate(traj):
        for k in ("rotmats", "trans", "angles", "seqs", "seqs_simplex"):
            out[f"step{i}_{k}"] = d[k]
    save("sample", **out)

    # ---- 5. training-style forward: the ga_encoder call inside FlowModel.forward is covered by (1);
    #         record the six losses for a fixed RNG state for the host-side loss arithmetic.
    ns.flow_model_mod.sample_from = lambda c: orc.categorical_from_uniform(c, torch.full(c.shape[:2], 0.5))
    np.random.seed(321)
    torch.manual_seed(321)
    losses = model(batch)
    save("forward_losses", **{k: v for k, v in losses.items()})


if __name__ == "__main__":
    main()
