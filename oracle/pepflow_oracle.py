"""CPU oracle for the PepFlow denoising hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A from-scratch restatement (fp32, CPU tensors, functional style over a plain
`state_dict`) of the reference algorithm for one sampling iteration: GAEncoder.forward
plus the manifold Euler update, and of the two once-per-sample embedders so the whole
`FlowModel.sample` can be timed as the CPU baseline.  Each function cites the reference
file:line it follows (paths relative to the reference root).

Who may import this file: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg.  The product package (pepflowww_b200/) never imports it; the
product path raises when its CUDA library is missing rather than falling back here.

Parity pin: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the UNMODIFIED reference run in the build container:
tests/golden/make_golden.py imports /root/reference and writes tests/golden/*.npz;
tests/test_oracle_golden.py checks every function here against those fixtures.

Array backend: CPU torch tensors used as a numpy-with-BLAS (matmul, einsum, elementwise);
no autograd, no nn.Module, no CUDA.
"""
import math

import torch

TWO_PI = 2.0 * math.pi

# ----------------------------------------------------------------------------- basics


def linear(x, w, b=None):
    """y = x W^T + b   (models_con/ipa_pytorch.py:116-181 is an nn.Linear)."""
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def time_embedding(t, dim=128, max_positions=2056):
    """models_con/utils.py:60-72 called from models_con/ga.py:79-85. t: [B] -> [B, dim]."""
    half = dim // 2
    tau = t * max_positions
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    arg = tau.float()[:, None] * freqs[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def angular_encoding(x, freq_bands):
    """pepflow/modules/common/layers.py:104-113. x: [..., d] -> [..., d*(1+2F)]."""
    xe = x[..., None]
    code = torch.cat([xe, torch.sin(xe * freq_bands), torch.cos(xe * freq_bands)], dim=-1)
    return code.reshape(*x.shape[:-1], -1)


def clamped_one_hot(x, num_classes):
    """pepflow/modules/common/layers.py:10-14."""
    ok = (x >= 0) & (x < num_classes)
    oh = torch.nn.functional.one_hot(x.clamp(0, num_classes - 1), num_classes)
    return oh * ok[..., None]


def seq_to_simplex(seqs, K=20, k=5.0):
    """models_con/flow_model.py:108-109."""
    return clamped_one_hot(seqs, K).float() * k * 2 - k


def categorical_from_uniform(probs, u):
    """Inverse-CDF draw standing in for `sample_from` (pepflow/modules/common/layers.py:17-22:
    multinomial(p + 1e-8, 1)).  probs [B,L,K], u [B,L] in [0,1).  The index is the number of
    prefix sums that are <= u * total, scanned left to right in fp32 (same order on the GPU)."""
    c = probs + 1e-8
    K = c.shape[-1]
    run = torch.zeros_like(c[..., 0])
    total = torch.zeros_like(run)
    for k in range(K):
        total = total + c[..., k]
    thr = u * total
    idx = torch.zeros(c.shape[:-1], dtype=torch.long)
    for k in range(K - 1):
        run = run + c[..., k]
        idx = idx + (run <= thr).long()
    return idx


# ----------------------------------------------------------------------------- rotations


def quat_to_rot(q):
    """openfold/utils/rigid_utils.py:185-205 (R = sum_ab q_a q_b QTR[a,b])."""
    a, b, c, d = q.unbind(-1)
    rows = [
        torch.stack([a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c], -1),
        torch.stack([2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b], -1),
        torch.stack([2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d], -1),
    ]
    return torch.stack(rows, dim=-2)


def rot_to_quat(R):
    """openfold/utils/rigid_utils.py:208-227: top eigenvector of the symmetric 4x4 K(R)/3."""
    xx, xy, xz = R[..., 0, 0], R[..., 0, 1], R[..., 0, 2]
    yx, yy, yz = R[..., 1, 0], R[..., 1, 1], R[..., 1, 2]
    zx, zy, zz = R[..., 2, 0], R[..., 2, 1], R[..., 2, 2]
    K = torch.stack([
        torch.stack([xx + yy + zz, zy - yz, xz - zx, yx - xy], -1),
        torch.stack([zy - yz, xx - yy - zz, xy + yx, xz + zx], -1),
        torch.stack([xz - zx, xy + yx, yy - xx - zz, yz + zy], -1),
        torch.stack([yx - xy, xz + zx, yz + zy, zz - xx - yy], -1),
    ], dim=-2) / 3.0
    _, vec = torch.linalg.eigh(K)
    return vec[..., -1]


def quat_times_vec(q, v):
    """q (x) (0, v)   openfold/utils/rigid_utils.py:230-275."""
    a, b, c, d = q.unbind(-1)
    x, y, z = v.unbind(-1)
    return torch.stack([
        -b * x - c * y - d * z,
        a * x + c * z - d * y,
        a * y - b * z + d * x,
        a * z + b * y - c * x,
    ], dim=-1)


def rot_apply(R, p):
    """R p for p [..., 3] (openfold/utils/rigid_utils.py:82-106)."""
    return torch.einsum("...ij,...j->...i", R, p)


class Frames:
    """Minimal stand-in for openfold Rigid/Rotation: exactly one of rot/quat is set."""

    def __init__(self, trans, rot=None, quat=None):
        self.trans, self.rot, self.quat = trans, rot, quat

    def rot_mats(self):
        return self.rot if self.rot is not None else quat_to_rot(self.quat)

    def quats(self):
        return self.quat if self.quat is not None else rot_to_quat(self.rot)


def rigid_compose_q_update(fr, upd, mask):
    """openfold/utils/rigid_utils.py:1039-1063 -> :587-616; normalisation at :331-332.
    upd [B,L,6] = (b,c,d | t-vec); mask [B,L,1]."""
    q = fr.quats()
    dq = quat_times_vec(q, upd[..., :3]) * mask
    qn = q + dq
    qn = qn / torch.linalg.norm(qn, dim=-1, keepdim=True)
    dt = rot_apply(fr.rot_mats(), upd[..., 3:]) * mask
    return Frames(fr.trans + dt, quat=qn)


# ----------------------------------------------------------------------------- SO(3) / torus


def so3_hat(w):
    """data/so3_utils.py:285-311."""
    x, y, z = w.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2)


def so3_exp(w, tol=1e-7):
    """rotvec -> rotmat, data/so3_utils.py:143-164 -> :88-140 (Rodrigues; Taylor below tol)."""
    th = torch.linalg.norm(w, dim=-1)[..., None, None]
    K = so3_hat(w)
    th2 = th * th
    small = th.abs() < tol
    s = torch.where(small, 1.0 - th2 / 6.0, torch.sin(th) / th)
    c = torch.where(small, 0.5 - th2 / 24.0, (1.0 - torch.cos(th)) / th2)
    eye = torch.eye(3).expand_as(K)
    return eye + s * K + c * (K @ K)


def so3_log(R):
    """rotmat -> rotvec, data/so3_utils.py:167-254 (+ angle_from_rotmat :257-282)."""
    S = R - R.transpose(-1, -2)
    v = torch.stack([S[..., 2, 1], S[..., 0, 2], S[..., 1, 0]], dim=-1)
    sin_t = torch.linalg.norm(v, dim=-1) / 2.0
    cos_t = (R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2] - 1.0) / 2.0
    th = torch.atan2(sin_t, cos_t)
    m0 = torch.isclose(th, torch.zeros_like(th)).to(th.dtype)
    mpi = torch.isclose(th, torch.full_like(th, math.pi), atol=1e-2).to(th.dtype)
    me = (1 - m0) * (1 - mpi)
    num = m0 / 2.0 + th * me
    den = (1.0 - th ** 2 / 6.0) * m0 + 2.0 * sin_t * me + mpi
    w = v * (num / den)[..., None]
    # theta ~ pi: omega omega^T = (I + R)/2, diagonal clamped at 0
    eye = torch.eye(3).expand_as(R)
    M = (eye + R) / 2.0
    M = M + (torch.relu(M) - M) * eye
    wpi = torch.sqrt(torch.diagonal(M, dim1=-2, dim2=-1))
    row = torch.argmax(torch.linalg.norm(M, dim=-1), dim=-1)
    sel = torch.take_along_dim(M, row[..., None, None], dim=-2).squeeze(-2)
    wpi = wpi * th[..., None] * torch.sign(sel)
    return w + wpi * mpi[..., None]


def calc_rot_vf(mat_t, mat_1):
    """data/so3_utils.py:486-497."""
    return so3_log(mat_t.transpose(-1, -2) @ mat_1)


def geodesic_t(t, mat, base_mat):
    """data/so3_utils.py:500-520:  base * Exp(t * Log(base^T mat))."""
    return base_mat @ so3_exp(t * calc_rot_vf(base_mat, mat))


def tor_logmap(x, y):
    """models_con/torus.py:8-9."""
    return torch.atan2(torch.sin(y - x), torch.cos(y - x))


def tor_geodesic_t(t, angles_1, angles_0):
    """models_con/torus.py:22-26 (expmap :5-6 is a mod 2pi)."""
    return (angles_0 + t * tor_logmap(angles_0, angles_1)) % TWO_PI


# ----------------------------------------------------------------------------- blocks

H, C, PQ, PV = 8, 128, 8, 12


def ipa_forward(sd, p, s, z, fr, mask):
    """InvariantPointAttention.forward, models_con/ipa_pytorch.py:316-484.
    sd/p: state dict and key prefix (e.g. 'ga_encoder.trunk.ipa_0.'); s [B,L,128]; z [B,L,L,64];
    fr Frames; mask [B,L] float."""
    B, L, _ = s.shape
    R, t = fr.rot_mats(), fr.trans
    q = linear(s, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, L, H, C)           # :347
    kv = linear(s, sd[p + "linear_kv.weight"], sd[p + "linear_kv.bias"]).view(B, L, H, 2 * C)     # :348-357
    k, v = kv[..., :C], kv[..., C:]

    def points(name, n):  # :360-387: output split in x|y|z planes, head-major inside a plane
        loc = linear(s, sd[p + name + ".weight"], sd[p + name + ".bias"]).view(B, L, 3, H, n).permute(0, 1, 3, 4, 2)
        return torch.einsum("blij,blhnj->blhni", R, loc) + t[:, :, None, None, :]

    q_pts = points("linear_q_points", PQ)
    kv_pts = points("linear_kv_points", PQ + PV)
    k_pts, v_pts = kv_pts[..., :PQ, :], kv_pts[..., PQ:, :]

    bias = linear(z, sd[p + "linear_b.weight"], sd[p + "linear_b.bias"])                           # :393 [B,L,L,H]
    a = torch.einsum("bihc,bjhc->bhij", q, k) * math.sqrt(1.0 / (3 * C))                            # :399-403
    a = a + math.sqrt(1.0 / 3) * bias.permute(0, 3, 1, 2)                                          # :404
    d2 = ((q_pts[:, :, None] - k_pts[:, None, :]) ** 2).sum(-1)                                    # :407-411 [B,L,L,H,PQ]
    hw = torch.nn.functional.softplus(sd[p + "head_weights"]) * math.sqrt(1.0 / (3 * (PQ * 9.0 / 2)))   # :412-417
    pt = (d2 * hw[None, None, None, :, None]).sum(-1) * (-0.5)                                     # :418-421
    a = a + pt.permute(0, 3, 1, 2)
    sq = mask[:, :, None] * mask[:, None, :]
    a = a + (1e5 * (sq - 1))[:, None]                                                              # :423-430
    a = torch.softmax(a, dim=-1)                                                                   # :431
    o = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, L, H * C)                                 # :437-442
    o_pt = torch.einsum("bhij,bjhpx->bihpx", a, v_pts)                                             # :445-452
    o_pt = torch.einsum("blji,blhpj->blhpi", R, o_pt - t[:, :, None, None, :])                     # :455 invert_apply
    o_norm = torch.sqrt((o_pt ** 2).sum(-1) + 1e-8).reshape(B, L, H * PV)                          # :458-460
    o_pt = o_pt.reshape(B, L, H * PV, 3)
    pair_z = linear(z, sd[p + "down_z.weight"], sd[p + "down_z.bias"])                             # :469
    o_pair = torch.einsum("bhij,bijc->bihc", a, pair_z).reshape(B, L, -1)                          # :470-473
    feats = torch.cat([o, o_pt[..., 0], o_pt[..., 1], o_pt[..., 2], o_norm, o_pair], dim=-1)       # :475
    return linear(feats, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def transformer_encoder_layer(sd, p, x, valid):
    """One post-norm torch.nn.TransformerEncoderLayer (d=128, 4 heads, ff=128, ReLU, dropout 0),
    built at models_con/ga.py:53-62.  valid [B,L] bool: keys with valid=False are masked."""
    B, L, D = x.shape
    nh = 4
    hd = D // nh
    qkv = linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, L, nh, hd).transpose(1, 2)
    k = k.view(B, L, nh, hd).transpose(1, 2)
    v = v.view(B, L, nh, hd).transpose(1, 2)
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    att = att.masked_fill(~valid[:, None, None, :], float("-inf"))
    att = torch.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).reshape(B, L, D)
    y = linear(y, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    x = layer_norm(x + y, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    f = linear(torch.relu(linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
               sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return layer_norm(x + f, sd[p + "norm2.weight"], sd[p + "norm2.bias"])


def node_transition(sd, p, s):
    """StructureModuleTransition.forward, models_con/ipa_pytorch.py:196-206."""
    h = torch.relu(linear(s, sd[p + "linear_1.weight"], sd[p + "linear_1.bias"]))
    h = torch.relu(linear(h, sd[p + "linear_2.weight"], sd[p + "linear_2.bias"]))
    h = linear(h, sd[p + "linear_3.weight"], sd[p + "linear_3.bias"])
    return layer_norm(s + h, sd[p + "ln.weight"], sd[p + "ln.bias"])


def edge_transition(sd, p, node, edge):
    """EdgeTransition.forward, models_con/ipa_pytorch.py:233-248."""
    B, L, _ = node.shape
    e = linear(node, sd[p + "initial_embed.weight"], sd[p + "initial_embed.bias"])
    x = torch.cat([edge, e[:, :, None, :].expand(B, L, L, -1), e[:, None, :, :].expand(B, L, L, -1)], dim=-1)
    h = torch.relu(linear(x, sd[p + "trunk.0.weight"], sd[p + "trunk.0.bias"]))
    h = torch.relu(linear(h, sd[p + "trunk.2.weight"], sd[p + "trunk.2.bias"]))
    y = linear(h + x, sd[p + "final_layer.weight"], sd[p + "final_layer.bias"])
    return layer_norm(y, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])


def mlp3(sd, p, x):
    h = torch.relu(linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    h = torch.relu(linear(h, sd[p + "2.weight"], sd[p + "2.bias"]))
    return linear(h, sd[p + "4.weight"], sd[p + "4.bias"])


def ga_encoder_forward(sd, t, rotmats_t, trans_t, angles_t, seqs_t, node_embed, edge_embed,
                       generate_mask, res_mask, p="ga_encoder.", num_blocks=6, return_node=False, trace=None):
    """GAEncoder.forward, models_con/ga.py:87-127.  t [B,1]; masks are long/float [B,L]."""
    B, L = seqs_t.shape
    m = res_mask.float()
    em = m[:, None, :] * m[:, :, None]
    temb = time_embedding(t[:, 0])[:, None, :].expand(B, L, -1)                                   # :79-85
    aenc = angular_encoding(angles_t, sd[p + "angles_embedder.freq_bands"])                      # :94
    x = torch.cat([node_embed, sd[p + "current_seq_embedder.weight"][seqs_t], temb, aenc], dim=-1)
    s = linear(torch.relu(linear(x, sd[p + "res_feat_mixer.0.weight"], sd[p + "res_feat_mixer.0.bias"])),
               sd[p + "res_feat_mixer.2.weight"], sd[p + "res_feat_mixer.2.bias"]) * m[..., None]   # :94-95
    fr = Frames(trans_t, rot=rotmats_t.float())                                                    # :96
    z = edge_embed
    valid = m > 0
    rec = (lambda name, x: trace.append((name, x.clone()))) if trace is not None else (lambda name, x: None)
    rec("mix", s)
    for b in range(num_blocks):
        tp = f"{p}trunk."
        ipa = ipa_forward(sd, f"{tp}ipa_{b}.", s, z, fr, m) * m[..., None]                          # :98-103
        rec(f"ipa_{b}", ipa)
        s = layer_norm(s + ipa, sd[f"{tp}ipa_ln_{b}.weight"], sd[f"{tp}ipa_ln_{b}.bias"])           # :104
        y = s
        for l in range(2):                                                                          # :105-106
            y = transformer_encoder_layer(sd, f"{tp}seq_tfmr_{b}.layers.{l}.", y, valid)
        rec(f"tfmr_{b}", y)
        s = s + linear(y, sd[f"{tp}post_tfmr_{b}.weight"], sd[f"{tp}post_tfmr_{b}.bias"])           # :107
        s = node_transition(sd, f"{tp}node_transition_{b}.", s) * m[..., None]                      # :108-109
        rec(f"node_{b}", s)
        upd = linear(s * m[..., None], sd[f"{tp}bb_update_{b}.linear.weight"], sd[f"{tp}bb_update_{b}.linear.bias"])
        fr = rigid_compose_q_update(fr, upd, m[..., None])                                          # :112-113
        rec(f"rot_{b}", fr.rot_mats())
        rec(f"trans_{b}", fr.trans)
        if b < num_blocks - 1:                                                                      # :115-118
            z = edge_transition(sd, f"{tp}edge_transition_{b}.", s, z) * em[..., None]
            rec(f"z_{b}", z)
    pred_trans = fr.trans
    pred_rot = fr.rot_mats()
    logits = mlp3(sd, p + "seq_net.", s)                                                           # :123
    angles = mlp3(sd, p + "angle_net.", s) % TWO_PI                                                # :124-125
    if return_node:
        return pred_rot, pred_trans, angles, logits, s, z
    return pred_rot, pred_trans, angles, logits


# ----------------------------------------------------------------------------- Euler step / loop


def denoise_postprocess(pred, gt, gen, u, torsions_mask):
    """models_con/flow_model.py:291-303: mask to peptide rows, draw residue types, mask torsions.
    pred = (R, x, ang, logits); gt = (R1, x1, ang1, seq1); gen bool [B,L]; u uniforms [B,L]."""
    R, x, ang, logits = pred
    R1, x1, a1, s1 = gt
    R = torch.where(gen[..., None, None], R, R1)
    x = torch.where(gen[..., None], x, x1)
    ang = torch.where(gen[..., None], ang, a1)
    seq = categorical_from_uniform(torch.softmax(logits, dim=-1), u)
    seq = torch.where(gen, seq, s1)
    ang = torch.where(torsions_mask[seq].bool(), ang, torch.zeros_like(ang))
    return R, x, ang, seq


def euler_update(state, clean, gt, noise0, gen, d_t, u, torsions_mask, K=20, k=5.0):
    """models_con/flow_model.py:316-333.  state = (R_t, x_t, ang_t, seq_t, simplex_t);
    clean = (R^, x^, a^, s^) after denoise_postprocess; noise0 = (x_0, simplex_0)."""
    R_t, x_t, a_t, s_t, sx_t = state
    Rh, xh, ah, sh = clean
    R1, x1, a1, s1 = gt
    x0, sx0 = noise0
    x2 = torch.where(gen[..., None], x_t + (xh - x0) * d_t, x1)                                    # :318-320
    R2 = torch.where(gen[..., None, None], geodesic_t(d_t * 10, Rh, R_t), R1)      # :322-323
    a2 = torch.where(gen[..., None], tor_geodesic_t(d_t, ah, a_t), a1)                             # :325-326
    sx2 = sx_t + (seq_to_simplex(sh, K, k) - sx0) * d_t                                            # :328
    s2 = categorical_from_uniform(torch.softmax(sx2, dim=-1), u)                                   # :329
    s2 = torch.where(gen, s2, s1)
    a2 = torch.where(torsions_mask[s2].bool(), a2, torch.zeros_like(a2))                           # :332-333
    return R2, x2, a2, s2, sx2


def sample_loop(sd, enc, noise, uniforms, gen, res_mask, num_steps, torsions_mask, K=20, k=5.0):
    """FlowModel.sample after encode + noise init, models_con/flow_model.py:276-374.
    enc = dict(rotmats_1, trans_1, angles_1, seqs_1, node_embed, edge_embed);
    noise = dict(rotmats_0, trans_0, angles_0, seqs_0, seqs_0_simplex) (already masked to gt on
    context rows); uniforms [num_steps, 2, B, L] (draw 0: s^, draw 1: s_{t2}).
    Returns the clean trajectory (list of dicts like the reference's clean_traj)."""
    R1, x1, a1, s1 = enc["rotmats_1"], enc["trans_1"], enc["angles_1"], enc["seqs_1"]
    gt = (R1, x1, a1, s1)
    state = (noise["rotmats_0"], noise["trans_0"], noise["angles_0"], noise["seqs_0"], noise["seqs_0_simplex"])
    noise0 = (noise["trans_0"], noise["seqs_0_simplex"])
    ts = torch.linspace(1.0e-2, 1.0, num_steps)
    B = s1.shape[0]
    traj = []
    gl, rl = gen.long(), res_mask.long()
    for n in range(num_steps):
        t = torch.ones(B, 1) * ts[n]
        pred = ga_encoder_forward(sd, t, state[0], state[1], state[2], state[3], enc["node_embed"],
                                  enc["edge_embed"], gl, rl)
        clean = denoise_postprocess(pred, gt, gen, uniforms[n, 0], torsions_mask)
        traj.append({"rotmats": clean[0], "trans": clean[1], "angles": clean[2], "seqs": clean[3],
                     "seqs_simplex": seq_to_simplex(clean[3], K, k),
                     "rotmats_1": R1, "trans_1": x1, "angles_1": a1, "seqs_1": s1})
        if n == num_steps - 1:
            break
        d_t = ts[n + 1] - ts[n]          # 0-dim fp32, as in the reference (:316)
        state = euler_update(state, clean, gt, noise0, gen, d_t, uniforms[n, 1], torsions_mask, K, k)
    return traj


# ----------------------------------------------------------------------------- embedders (once per sample)


def construct_3d_basis(center, p1, p2):
    """pepflow/modules/common/geometry.py:89-111 (columns e1,e2,e3; eps 1e-6 in the norms)."""
    v1 = p1 - center
    e1 = v1 / (torch.linalg.norm(v1, dim=-1, keepdim=True) + 1e-6)
    v2 = p2 - center
    u2 = v2 - (e1 * v2).sum(-1, keepdim=True) * e1
    e2 = u2 / (torch.linalg.norm(u2, dim=-1, keepdim=True) + 1e-6)
    e3 = torch.cross(e1, e2, dim=-1)
    return torch.stack([e1, e2, e3], dim=-1)


def dihedral(p0, p1, p2, p3):
    """pepflow/modules/common/geometry.py:296-313."""
    v0, v1, v2 = p2 - p1, p0 - p1, p3 - p2
    u1 = torch.cross(v0, v1, dim=-1)
    n1 = u1 / torch.linalg.norm(u1, dim=-1, keepdim=True)
    u2 = torch.cross(v0, v2, dim=-1)
    n2 = u2 / torch.linalg.norm(u2, dim=-1, keepdim=True)
    sgn = torch.sign((torch.cross(v1, v2, dim=-1) * v0).sum(-1))
    return torch.nan_to_num(sgn * torch.acos((n1 * n2).sum(-1).clamp(-0.999999, 0.999999)))


def _mlp(sd, p, x, idxs):
    for n, i in enumerate(idxs):
        x = linear(x, sd[f"{p}{i}.weight"], sd[f"{p}{i}.bias"])
        if n < len(idxs) - 1:
            x = torch.relu(x)
    return x


def node_embedder(sd, aa, res_nb, chain_nb, pos, mask_atoms, structure_mask, sequence_mask, p="node_embedder."):
    """NodeEmbedder.forward, models_con/node.py:35-105 (22 aa slots x 15 atoms x 3 placed coordinates)."""
    N, L = aa.shape
    A = 15
    mask_res = mask_atoms[:, :, 1]
    if sequence_mask is not None:
        aa = torch.where(sequence_mask, aa, torch.full_like(aa, 20))
    aa_feat = sd[p + "aatype_embed.weight"][aa]
    R = construct_3d_basis(pos[:, :, 1], pos[:, :, 2], pos[:, :, 0])
    t = pos[:, :, 1]
    crd = torch.einsum("nlji,nlaj->nlai", R, pos - t[:, :, None, :])           # geometry.py:136-155
    crd = torch.where(mask_atoms[..., None], crd, torch.zeros_like(crd))
    place = torch.nn.functional.one_hot(aa, 22).to(crd.dtype)                   # node.py:66-71
    crd_feat = (place[:, :, :, None, None] * crd[:, :, None, :, :]).reshape(N, L, 22 * A * 3)
    if structure_mask is not None:
        crd_feat = crd_feat * structure_mask[:, :, None]
    # backbone dihedrals  geometry.py:355-390; terminus flags topology.py:5-24
    consec = ((res_nb[:, 1:] - res_nb[:, :-1]).abs() == 1) & (chain_nb[:, 1:] == chain_nb[:, :-1]) & mask_res[:, :-1]
    pad = torch.nn.functional.pad
    n_term = pad(~consec, (1, 0), value=True)
    c_term = pad(~consec, (0, 1), value=True)
    pN, pCA, pC = pos[:, :, 0], pos[:, :, 1], pos[:, :, 2]
    omega = pad(dihedral(pCA[:, :-1], pC[:, :-1], pN[:, 1:], pCA[:, 1:]), (1, 0))
    phi = pad(dihedral(pC[:, :-1], pN[:, 1:], pCA[:, 1:], pC[:, 1:]), (1, 0))
    psi = pad(dihedral(pN[:, :-1], pCA[:, :-1], pC[:, :-1], pN[:, 1:]), (0, 1))
    dmask = torch.stack([~n_term, ~n_term, ~c_term], dim=-1)
    dih = torch.stack([omega, phi, psi], dim=-1) * dmask
    dfeat = angular_encoding(dih[..., None], sd[p + "dihed_embed.freq_bands"]) * dmask[..., None]
    dfeat = dfeat.reshape(N, L, -1)
    if structure_mask is not None:
        dm = structure_mask & torch.roll(structure_mask, 1, 1) & torch.roll(structure_mask, -1, 1)
        dfeat = dfeat * dm[:, :, None]
    out = _mlp(sd, p + "mlp.", torch.cat([aa_feat, crd_feat, dfeat], dim=-1), [0, 2, 4, 6])
    return out * mask_res[:, :, None]


def edge_embedder(sd, aa, res_nb, chain_nb, pos, mask_atoms, structure_mask, sequence_mask, p="edge_embedder."):
    """EdgeEmbedder.forward, models_con/edge.py:39-112."""
    N, L = aa.shape
    mask_res = mask_atoms[:, :, 1]
    mask_pair = mask_res[:, :, None] * mask_res[:, None, :]
    psm = structure_mask[:, :, None] * structure_mask[:, None, :] if structure_mask is not None else None
    if sequence_mask is not None:
        aa = torch.where(sequence_mask, aa, torch.full_like(aa, 20))
    aa_pair = aa[:, :, None] * 22 + aa[:, None, :]
    f_aa = sd[p + "aa_pair_embed.weight"][aa_pair]
    same = chain_nb[:, :, None] == chain_nb[:, None, :]
    rel = torch.clamp(res_nb[:, :, None] - res_nb[:, None, :], -32, 32)
    f_rel = sd[p + "relpos_embed.weight"][rel + 32] * same[..., None]
    d = torch.linalg.norm(pos[:, :, None, :, None] - pos[:, None, :, None, :], dim=-1).reshape(N, L, L, -1) / 10
    c = torch.nn.functional.softplus(sd[p + "aapair_to_distcoef.weight"][aa_pair])
    g = torch.exp(-1 * c * d ** 2)
    map_ = (mask_atoms[:, :, None, :, None] * mask_atoms[:, None, :, None, :]).reshape(N, L, L, -1)
    f_d = torch.relu(_mlp(sd, p + "distance_embed.", g * map_, [0, 2]))
    if psm is not None:
        f_d = f_d * psm[..., None]
    pN, pCA, pC = pos[:, :, 0], pos[:, :, 1], pos[:, :, 2]
    ex = lambda x, dim: x.unsqueeze(dim).expand(N, L, L, 3)
    phi = dihedral(ex(pC, 2), ex(pN, 1), ex(pCA, 1), ex(pC, 1))                 # geometry.py:393-418
    psi = dihedral(ex(pN, 2), ex(pCA, 2), ex(pC, 2), ex(pN, 1))
    f_h = angular_encoding(torch.stack([phi, psi], dim=-1), sd[p + "dihedral_embed.freq_bands"])
    if psm is not None:
        f_h = f_h * psm[..., None]
    out = _mlp(sd, p + "out_mlp.", torch.cat([f_aa, f_rel, f_d, f_h], dim=-1), [0, 2, 4])
    return out * mask_pair[..., None]


def encode(sd, batch):
    """FlowModel.encode, models_con/flow_model.py:75-93 (sample_structure = sample_sequence = True)."""
    pos = batch["pos_heavyatom"]
    rot1 = construct_3d_basis(pos[:, :, 1], pos[:, :, 2], pos[:, :, 0])
    ctx = batch["mask_heavyatom"][:, :, 1] & ~batch["generate_mask"]
    args = (batch["aa"], batch["res_nb"], batch["chain_nb"], pos, batch["mask_heavyatom"], ctx, ctx)
    return {"rotmats_1": rot1, "trans_1": pos[:, :, 1], "angles_1": batch["torsion_angle"], "seqs_1": batch["aa"],
            "node_embed": node_embedder(sd, *args), "edge_embed": edge_embedder(sd, *args)}


def zero_center_part(pos, gen_mask, res_mask):
    """models_con/flow_model.py:95-106."""
    g = gen_mask.float()
    center = (pos * g[..., None]).sum(1) / (g.sum(-1, keepdim=True) + 1e-8)
    return (pos - center[:, None]) * res_mask.float()[..., None], center[:, None]


# ----------------------------------------------------------------------------- post-sampling reconstruction
# SURVEY.md section 8f rank 3.  `tables` = the dict of pepflowww_b200/data/restype_rigid_tables.npz as CPU tensors
# (rigid_rot [21,8,3,3], rigid_trans [21,8,3], atom_group [21,14], atom_pos [21,14,3], bb_coords [21,3,3],
# bb_oxygen [21,3]) - the reference's pepflow/modules/protein/constants.py:665-668,878-879.


def _rot_x(angle):
    """[[1,0,0],[0,c,-s],[0,s,c]] per angle (models_con/torsion.py:68-96, geometry.py:471-479)."""
    s, c = torch.sin(angle), torch.cos(angle)
    R = torch.zeros(angle.shape + (3, 3), dtype=angle.dtype)
    R[..., 0, 0] = 1.0
    R[..., 1, 1] = c
    R[..., 1, 2] = -s
    R[..., 2, 1] = s
    R[..., 2, 2] = c
    return R


def _compose(R1, t1, R2, t2):
    """(R1, t1) o (R2, t2) (pepflow/modules/common/geometry.py:175-180)."""
    return R1 @ R2, (R1 @ t2[..., None])[..., 0] + t1


def full_atom_reconstruction(tables, R_bb, t_bb, angles, aa):
    """models_con/torsion.py:140-226: (pos14 [B,L,14,3], R [B,L,6,3,3], t [B,L,6,3]); compose_chain folds from the
    right (geometry.py:183-189), so each torsion frame is parent o (group o Rx)."""
    rr, rt = tables["rigid_rot"][aa], tables["rigid_trans"][aa]         # [B,L,8,3,3], [B,L,8,3]
    zero = torch.zeros_like(t_bb)
    frames = [(R_bb, t_bb)]
    for f in range(1, 6):                                                # psi, chi1..chi4 = rigid groups 3..7
        parent = frames[0] if f <= 2 else frames[f - 1]
        inner = _compose(rr[:, :, f + 2], rt[:, :, f + 2], _rot_x(angles[:, :, f - 1]), zero)
        frames.append(_compose(parent[0], parent[1], inner[0], inner[1]))
    group = tables["atom_group"][aa].long()                              # [B,L,14]
    fidx = torch.where(group < 3, torch.zeros_like(group), group - 2)
    R_all = torch.stack([f[0] for f in frames], dim=2)
    t_all = torch.stack([f[1] for f in frames], dim=2)
    R_atom = torch.gather(R_all, 2, fidx[..., None, None].expand(-1, -1, -1, 3, 3))
    t_atom = torch.gather(t_all, 2, fidx[..., None].expand(-1, -1, -1, 3))
    pos14 = (R_atom @ tables["atom_pos"][aa][..., None])[..., 0] + t_atom
    return pos14, R_all, t_all


def reconstruct_backbone(tables, R, t, aa, chain_nb, res_nb, mask):
    """pepflow/modules/common/geometry.py:446-489: N, CA, C from the backbone frame, O from the psi frame with psi
    measured on the rebuilt backbone (:355-390; terminus flags pepflow/modules/common/topology.py:5-24)."""
    aa = aa.clamp(0, 20)
    bb = (R[:, :, None] @ tables["bb_coords"][aa][..., None])[..., 0] + t[:, :, None]     # [B,L,3,3]
    consec = ((res_nb[:, 1:] - res_nb[:, :-1]).abs() == 1) & (chain_nb[:, 1:] == chain_nb[:, :-1]) & mask[:, :-1].bool()
    psi = dihedral(bb[:, :-1, 0], bb[:, :-1, 1], bb[:, :-1, 2], bb[:, 1:, 0]) * consec
    psi = torch.cat([psi, torch.zeros_like(psi[:, :1])], dim=1)
    R_psi, t_psi = _compose(R, t, _rot_x(psi), torch.zeros_like(t))
    O = (R_psi @ tables["bb_oxygen"][aa][..., None])[..., 0] + t_psi
    return torch.cat([bb, O[:, :, None]], dim=2)


# ----------------------------------------------------------------------------- training-loss arithmetic
# SURVEY.md section 8f rank 4.


def flow_losses(sd, batch, noise, torsions_mask, cfg_min_clip=0.9, sigma=1.0, K=20, k=5.0):
    """FlowModel.forward, models_con/flow_model.py:111-227, with the corruption injected instead of drawn:
    noise = dict(t [B,1] (already in [min_t, 1 - min_t]), trans_0 [B,L,3] raw normal, rotmats_0, angles_0,
    seqs_0_simplex (already scaled by k), u_t, u_pred [B,L] uniforms of the two categorical draws)."""
    enc = encode(sd, batch)
    gen = batch["generate_mask"]
    gm, rm = gen.float(), batch["res_mask"].float()
    R1, x1, a1, s1 = enc["rotmats_1"], enc["trans_1"], enc["angles_1"], enc["seqs_1"]
    t = noise["t"]
    x0c, _ = zero_center_part(noise["trans_0"] * sigma, gen, batch["res_mask"])                      # :129-130
    x_t = torch.where(gen[..., None], (1 - t[..., None]) * x0c + t[..., None] * x1, x1)              # :131-132
    R_t = torch.where(gen[..., None, None], geodesic_t(t[..., None], R1, noise["rotmats_0"]), R1)    # :134-136
    a_t = torch.where(gen[..., None], tor_geodesic_t(t[..., None], a1, noise["angles_0"]), a1)       # :138-140
    sx1 = seq_to_simplex(s1, K, k)
    sx_t = torch.where(gen[..., None], (1 - t[..., None]) * noise["seqs_0_simplex"] + t[..., None] * sx1, sx1)
    s_t = torch.where(gen, categorical_from_uniform(torch.softmax(sx_t, -1), noise["u_t"]), s1)     # :147-152
    pR, px, pa, logits = ga_encoder_forward(sd, t, R_t, x_t, a_t, s_t, enc["node_embed"], enc["edge_embed"],
                                            gen.long(), batch["res_mask"].long())                   # :159
    ps = categorical_from_uniform(torch.softmax(logits, -1), noise["u_pred"])
    ps = torch.where(gen, ps, s1.clamp(0, 19))                                                       # :160-161
    scale = 1 / (1 - torch.clamp(t[..., None], max=cfg_min_clip))                                    # :165
    gsum = gm.sum(-1) + 1e-8
    per = lambda x: (x / gsum).mean()
    trans_loss = per(((px - x1) ** 2 * gm[..., None]).sum((-1, -2)))                                 # :168-169
    rot_loss = per((((calc_rot_vf(R_t, R1) - calc_rot_vf(R_t, pR)) * scale) ** 2 * gm[..., None]).sum((-1, -2)))
    ideal = torch.tensor([[-0.525, 1.363, 0.0], [0.0, 0.0, 0.0], [1.526, 0.0, 0.0]])               # N, CA, C (:178-179)
    bb = lambda R, x: torch.einsum("blij,aj->blai", R, ideal) + x[:, :, None]
    bb_loss = per(((bb(R1, x1) - bb(pR, px)) ** 2 * gm[..., None, None]).sum((-1, -2, -3)))          # :183-187
    ce = torch.nn.functional.cross_entropy(logits.reshape(-1, K), s1.clamp(0, 19).reshape(-1), reduction="none")
    seqs_loss = per((ce.view(logits.shape[:-1]) * gm).sum(-1))                                       # :191-193
    aml = torsions_mask[ps]
    aml = torch.cat([aml, aml], -1).bool() & gen[..., None]                                          # :199-202
    asum = aml.sum((-1, -2)) + 1e-8
    vec = lambda x: torch.cat([torch.sin(x), torch.cos(x)], -1)
    angle_loss = ((((vec(tor_logmap(a_t, a1)) - vec(tor_logmap(a_t, pa))) * scale) ** 2 * aml).sum((-1, -2)) / asum).mean()
    torsion_loss = (((vec(pa) - vec(a1)) ** 2 * aml).sum((-1, -2)) / asum).mean()                    # :213-218
    return {"trans_loss": trans_loss, "rot_loss": rot_loss, "bb_atom_loss": bb_loss, "seqs_loss": seqs_loss,
            "angle_loss": angle_loss, "torsion_loss": torsion_loss}


def _torsion_raw(p0, p1, p2, p3):
    """models_con/torsion.py:13-29 (_get_torsion): like dihedral() but degenerate geometry stays NaN."""
    v0, v1, v2 = p2 - p1, p0 - p1, p3 - p2
    u1 = torch.cross(v0, v1, dim=-1)
    n1 = u1 / torch.linalg.norm(u1, dim=-1, keepdim=True)
    u2 = torch.cross(v0, v2, dim=-1)
    n2 = u2 / torch.linalg.norm(u2, dim=-1, keepdim=True)
    sgn = torch.sign((torch.cross(v1, v2, dim=-1) * v0).sum(-1))
    return sgn * torch.acos((n1 * n2).sum(-1).clamp(-0.999999, 0.999999))


def torsion_angles(tables, pos, aa):
    """get_torsion_angle, models_con/torsion.py:31-66, vectorised over residues: pos [..., A, 3], aa [...] ->
    (torsion [..., 5] in [0, 2 pi), mask [..., 5]).  psi = torsion(N, CA, C, O); chi_i from tables['chi_atoms']."""
    known = (aa >= 0) & (aa < 20)
    idx = tables["chi_atoms"][aa.clamp(0, 20)].long()                    # [..., 4, 4]
    have = (idx[..., 0] >= 0) & known[..., None]
    g = lambda j: torch.gather(pos, -2, idx[..., j].clamp(min=0)[..., None].expand(idx.shape[:-1] + (3,)))
    chi = _torsion_raw(g(0), g(1), g(2), g(3))
    chi = torch.where(have, chi, torch.full_like(chi, float("inf")))
    psi = _torsion_raw(pos[..., 0, :], pos[..., 1, :], pos[..., 2, :], pos[..., 3, :])
    tor = torch.cat([psi[..., None], chi], dim=-1)
    mask = torch.isfinite(tor) & known[..., None]
    tor = torch.where(mask, tor, torch.zeros_like(tor)) % TWO_PI
    return tor, mask
