"""Tensor-level wrappers over the C ABI (one function per entry point of include/pepflow_b200.h).

All inputs must be contiguous CUDA tensors; outputs are freshly allocated torch tensors on the same
device.  Kernels are enqueued on torch's current stream.
"""
import torch

from . import _lib
from ._lib import check, ptr, stream

F32, I64, U8 = torch.float32, torch.int64, torch.uint8


def _c(t, dtype=F32):
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


_linear_ws = {}


def linear(x, w, b=None, residual=None, rowmask=None, act=0, out=None):
    """y = act(x W^T + b) (+ residual) (* rowmask[:, None]).  x [..., K], w [N, K]."""
    lib = _lib.lib_for(x.device)
    K = x.shape[-1]
    N = w.shape[0]
    x2 = _c(x).reshape(-1, K)
    M = x2.shape[0]
    y = out if out is not None else torch.empty(M, N, device=x.device, dtype=F32)
    res = _c(residual).reshape(M, N) if residual is not None else None
    rm = _c(rowmask).reshape(M) if rowmask is not None else None
    # scratch for the tcgen05 GEMM's packed weight tiles (used when "gemm_impl" = 2 and K = 128)
    nbytes = int(lib.pf_linear_workspace_bytes(N))
    ws = _linear_ws.get(x.device)
    if ws is None or ws.numel() < nbytes:
        ws = _linear_ws[x.device] = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    check(lib.pf_linear_ws(ptr(x2), ptr(_c(w)), ptr(_c(b), allow_none=True), ptr(res, allow_none=True),
                           ptr(rm, allow_none=True), ptr(y), M, K, N, int(act), ptr(ws, dtype=torch.uint8), nbytes,
                           stream()))
    return y.reshape(*x.shape[:-1], N)


def add_layernorm(a, b, gamma, beta, rowmask=None):
    lib = _lib.lib_for(a.device)
    N = a.shape[-1]
    a2 = _c(a).reshape(-1, N)
    M = a2.shape[0]
    b2 = _c(b).reshape(M, N) if b is not None else None
    rm = _c(rowmask).reshape(M) if rowmask is not None else None
    y = torch.empty_like(a2)
    check(lib.pf_add_layernorm(ptr(a2), ptr(b2, allow_none=True), ptr(_c(gamma)), ptr(_c(beta)),
                               ptr(rm, allow_none=True), ptr(y), M, N, stream()))
    return y.reshape(a.shape)


def mix_features(node_embed, seq_emb, seqs, t, time_freqs, angles, ang_freqs):
    lib = _lib.lib_for(node_embed.device)
    B, L = seqs.shape
    x = torch.empty(B, L, 629, device=node_embed.device, dtype=F32)
    check(lib.pf_mix_features(ptr(_c(node_embed)), ptr(_c(seq_emb)), ptr(_c(seqs, I64), I64), ptr(_c(t).reshape(B)),
                              ptr(_c(time_freqs)), ptr(_c(angles)), ptr(_c(ang_freqs)), ptr(x), B, L, stream()))
    return x


def ipa_points(proj, rot, trans):
    lib = _lib.lib_for(proj.device)
    B, L = proj.shape[:2]
    pts = torch.empty(B, L, 8, 28, 3, device=proj.device, dtype=F32)
    check(lib.pf_ipa_points(ptr(_c(proj)), ptr(_c(rot).reshape(B, L, 9)), ptr(_c(trans)), ptr(pts), B, L, stream()))
    return pts


def ipa_attention(proj, pts, z, w_b, b_b, w_dz, b_dz, head_w, rot, trans, mask):
    lib = _lib.lib_for(proj.device)
    B, L = proj.shape[:2]
    feats = torch.empty(B, L, 1536, device=proj.device, dtype=F32)
    nbytes = lib.pf_ipa_attention_workspace_bytes(B, L)
    ws = torch.empty(nbytes, device=proj.device, dtype=U8)
    check(lib.pf_ipa_attention(ptr(_c(proj)), ptr(_c(pts)), ptr(_c(z)), ptr(_c(w_b)), ptr(_c(b_b)), ptr(_c(w_dz)),
                               ptr(_c(b_dz)), ptr(_c(head_w)), ptr(_c(rot).reshape(B, L, 9)), ptr(_c(trans)),
                               ptr(_c(mask)), ptr(feats), ptr(ws, U8), nbytes, B, L, stream()))
    return feats


def seq_attention(qkv, mask):
    lib = _lib.lib_for(qkv.device)
    B, L = qkv.shape[:2]
    ctx = torch.empty(B, L, 128, device=qkv.device, dtype=F32)
    check(lib.pf_seq_attention(ptr(_c(qkv)), ptr(_c(mask)), ptr(ctx), B, L, stream()))
    return ctx


def rigid_update(quat, rot, trans, upd, mask):
    """Exactly one of quat / rot is given.  Returns (quat', rot', trans')."""
    lib = _lib.lib_for(trans.device)
    n = trans.numel() // 3
    dev = trans.device
    q_o = torch.empty(*trans.shape[:-1], 4, device=dev, dtype=F32)
    r_o = torch.empty(*trans.shape[:-1], 3, 3, device=dev, dtype=F32)
    t_o = torch.empty_like(trans, dtype=F32)
    check(lib.pf_rigid_update(ptr(_c(quat), allow_none=True), ptr(_c(rot), allow_none=True), ptr(_c(trans)),
                              ptr(_c(upd)), ptr(_c(mask)), ptr(q_o), ptr(r_o), ptr(t_o), n, stream()))
    return q_o, r_o, t_o


def edge_transition(s, z, w_init, b_init, w1, b1, w2, b2, wf, bf, ln_g, ln_b, mask, out=None):
    lib = _lib.lib_for(z.device)
    B, L = s.shape[:2]
    nbytes = lib.pf_edge_transition_workspace_bytes(B, L)
    ws = torch.empty(nbytes, device=z.device, dtype=U8)
    z_out = out if out is not None else torch.empty_like(z)
    check(lib.pf_edge_transition(ptr(_c(s)), ptr(_c(z)), ptr(_c(w_init)), ptr(_c(b_init)), ptr(_c(w1)), ptr(_c(b1)),
                                 ptr(_c(w2)), ptr(_c(b2)), ptr(_c(wf)), ptr(_c(bf)), ptr(_c(ln_g)), ptr(_c(ln_b)),
                                 ptr(_c(mask)), ptr(z_out), ptr(ws, U8), nbytes, B, L, stream()))
    return z_out


def mod_2pi(x):
    lib = _lib.lib_for(x.device)
    x = _c(x)
    y = torch.empty_like(x)
    check(lib.pf_mod_2pi(ptr(x), ptr(y), x.numel(), stream()))
    return y


def quat_to_rot(q):
    lib = _lib.lib_for(q.device)
    q = _c(q)
    r = torch.empty(*q.shape[:-1], 3, 3, device=q.device, dtype=F32)
    check(lib.pf_quat_to_rot(ptr(q), ptr(r), q.numel() // 4, stream()))
    return r


def so3_log(rot):
    lib = _lib.lib_for(rot.device)
    rot = _c(rot)
    w = torch.empty(*rot.shape[:-2], 3, device=rot.device, dtype=F32)
    check(lib.pf_so3_log(ptr(rot), ptr(w), rot.numel() // 9, stream()))
    return w


def so3_exp(w):
    lib = _lib.lib_for(w.device)
    w = _c(w)
    r = torch.empty(*w.shape[:-1], 3, 3, device=w.device, dtype=F32)
    check(lib.pf_so3_exp(ptr(w), ptr(r), w.numel() // 3, stream()))
    return r


def so3_geodesic(t, mat, base):
    """t broadcastable over the leading dims of mat: one value per `group` consecutive matrices."""
    lib = _lib.lib_for(mat.device)
    mat, base = _c(mat), _c(base)
    n = mat.numel() // 9
    t = _c(torch.as_tensor(t, device=mat.device)).reshape(-1)
    if n % max(t.numel(), 1) != 0:
        raise ValueError(f"Incompatible shapes: t={tuple(t.shape)}, mat={tuple(mat.shape)}")
    out = torch.empty_like(mat)
    check(lib.pf_so3_geodesic(ptr(t), ptr(mat), ptr(base), ptr(out), n, n // t.numel(), stream()))
    return out


def tor_geodesic(t, ang1, ang0):
    lib = _lib.lib_for(ang1.device)
    ang1, ang0 = _c(ang1), _c(ang0)
    d = ang1.shape[-1]
    n = ang1.numel() // d
    t = _c(torch.as_tensor(t, device=ang1.device)).reshape(-1)
    out = torch.empty_like(ang1)
    check(lib.pf_tor_geodesic(ptr(t), ptr(ang1), ptr(ang0), ptr(out), n, n // t.numel(), d, stream()))
    return out


def denoise_post(pred, gt, gen_mask_u8, tmask, uniforms, seed, counter, clean, simplex_k):
    """pred = (rot, trans, ang, logits); gt = (rot1, trans1, ang1, seq1); clean = output tensors
    (rot, trans, ang, seq, simplex) written in place."""
    lib = _lib.lib_for(pred[0].device)
    n = gt[3].numel()
    check(lib.pf_denoise_post(ptr(pred[0]), ptr(pred[1]), ptr(pred[2]), ptr(pred[3]), ptr(gt[0]), ptr(gt[1]),
                              ptr(gt[2]), ptr(gt[3], I64), ptr(gen_mask_u8, U8), ptr(tmask),
                              ptr(uniforms, allow_none=True), int(seed), int(counter), ptr(clean[0]), ptr(clean[1]),
                              ptr(clean[2]), ptr(clean[3], I64), ptr(clean[4]), n, float(simplex_k), stream()))


def euler_step(state, clean, noise0, gt, gen_mask_u8, tmask, uniforms, seed, counter, d_t, out, simplex_k):
    """state = (rot, trans, ang, simplex); clean = (rot, trans, ang, seq); noise0 = (trans0, simplex0);
    gt = (rot1, trans1, ang1, seq1); out = (rot, trans, ang, seq, simplex) (may alias state)."""
    lib = _lib.lib_for(state[0].device)
    n = gt[3].numel()
    check(lib.pf_euler_step(ptr(state[0]), ptr(state[1]), ptr(state[2]), ptr(state[3]), ptr(clean[0]), ptr(clean[1]),
                            ptr(clean[2]), ptr(clean[3], I64), ptr(noise0[0]), ptr(noise0[1]), ptr(gt[0]), ptr(gt[1]),
                            ptr(gt[2]), ptr(gt[3], I64), ptr(gen_mask_u8, U8), ptr(tmask),
                            ptr(uniforms, allow_none=True), int(seed), int(counter), float(d_t), ptr(out[0]),
                            ptr(out[1]), ptr(out[2]), ptr(out[3], I64), ptr(out[4]), n, float(simplex_k), stream()))


def ga_encoder_workspace_bytes(B, L):
    return int(_lib.load().pf_ga_encoder_workspace_bytes(B, L))


def ga_encoder_forward(weights_struct, t, rot_t, trans_t, angles_t, seqs_t, node_embed, edge_embed, res_mask,
                       workspace, out=None, node_out=None):
    """Composite GAEncoder.forward.  All tensors contiguous fp32 CUDA (seqs int64).  Returns
    (pred_rot [B,L,3,3], pred_trans, pred_angles, logits)."""
    lib = _lib.lib_for(edge_embed.device)
    B, L = seqs_t.shape
    dev = edge_embed.device
    if out is None:
        out = (torch.empty(B, L, 3, 3, device=dev, dtype=F32), torch.empty(B, L, 3, device=dev, dtype=F32),
               torch.empty(B, L, 5, device=dev, dtype=F32), torch.empty(B, L, 20, device=dev, dtype=F32))
    import ctypes
    check(lib.pf_ga_encoder_forward(ctypes.addressof(weights_struct), ptr(t), ptr(rot_t), ptr(trans_t),
                                    ptr(angles_t), ptr(seqs_t, I64), ptr(node_embed), ptr(edge_embed), ptr(res_mask),
                                    ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]),
                                    ptr(node_out, allow_none=True), ptr(workspace, U8), workspace.numel(), B, L,
                                    stream()))
    return out


def seq_transformer_forward(weights_struct, block, x, res_mask, workspace):
    """The two post-norm encoder layers of block `block` (ga.py:105-106) through the composite's own kernels."""
    import ctypes
    lib = _lib.lib_for(x.device)
    B, L = x.shape[:2]
    y = torch.empty(B, L, 128, device=x.device, dtype=F32)
    check(lib.pf_seq_transformer_forward(ctypes.addressof(weights_struct), int(block), ptr(_c(x)), ptr(_c(res_mask)),
                                         ptr(y), ptr(workspace, U8), workspace.numel(), B, L, stream(x.device)))
    return y


def zero_center(pos, gen_mask, res_mask):
    """FlowModel.zero_center_part (flow_model.py:95-106) as one kernel: returns (centered * res_mask, center [B,1,3])."""
    lib = _lib.lib_for(pos.device)
    B, L = pos.shape[:2]
    out = _c(pos).clone()
    center = torch.empty(B, 1, 3, device=pos.device, dtype=F32)
    check(lib.pf_zero_center(ptr(out), ptr(_c(gen_mask, torch.bool).view(U8), U8), ptr(_c(res_mask)), ptr(center), B, L,
                             stream(pos.device)))
    return out, center


def sampler_step(sampler_struct, device):
    """One iteration of the FlowModel.sample loop (pf_sampler_step): denoiser + post-processing + Euler update."""
    import ctypes
    lib = _lib.lib_for(device)
    check(lib.pf_sampler_step(ctypes.addressof(sampler_struct), stream(device)))


def edge_embed(aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, consts):
    """EdgeEmbedder.forward (models_con/edge.py:39-112) as one fused kernel (pf_edge_embed).  `consts` is the tuple of
    host-prepared constants in the order of the C prototype (EdgeEmbedder._kernel_constants)."""
    lib = _lib.lib_for(pos_atoms.device)
    N, L = aa.shape
    pos = _c(pos_atoms)
    m = _c(mask_atoms, torch.bool).view(U8)
    sm = _c(structure_mask, torch.bool).view(U8) if structure_mask is not None else None
    out = torch.empty(N, L, L, 64, device=pos.device, dtype=F32)
    cs = [_c(c) for c in consts]
    check(lib.pf_edge_embed(ptr(_c(aa, I64), I64), ptr(_c(res_nb, I64), I64), ptr(_c(chain_nb, I64), I64), ptr(pos),
                            ptr(m, U8), ptr(sm, U8, allow_none=True), *[ptr(c) for c in cs], ptr(out), N, L,
                            pos.shape[2], stream()))
    return out


def node_embed(aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, consts):
    """NodeEmbedder.forward (models_con/node.py:35-105) as one fused kernel (pf_node_embed); `consts` in the order of
    the C prototype (NodeEmbedder._kernel_constants)."""
    lib = _lib.lib_for(pos_atoms.device)
    N, L = aa.shape
    pos = _c(pos_atoms)
    m = _c(mask_atoms, torch.bool).view(U8)
    sm = _c(structure_mask, torch.bool).view(U8) if structure_mask is not None else None
    out = torch.empty(N, L, 128, device=pos.device, dtype=F32)
    cs = [_c(c) for c in consts]
    check(lib.pf_node_embed(ptr(_c(aa, I64), I64), ptr(_c(res_nb, I64), I64), ptr(_c(chain_nb, I64), I64), ptr(pos),
                            ptr(m, U8), ptr(sm, U8, allow_none=True), *[ptr(c) for c in cs], ptr(out), N, L,
                            pos.shape[2], stream()))
    return out


def full_atom_reconstruction(rot, trans, angles, aa, tables, want_frames=True, want_mask=False):
    """full_atom_reconstruction (models_con/torsion.py:140-226) as one kernel (pf_full_atom_reconstruction).
    `tables` = constants.rigid_tables(device).  Returns (pos14 [B,L,14,3], R [B,L,6,3,3] | None, t [B,L,6,3] | None,
    mask [B,L,15] bool | None)."""
    lib = _lib.lib_for(rot.device)
    B, L = aa.shape
    n = B * L
    dev = rot.device
    pos14 = torch.empty(B, L, 14, 3, device=dev, dtype=F32)
    R_ret = torch.empty(B, L, 6, 3, 3, device=dev, dtype=F32) if want_frames else None
    t_ret = torch.empty(B, L, 6, 3, device=dev, dtype=F32) if want_frames else None
    mask = torch.empty(B, L, 15, device=dev, dtype=U8) if want_mask else None
    check(lib.pf_full_atom_reconstruction(
        ptr(_c(rot)), ptr(_c(trans)), ptr(_c(angles)), ptr(_c(aa, I64), I64), ptr(tables["rigid_rot"]),
        ptr(tables["rigid_trans"]), ptr(tables["atom_group"], torch.int32), ptr(tables["atom_pos"]),
        ptr(tables["heavyatom_mask"], U8), ptr(pos14), ptr(R_ret, allow_none=True), ptr(t_ret, allow_none=True),
        ptr(mask, U8, allow_none=True), n, stream()))
    return pos14, R_ret, t_ret, (mask.view(torch.bool) if want_mask else None)


def reconstruct_backbone(rot, trans, aa, chain_nb, res_nb, mask, tables):
    """reconstruct_backbone (pepflow/modules/common/geometry.py:446-489) as one kernel: [B,L,4,3] = N, CA, C, O."""
    lib = _lib.lib_for(rot.device)
    B, L = aa.shape
    out = torch.empty(B, L, 4, 3, device=rot.device, dtype=F32)
    m = _c(mask, torch.bool).view(U8)
    check(lib.pf_reconstruct_backbone(ptr(_c(rot)), ptr(_c(trans)), ptr(_c(aa, I64), I64), ptr(_c(chain_nb, I64), I64),
                                      ptr(_c(res_nb, I64), I64), ptr(m, U8), ptr(tables["bb_coords"]),
                                      ptr(tables["bb_oxygen"]), ptr(out), B, L, stream()))
    return out


def torsion_angles(pos_atoms, aa, tables):
    """get_torsion_angle (models_con/torsion.py:49-66) for every residue of pos_atoms [..., A, 3] / aa [...] in one
    launch (pf_torsion_angles): (torsion [..., 5] in [0, 2 pi), mask [..., 5] bool)."""
    lib = _lib.lib_for(pos_atoms.device)
    pos = _c(pos_atoms)
    n = aa.numel()
    tor = torch.empty(*aa.shape, 5, device=pos.device, dtype=F32)
    mask = torch.empty(*aa.shape, 5, device=pos.device, dtype=U8)
    check(lib.pf_torsion_angles(ptr(pos), ptr(_c(aa, I64), I64), ptr(tables["chi_atoms"], torch.int32), ptr(tor),
                                ptr(mask, U8), n, pos.shape[-2], stream()))
    return tor, mask.view(torch.bool)


# ---- device scoping -----------------------------------------------------------------------------------------------------
# Every wrapper above launches on "torch's current stream", which only means the tensors' stream while their device is the
# current one.  Wrap each public op so that a call with tensors on another device runs under torch.cuda.device(that device)
# (and therefore on that device's current stream); on the usual single-device process this is one integer comparison.
def _device_scoped(fn):
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


for _name, _fn in list(globals().items()):
    if callable(_fn) and getattr(_fn, "__module__", None) == __name__ and not _name.startswith("_"):
        globals()[_name] = _device_scoped(_fn)
del _name, _fn
