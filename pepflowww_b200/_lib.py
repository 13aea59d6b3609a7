"""ctypes binding of libpepflow_b200.so (the C ABI declared in include/pepflow_b200.h).

The product path has NO fallback: if the library is missing, or a tensor is not a CUDA tensor,
these wrappers raise.  Build with `python -m pepflowww_b200.build` (or __graft_entry__.build()).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpepflow_b200.so")

_p = C.c_void_p
_i = C.c_int
_sz = C.c_size_t
_u64 = C.c_uint64
_f = C.c_float

# name -> (restype, argtypes); every symbol include/pepflow_b200.h declares
SIGNATURES = {
    "pf_version": (_i, []),
    "pf_strerror": (C.c_char_p, [_i]),
    "pf_init": (_i, [_i]),
    "pf_check_config": (_i, [_i] * 8),
    "pf_set_option": (_i, [C.c_char_p, _i]),
    "pf_get_option": (_i, [C.c_char_p]),
    "pf_launch_count": (C.c_int64, []),
    "pf_reset_launch_count": (None, []),
    "pf_debug_buffer": (_i, [_p, _sz]),
    "pf_ga_prepack_bytes": (_sz, [_p]),
    "pf_ga_prepack": (_i, [_p, _p, _sz, _p]),
    "pf_profile_enable": (_i, [_i]),
    "pf_profile_read": (_i, [_p, _p, _p, _p]),
    "pf_profile_read_category": (_i, [_i, _p, _p]),
    "pf_linear": (_i, [_p] * 6 + [_i] * 4 + [_p]),
    "pf_linear_workspace_bytes": (_sz, [_i]),
    "pf_linear_ws": (_i, [_p] * 6 + [_i] * 4 + [_p, _sz, _p]),
    "pf_add_layernorm": (_i, [_p] * 6 + [_i] * 2 + [_p]),
    "pf_mix_features": (_i, [_p] * 8 + [_i] * 2 + [_p]),
    "pf_ipa_points": (_i, [_p] * 4 + [_i] * 2 + [_p]),
    "pf_ipa_attention_workspace_bytes": (_sz, [_i, _i]),
    "pf_ipa_attention": (_i, [_p] * 13 + [_sz] + [_i] * 2 + [_p]),
    "pf_seq_attention": (_i, [_p] * 3 + [_i] * 2 + [_p]),
    "pf_rigid_update": (_i, [_p] * 8 + [_i] + [_p]),
    "pf_edge_transition_workspace_bytes": (_sz, [_i, _i]),
    "pf_edge_transition": (_i, [_p] * 15 + [_sz] + [_i] * 2 + [_p]),
    "pf_mod_2pi": (_i, [_p, _p, _i, _p]),
    "pf_quat_to_rot": (_i, [_p, _p, _i, _p]),
    "pf_so3_log": (_i, [_p, _p, _i, _p]),
    "pf_so3_exp": (_i, [_p, _p, _i, _p]),
    "pf_so3_geodesic": (_i, [_p] * 4 + [_i, _i, _p]),
    "pf_tor_geodesic": (_i, [_p] * 4 + [_i, _i, _i, _p]),
    "pf_denoise_post": (_i, [_p] * 11 + [_u64, _u64] + [_p] * 5 + [_i, _f, _p]),
    "pf_euler_step": (_i, [_p] * 17 + [_u64, _u64, _f] + [_p] * 5 + [_i, _f, _p]),
    "pf_ga_encoder_workspace_bytes": (_sz, [_i, _i]),
    "pf_ga_encoder_forward": (_i, [_p] * 15 + [_sz] + [_i] * 2 + [_p]),
    "pf_edge_embed": (_i, [_p] * 21 + [_i] * 3 + [_p]),
    "pf_node_embed": (_i, [_p] * 16 + [_i] * 3 + [_p]),
    "pf_full_atom_reconstruction": (_i, [_p] * 13 + [C.c_longlong, _p]),
    "pf_reconstruct_backbone": (_i, [_p] * 9 + [_i] * 2 + [_p]),
    "pf_torsion_angles": (_i, [_p] * 5 + [C.c_longlong, _i, _p]),
    "pf_seq_transformer_forward": (_i, [_p, _i, _p, _p, _p, _p, _sz, _i, _i, _p]),
    "pf_sampler_step": (_i, [_p, _p]),
    "pf_zero_center": (_i, [_p] * 4 + [_i, _i, _p]),
}

# enum sizes of include/pepflow_b200.h
PF_G_NSLOTS = 19
PF_B_NSLOTS = 57
PF_MAX_BLOCKS = 8

G_SLOTS = ["MIX0_W", "MIX0_B", "MIX2_W", "MIX2_B", "SEQ_EMB", "ANG_FREQS", "TIME_FREQS",
           "SEQNET0_W", "SEQNET0_B", "SEQNET2_W", "SEQNET2_B", "SEQNET4_W", "SEQNET4_B",
           "ANGNET0_W", "ANGNET0_B", "ANGNET2_W", "ANGNET2_B", "ANGNET4_W", "ANGNET4_B"]
B_SLOTS = ["PROJ_W", "PROJ_B", "LINB_W", "LINB_B", "DOWNZ_W", "DOWNZ_B", "HEAD_W", "OUT_W", "OUT_B", "IPA_LN_G", "IPA_LN_B"]
for _l in (0, 1):
    B_SLOTS += [f"T{_l}_{n}" for n in ("IN_W", "IN_B", "OUT_W", "OUT_B", "L1_W", "L1_B", "L2_W", "L2_B",
                                       "N1_G", "N1_B", "N2_G", "N2_B")]
B_SLOTS += ["POST_W", "POST_B", "NT1_W", "NT1_B", "NT2_W", "NT2_B", "NT3_W", "NT3_B", "NT_LN_G", "NT_LN_B", "BB_W", "BB_B",
            "ET_INIT_W", "ET_INIT_B", "ET_W1", "ET_B1", "ET_W2", "ET_B2", "ET_WF", "ET_BF", "ET_LN_G", "ET_LN_B"]
assert len(G_SLOTS) == PF_G_NSLOTS and len(B_SLOTS) == PF_B_NSLOTS


class GaWeights(C.Structure):
    _fields_ = [("num_blocks", C.c_int32), ("reserved", C.c_int32),
                ("g", _p * PF_G_NSLOTS), ("blk", (_p * PF_B_NSLOTS) * PF_MAX_BLOCKS),
                ("prepacked", _p), ("prepacked_bytes", C.c_uint64)]


class Sampler(C.Structure):
    """pf_sampler of include/pepflow_b200.h (field order is the ABI)."""
    _fields_ = ([("weights", _p), ("node_embed", _p), ("edge_embed", _p), ("res_mask", _p), ("workspace", _p),
                 ("workspace_bytes", C.c_uint64)] +
                [(n, _p) for n in ("rot1", "trans1", "ang1", "seq1", "gen_mask", "torsions_mask", "trans0", "simplex0",
                                   "rot_t", "trans_t", "ang_t", "seq_t", "simplex_t", "pred_rot", "pred_trans", "pred_ang",
                                   "logits", "traj_rot", "traj_trans", "traj_ang", "traj_seq", "traj_simplex", "ts",
                                   "uniforms", "step", "t_cur")] +
                [("seed", C.c_uint64), ("num_steps", C.c_int32), ("B", C.c_int32), ("L", C.c_int32),
                 ("sample_bb", C.c_int32), ("sample_ang", C.c_int32), ("sample_seq", C.c_int32),
                 ("simplex_k", C.c_float), ("reserved", C.c_int32)])


_lib = None
_inited = set()


def load():
    """Loads the shared library (no GPU needed) and sets the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built "
            "(run `python -m pepflowww_b200.build`). There is no CPU or PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise RuntimeError(f"pepflow_b200: {load().pf_strerror(status).decode()} (status {status})")


def lib_for(device):
    """Library handle initialised for `device` (a CUDA torch.device)."""
    if device.type != "cuda":
        raise RuntimeError("pepflow_b200 kernels need CUDA tensors (no CPU fallback); got device %s" % device)
    lib = load()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited:
        check(lib.pf_init(idx))
        _inited.add(idx)
    return lib


def ptr(t, dtype=torch.float32, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError("tensor required")
    if not t.is_cuda:
        raise RuntimeError("pepflow_b200 kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream(device=None):
    """Raw handle of torch's current stream on `device` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def set_option(name, value):
    check(load().pf_set_option(name.encode(), int(value)))


def get_option(name):
    return load().pf_get_option(name.encode())


def launch_count():
    return int(load().pf_launch_count())


def reset_launch_count():
    load().pf_reset_launch_count()


def profile_enable(on):
    check(load().pf_profile_enable(int(bool(on))))


def profile_read():
    """{'ipa': (ms, launches), 'edge': ..., 'ipa_pack': ...} since the last read (synchronises)."""
    out = {}
    for cat, name in enumerate(("ipa", "edge", "ipa_pack")):
        ms, n = C.c_double(0), C.c_int64(0)
        check(load().pf_profile_read_category(cat, C.byref(ms), C.byref(n)))
        out[name] = (ms.value, n.value)
    return out
