"""Small host helpers kept from the reference surface.

process_dic: models_con/utils.py:16-23 (strip the DDP 'module.' prefix so model1.pt/model2.pt
load unchanged).  recursive_to: pepflow/utils/train.py:126-140.  seed_all: pepflow/utils/misc.py:68-73.
deterministic_state_dict: test/bench helper - reproducible non-degenerate weights (the
reference's init='final' layers are all-zero at fresh init, SURVEY.md finding 10).
"""
import random
import zlib

import numpy as np
import torch


def process_dic(state_dict):
    out = {}
    for k, v in state_dict.items():
        out[k[7:] if "module" in k else k] = v
    return out


def load_checkpoint(path, map_location="cpu"):
    """torch.load of a training checkpoint {config, model, optimizer, scheduler, iteration} (train_ddp.py:206-212;
    loaded at inference.py:61-65).  The reference pickles its config as easydict.EasyDict; easydict is not in this
    image, so a stand-in module whose EasyDict is config.AttrDict is registered for the duration of the load.
    weights_only=False: the config object is not a tensor (torch >= 2.6 defaults to weights_only=True and would reject
    it) - only load checkpoints you trust, as with the reference."""
    import sys
    import types

    from .config import AttrDict
    shim = None
    if "easydict" not in sys.modules:
        shim = types.ModuleType("easydict")
        shim.EasyDict = AttrDict
        sys.modules["easydict"] = shim
    try:
        return torch.load(path, map_location=map_location, weights_only=False)
    finally:
        if shim is not None:
            sys.modules.pop("easydict", None)


def plain_config(obj):
    """Nested AttrDict / EasyDict -> plain dict / list (what checkpoint_dict stores: loadable without this package)."""
    if isinstance(obj, dict):
        return {k: plain_config(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(plain_config(v) for v in obj)
    return obj


def recursive_to(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, list):
        return [recursive_to(o, device) for o in obj]
    if isinstance(obj, tuple):
        return tuple(recursive_to(o, device) for o in obj)
    if isinstance(obj, dict):
        return {k: recursive_to(v, device) for k, v in obj.items()}
    return obj


def seed_all(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def deterministic_state_dict(reference_sd, seed=114514):
    """Returns a new state dict with the same keys/shapes/dtypes as `reference_sd`, filled from
    numpy PCG64 streams keyed by (seed, crc32(key)) - independent of torch's RNG and of key order.
    Matrices: U(-a, a), a = sqrt(3/fan_in) (unit-variance preserving); biases U(-0.1, 0.1);
    LayerNorm / norm weights 1 + U(-0.1, 0.1); embeddings U(-1, 1); IPA head_weights 0.54 + U(-0.3, 0.3);
    buffers (freq_bands) are kept."""
    out = {}
    for key, ref in reference_sd.items():
        if key.endswith("freq_bands"):
            out[key] = ref.clone()
            continue
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))
        u = torch.from_numpy(rng.random(tuple(ref.shape), dtype=np.float32) * 2.0 - 1.0)
        leaf = key.split(".")[-1]
        if key.endswith("head_weights"):
            val = 0.541324854612918 + 0.3 * u
        elif "embed.weight" in key or key.endswith("current_seq_embedder.weight"):
            val = u
        elif key.endswith("aapair_to_distcoef.weight"):
            val = 0.5 * u
        elif ref.dim() >= 2:
            val = u * float(np.sqrt(3.0 / ref.shape[-1]))
        elif leaf == "bias" or leaf == "in_proj_bias":
            val = 0.1 * u
        elif leaf == "weight":  # 1-D weight = a norm scale
            val = 1.0 + 0.1 * u
        else:
            val = 0.1 * u
        out[key] = val.to(ref.dtype)
    return out
