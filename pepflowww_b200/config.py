"""Config surface: configs/learn_angle.yaml loads unchanged.

Mirrors pepflow/utils/misc.py:110-114 (load_config -> (EasyDict, name)); easydict is
not in the image, so AttrDict restates the attribute-dict behaviour the reference relies on.
"""
import os

import yaml


class AttrDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(AttrDict(x) if isinstance(x, dict) else x for x in v)
        dict.__setitem__(self, k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


EasyDict = AttrDict

DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "learn_angle.yaml")


def load_config(config_path=DEFAULT_CONFIG):
    with open(config_path, "r") as f:
        config = AttrDict(yaml.safe_load(f))
    base = os.path.basename(config_path)
    return config, base[: base.rfind(".")]
