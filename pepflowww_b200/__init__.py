"""pepflowww_b200: B200-native (sm_100a) implementation of PepFlow's flow-matching denoising hot path.

Public surface mirrors the reference: FlowModel (forward / sample / encode), GAEncoder, the IPA blocks,
so3_utils / torus maps, the pep_dataloader batch schema and configs/learn_angle.yaml.
The arithmetic runs in hand-written CUDA kernels behind the C ABI of include/pepflow_b200.h
(libpepflow_b200.so, built in-tree by `python -m pepflowww_b200.build`).
"""
from .config import load_config  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "FlowModel":
        from .flow_model import FlowModel
        return FlowModel
    if name == "GAEncoder":
        from .ga import GAEncoder
        return GAEncoder
    raise AttributeError(name)
