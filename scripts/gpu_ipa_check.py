"""Diagnostic (not a test): IPA attention variants against each other + timing.  python scripts/gpu_ipa_check.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pepflowww_b200 import _lib  # noqa: E402
from pepflowww_b200.config import load_config  # noqa: E402
from pepflowww_b200.flow_model import FlowModel  # noqa: E402
from pepflowww_b200.rigid import create_rigid  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def rand_rot(n, g):
    q = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    a, b, c, d = q.unbind(-1)
    return torch.stack([a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c,
                        2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b,
                        2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d], -1).view(n, 3, 3)


def main():
    dev = torch.device("cuda:0")
    cfg, _ = load_config()
    model = FlowModel(cfg.model).eval()
    model.load_state_dict(deterministic_state_dict(model.state_dict(), 114514))
    model = model.to(dev)
    ipa = model.ga_encoder.trunk["ipa_1"]
    for B, L in [(2, 30), (1, 7), (3, 37), (2, 140), (2, 271), (64, 271)]:
        g = torch.Generator().manual_seed(L)
        s = torch.randn(B, L, 128, generator=g).to(dev)
        z = torch.randn(B, L, L, 64, generator=g).to(dev) if B < 64 else torch.randn(B, L, L, 64, device=dev)
        R = rand_rot(B * L, g).view(B, L, 3, 3).to(dev)
        x = (torch.randn(B, L, 3, generator=g) * 8.0).to(dev)
        m = (torch.rand(B, L, generator=g) > 0.15).float().to(dev)
        rig = create_rigid(R, x)
        outs = {}
        for impl in ((0, 3, 4) if B < 64 else (3, 4)):
            _lib.set_option("ipa_impl", impl)
            with torch.no_grad():
                outs[impl] = ipa(s, z, rig, m)
            torch.cuda.synchronize()
        ref = outs[0] if 0 in outs else outs[3]
        print(f"B={B} L={L}: " + "  ".join(f"v{k} vs ref {rel(v * m[..., None], ref * m[..., None]):.2e}" for k, v in outs.items()), flush=True)
        if B == 64:
            for impl in (3, 4):
                _lib.set_option("ipa_impl", impl)
                _lib.profile_enable(True)
                with torch.no_grad():
                    for _ in range(3):
                        ipa(s, z, rig, m)
                torch.cuda.synchronize()
                ipa_ms, ipa_n = _lib.profile_read()["ipa"]
                _lib.profile_enable(False)
                print(f"impl {impl}: attention kernel {ipa_ms / ipa_n:.3f} ms (B=64, L=271)", flush=True)


if __name__ == "__main__":
    main()
