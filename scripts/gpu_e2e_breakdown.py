"""Diagnostic: where the end-to-end FlowModel.sample time goes (encode / Euler loop / trajectory D2H)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pepflowww_b200.config import load_config  # noqa: E402
from pepflowww_b200.flow_model import FlowModel  # noqa: E402
from pepflowww_b200.pep_dataloader import synthetic_batch  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict, recursive_to  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cfg, _ = load_config()
    model = FlowModel(cfg.model).eval()
    model.load_state_dict(deterministic_state_dict(model.state_dict(), 114514))
    model = model.to(dev)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    host = synthetic_batch(B, 256, 15, seed=0)
    host = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch = recursive_to(host, dev)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        with torch.no_grad():
            enc = model.encode(batch)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        traj = model.sample(batch, num_steps=200, seed=1)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"rep {rep}: H2D {1e3 * (t1 - t0):.1f} ms, encode {1e3 * (t2 - t1):.1f} ms, sample(200) total {1e3 * (t3 - t2):.1f} ms "
              f"(includes its own encode)", flush=True)
    # inside sample: time the loop alone
    smp = model.sampler_init(batch, num_steps=200, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for n in range(199):
        smp.step(n)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"Euler loop alone: {1e3 * (t1 - t0):.1f} ms for 199 steps", flush=True)


if __name__ == "__main__":
    main()
