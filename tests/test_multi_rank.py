"""world_size-2 gloo tests (CPU) of the sharding / reduction plumbing used by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pepflowww_b200.dist_utils import gather_trajectory_step, max_over_ranks, shard_range, world_info
        from pepflowww_b200.pep_dataloader import synthetic_batch
        assert world_info() == (rank, world)
        lo, hi = shard_range(5, rank, world)
        batch = synthetic_batch(hi - lo, 12, 4, seed=0, first_index=lo)
        slowest = max_over_ranks(10.0 + rank)
        step = {"seqs": batch["aa"].clone(), "trans": batch["pos_heavyatom"][:, :, 1].clone()}
        full = gather_trajectory_step(step, dst=0)
        if rank == 0:
            ref = synthetic_batch(5, 12, 4, seed=0, first_index=0)
            ok = torch.equal(full["seqs"], ref["aa"]) and torch.equal(full["trans"], ref["pos_heavyatom"][:, :, 1])
            out.put((slowest, (lo, hi), bool(ok)))
        else:
            assert full is None
            out.put((slowest, (lo, hi), True))
    finally:
        dist.destroy_process_group()


def test_shards_are_disjoint_cover_and_reduce_max():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] == 11.0 for r in res)                 # MAX over ranks
    assert sorted(r[1] for r in res) == [(0, 3), (3, 5)]   # contiguous, disjoint, covering
    assert all(r[2] for r in res)                          # gathered shards == the unsharded batch


@pytest.mark.parametrize("n,world", [(512, 8), (5, 2), (3, 4), (0, 2)])
def test_shard_range_properties(n, world):
    from pepflowww_b200.dist_utils import shard_range
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


class _DenoiserLoss(torch.nn.Module):
    """Loss dict over the denoiser's autograd formulation - what FlowModel.forward feeds train_step - on CPU tensors."""

    def __init__(self):
        super().__init__()
        from pepflowww_b200.config import load_config
        from pepflowww_b200.ga import GAEncoder
        cfg, _ = load_config()
        self.cfg = cfg
        self.ga_encoder = GAEncoder(cfg.model.encoder.ipa)

    def forward(self, b):
        R, x, ang, logits = self.ga_encoder.forward_autograd(b["t"], b["R"], b["x"], b["ang"], b["seq"], b["node"], b["edge"],
                                                             b["mask"], b["mask"])
        return {"trans_loss": (x ** 2).mean(), "rot_loss": ((R - b["R"]) ** 2).mean(), "bb_atom_loss": x.abs().mean(),
                "seqs_loss": torch.nn.functional.cross_entropy(logits.reshape(-1, 20), b["seq"].reshape(-1).clamp(0, 19)),
                "angle_loss": torch.sin(ang).pow(2).mean(), "torsion_loss": torch.cos(ang).mean()}


def _ddp_inputs(seed, B=1, L=9):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, L, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    a, b, c, d = q.unbind(-1)
    R = torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c), 2 * (b * c + a * d),
                     a * a - b * b + c * c - d * d, 2 * (c * d - a * b), 2 * (b * d - a * c), 2 * (c * d + a * b),
                     a * a - b * b - c * c + d * d], -1).reshape(B, L, 3, 3)
    return {"t": torch.rand(B, 1, generator=g), "R": R, "x": torch.randn(B, L, 3, generator=g) * 3,
            "ang": torch.rand(B, L, 5, generator=g) * 6.28, "seq": torch.randint(0, 20, (B, L), generator=g),
            "node": torch.randn(B, L, 128, generator=g), "edge": torch.randn(B, L, L, 64, generator=g),
            "mask": torch.ones(B, L, dtype=torch.long)}


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torch.nn.parallel import DistributedDataParallel as DDP
        from pepflowww_b200 import train
        torch.manual_seed(0)                                   # same initial weights on every rank
        net = _DenoiserLoss()
        with torch.no_grad():                                  # 'final'-initialised layers are zero: give them signal
            for p in net.parameters():
                if float(p.abs().sum()) == 0:
                    p.normal_(0, 0.02, generator=torch.Generator().manual_seed(p.numel()))
        weights = net.cfg.train.loss_weights
        # local (unreduced) gradients of every rank's batch, computed without DDP
        local = []
        for r in range(world):
            net.zero_grad()
            train.sum_weighted_losses(net(_ddp_inputs(114514 + 100 * r)), weights).backward()
            local.append(torch.cat([p.grad.flatten() for p in net.parameters()]).clone())
        net.zero_grad()
        ddp = DDP(net)
        train.sum_weighted_losses(ddp(_ddp_inputs(114514 + 100 * rank)), weights).backward()
        got = torch.cat([p.grad.flatten() for p in net.parameters()])
        want = sum(local) / world
        err = float((got - want).abs().max() / want.abs().max())
        # one optimiser step through the harness keeps the replicas identical
        opt = train.get_optimizer(net.cfg.train.optimizer, ddp)
        net.zero_grad()
        loss, parts, gnorm = train.train_step(ddp, _ddp_inputs(7 + rank), opt, weights, net.cfg.train.max_grad_norm)
        digest = float(sum(p.double().sum() for p in net.parameters()))
        # the loader of the harness: disjoint shards of one epoch, collated to the PaddingCollate schema
        from pepflowww_b200.pep_dataloader import SyntheticPepDataset
        loader = train.make_loader(SyntheticPepDataset(8, 6, 3, seed=0), 2, rank, world, seed=114514)
        ids = [i for b in loader for i in b["id"]]
        first = next(train.inf_iterator(loader))
        assert first["aa"].shape == (2, 16) and first["res_mask"].sum() == 18          # 9 residues padded to 16
        out.put((rank, err, digest, float(loss), sorted(parts), int(got.numel()), ids))
    finally:
        dist.destroy_process_group()


def test_ddp_gradient_allreduce_on_the_denoiser():
    """cfg5 plumbing (train_ddp.py:94,117-150): DistributedDataParallel over the denoiser's autograd formulation averages
    the per-rank gradients (== mean of the unreduced local gradients), and train_step keeps the replicas in lockstep."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] < 1e-5 for r in res), res
    assert res[0][2] == res[1][2]                              # identical parameters after the step
    assert res[0][4] == ["angle_loss", "bb_atom_loss", "rot_loss", "seqs_loss", "torsion_loss", "trans_loss"]
    assert res[0][5] > 6_000_000                               # the ga_encoder's parameters all took part
    assert len(res[0][6]) == 4 and not set(res[0][6]) & set(res[1][6]) and len(set(res[0][6]) | set(res[1][6])) == 8
