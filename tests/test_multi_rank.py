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
