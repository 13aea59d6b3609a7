"""Multi-GPU plumbing of the sampling path: complexes are independent, so the global batch
is sharded over ranks with NO data-path collective; only the timing is reduced (MAX over ranks) and, when a
caller wants the whole trajectory on one host, the per-rank results are gathered.  One process per GPU,
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world_info():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(global_batch, rank, world):
    """Contiguous block of complex indices owned by `rank`: sizes differ by at most one, blocks are disjoint and
    cover range(global_batch) (the replicas of inference.py:73 / the complexes of a names.txt chunk)."""
    base, rem = divmod(int(global_batch), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(x, device=None):
    """MAX-reduce a host scalar over ranks (device-side timing of a multi-GPU run is the slowest rank's)."""
    rank, world = world_info()
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_trajectory_step(step, dst=0):
    """Concatenate one trajectory entry (dict of CPU tensors with the batch on dim 0) from all ranks on `dst`
    in shard order; other ranks get None.  Not on the timed path."""
    rank, world = world_info()
    if world == 1:
        return step
    parts = [None] * world if rank == dst else None
    dist.gather_object(step, parts, dst=dst)
    if rank != dst:
        return None
    return {k: torch.cat([p[k] for p in parts], dim=0) for k in step}
