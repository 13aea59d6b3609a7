"""Data-parallel training harness with the reference's train_ddp.py surface (train_ddp.py:40-215; helpers
pepflow/utils/train.py:11-53,143-155): per-rank seed `seed + 100 * rank`, DistributedDataParallel over the whole
FlowModel (one bucketed all-reduce of the 6.88 M fp32 gradients per iteration - NCCL over NVLink on the GPU box, gloo in
the CPU tests), weighted sum of the six losses, NaN-gradient rescue, gradient clipping, Adam + plateau scheduler, and the
checkpoint dict {config, model, optimizer, scheduler, iteration} that inference.py:61-65 loads.

SURVEY.md section 8f rank 4 / cfg5.  The forward under autograd goes through ga_autograd.denoiser_autograd (torch
ops over the same parameters the kernels read); hand-written backward kernels are the open part of that row.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        -m pepflowww_b200.train --iters 10 --batch-size 32 --pocket 48 --peptide 12
trains on synthetic complexes (the PepMerge LMDB is not available offline) and prints one JSON line per run."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
from torch.nn.utils import clip_grad_norm_


def sum_weighted_losses(losses, weights):
    total = 0
    for k, v in losses.items():
        total = total + (v if weights is None else weights[k] * v)
    return total


def get_optimizer(cfg, model):
    if cfg.type == "adam":
        return torch.optim.Adam(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay, betas=(cfg.beta1, cfg.beta2),
                                capturable=bool(cfg.get("capturable", False)))
    if cfg.type == "adamw":
        return torch.optim.AdamW(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay)
    raise NotImplementedError("Optimizer not supported: %s" % cfg.type)


def get_scheduler(cfg, optimizer):
    if cfg.type == "plateau":
        return torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, factor=cfg.factor, patience=cfg.patience, min_lr=cfg.min_lr)
    if cfg.type == "multistep":
        return torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=cfg.milestones, gamma=cfg.gamma)
    if cfg.type == "exp":
        return torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=cfg.gamma)
    raise NotImplementedError("Scheduler not supported: %s" % cfg.type)


def train_step(model, batch, optimizer, loss_weights, max_grad_norm):
    """One iteration of train_ddp.py:117-150.  Returns (loss, loss_dict, gradient norm before clipping)."""
    model.train()
    loss_dict = model(batch)
    loss = sum_weighted_losses(loss_dict, loss_weights)
    loss.backward()
    for p in model.parameters():                       # rescue for NaN gradients, train_ddp.py:139-142
        if p.grad is not None:
            torch.nan_to_num_(p.grad, nan=0.0, posinf=float("inf"), neginf=float("-inf"))
    grad_norm = clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach(), {k: v.detach() for k, v in loss_dict.items()}, grad_norm


class GraphedTrainStep:
    """One training iteration (forward, weighted loss, backward, NaN rescue, clipping, optimizer step) captured ONCE in a
    CUDA graph and replayed on a static input batch - the iteration is otherwise bound by the ~6,000 kernel launches of
    the autograd formulation, not by arithmetic.  Needs: fixed batch shapes, an optimizer built with
    capturable=True, and (with DistributedDataParallel) >= 11 eager iterations before capture so that DDP's bucket
    views / reducer state are final; the NCCL all-reduces of the gradient buckets are captured as graph nodes."""

    def __init__(self, model, optimizer, loss_weights, max_grad_norm, example_batch):
        self.model, self.optimizer = model, optimizer
        self.static = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in example_batch.items()}
        dev = next(model.parameters()).device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        model.train()
        with torch.cuda.stream(side):
            optimizer.zero_grad(set_to_none=True)      # backward inside the capture allocates the grads in the graph's pool
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
                loss_dict = model(self.static)
                loss = sum_weighted_losses(loss_dict, loss_weights)
                loss.backward()
                for p in model.parameters():
                    if p.grad is not None:
                        torch.nan_to_num_(p.grad, nan=0.0, posinf=float("inf"), neginf=float("-inf"))
                self.grad_norm = clip_grad_norm_(model.parameters(), max_grad_norm)
                optimizer.step()
            self.loss, self.loss_dict = loss.detach(), {k: v.detach() for k, v in loss_dict.items()}
        torch.cuda.current_stream(dev).wait_stream(side)

    def __call__(self, batch):
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss, self.loss_dict, self.grad_norm


def nccl_overlap_summary(prof, iters):
    """From a torch.profiler trace of `iters` iterations: device time of the NCCL kernels (DDP's gradient all-reduce)
    per iteration and the share of it that runs while a compute kernel is executing on another stream."""
    ker = [e for e in prof.events() if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()
           and e.time_range is not None]
    ivs = [(e.name, e.time_range.start, e.time_range.end) for e in ker]
    nccl = [(a, b) for n, a, b in ivs if "nccl" in n.lower()]
    comp = sorted((a, b) for n, a, b in ivs if "nccl" not in n.lower() and "memcpy" not in n.lower())
    merged = []
    for a, b in comp:
        if merged and a <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], b)
        else:
            merged.append([a, b])
    over = 0.0
    for a, b in nccl:
        for c, d in merged:
            if d <= a:
                continue
            if c >= b:
                break
            over += min(b, d) - max(a, c)
    total = sum(b - a for a, b in nccl)
    names = sorted({n for n, _, _ in ivs if "nccl" in n.lower()})
    return {"nccl_kernel_ms_per_iter": total / 1e3 / max(1, iters), "nccl_launches_per_iter": len(nccl) / max(1, iters),
            "overlapped_with_compute": (over / total) if total > 0 else None, "kernels": names[:4],
            "compute_kernel_launches_per_iter": len(comp) / max(1, iters)}


def inf_iterator(iterable):
    """Endless pass over a DataLoader (pepflow/utils/misc.py inf_iterator, used at train_ddp.py:91)."""
    while True:
        for x in iterable:
            yield x


def make_loader(dataset, batch_size, rank=0, world=1, num_workers=0, seed=0):
    """DistributedSampler(shuffle=True) + DataLoader(PaddingCollate()) as train_ddp.py:88-91: every rank draws a
    disjoint shard of each epoch's permutation."""
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler

    from .pep_dataloader import PaddingCollate
    sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed)
    return DataLoader(dataset, batch_size=batch_size, collate_fn=PaddingCollate(), sampler=sampler,
                      num_workers=num_workers, pin_memory=torch.cuda.is_available())


def checkpoint_dict(config, model, optimizer, scheduler, iteration):
    """train_ddp.py:206-212: `model` is what the training loop holds (the DDP wrapper when distributed, so its keys carry
    the `module.` prefix exactly like the reference's checkpoints; utils.process_dic strips it on load); `iteration` is
    the last completed iteration; the config goes in as plain dicts so the file loads without this package."""
    from .utils import plain_config
    return {"config": plain_config(config), "model": model.state_dict(), "optimizer": optimizer.state_dict(),
            "scheduler": scheduler.state_dict(), "iteration": iteration}


def main(argv=None):
    from .config import load_config
    from .flow_model import FlowModel
    from .pep_dataloader import SyntheticPepDataset
    from .utils import recursive_to, seed_all

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--pocket", type=int, default=48)
    ap.add_argument("--peptide", type=int, default=12)
    ap.add_argument("--save", default=None)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--dataset-size", type=int, default=4096, help="synthetic complexes per epoch")
    ap.add_argument("--graph", action="store_true", help="capture the iteration in a CUDA graph (fixed batch shapes)")
    ap.add_argument("--tf32", action="store_true", help="TF32 tensor-core matmuls in the autograd path (not the reference's numerics)")
    ap.add_argument("--profile", type=int, default=0, help="profile this many extra iterations (NCCL time / overlap)")
    ap.add_argument("--out", default=None, help="append the JSON line to this file")
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("pepflowww_b200.train needs a CUDA device (the loss path launches the CUDA kernels)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    config, _ = load_config(args.config) if args.config else load_config()
    seed_all(config.train.seed + 100 * rank)
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    net = FlowModel(config.model).to(dev)
    if world > 1 and args.graph:                        # DDP for graph capture is built on a side stream
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            model = DDP(net, device_ids=[local_rank], broadcast_buffers=False)
        torch.cuda.current_stream(dev).wait_stream(side)
    else:
        model = DDP(net, device_ids=[local_rank], broadcast_buffers=False) if world > 1 else net
    if args.graph:
        config.train.optimizer["capturable"] = True
    optimizer = get_optimizer(config.train.optimizer, model)
    scheduler = get_scheduler(config.train.scheduler, optimizer)
    bs = args.batch_size or config.train.batch_size
    it_first = 1
    if args.resume:                                     # train_ddp.py:105-114
        from .utils import load_checkpoint, process_dic
        ckpt = load_checkpoint(args.resume, map_location=dev)
        it_first = ckpt["iteration"] + 1                # the checkpoint holds the last completed iteration
        net.load_state_dict(process_dic(ckpt["model"]))  # reference checkpoints come from the DDP wrapper ('module.')
        optimizer.load_state_dict(ckpt["optimizer"])
        scheduler.load_state_dict(ckpt["scheduler"])
    dataset = SyntheticPepDataset(args.dataset_size, args.pocket, args.peptide, seed=0)
    train_iterator = inf_iterator(make_loader(dataset, bs, rank, world, seed=config.train.seed))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    last = None
    warmup = max(args.warmup, 11) if (args.graph and world > 1) else args.warmup
    graphed = None

    def one_iteration():
        batch = recursive_to(next(train_iterator), dev)
        if graphed is not None:
            return graphed(batch)
        return train_step(model, batch, optimizer, config.train.loss_weights, config.train.max_grad_norm)

    for it in range(it_first, it_first + warmup + args.iters):
        if it == it_first + warmup:
            if args.graph:
                graphed = GraphedTrainStep(model, optimizer, config.train.loss_weights, config.train.max_grad_norm,
                                           recursive_to(next(train_iterator), dev))
                graphed(recursive_to(next(train_iterator), dev))          # one replay outside the timed region
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev0.record()
        last = one_iteration()
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / max(1, args.iters)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nccl = None
    if args.profile > 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(args.profile):
                one_iteration()
            torch.cuda.synchronize()
        if rank == 0:
            nccl = nccl_overlap_summary(prof, args.profile)
    if rank == 0:
        if args.save:
            torch.save(checkpoint_dict(config, model, optimizer, scheduler, it_first + warmup + args.iters - 1),
                       args.save)
        line = json.dumps({"metric": "training samples/sec (flow-matching loss, fwd+bwd+Adam)", "unit": "samples/s",
                           "value": bs * world / (float(ms) / 1e3), "ms_per_iter": float(ms), "n_gpus": world,
                           "batch_per_gpu": bs, "residues": args.pocket + args.peptide,
                           "padded_residues": -(-(args.pocket + args.peptide) // 8) * 8, "iters": args.iters,
                           "cuda_graph": bool(args.graph), "tf32_matmul": bool(args.tf32), "loss": float(last[0]),
                           "grad_norm": float(last[2]), "ddp_allreduce": nccl, "time": time.strftime("%Y-%m-%d %H:%M:%S")})
        print(line)
        if args.out:
            with open(args.out, "a") as fh:
                fh.write(line + "\n")
    if world > 1:
        if args.graph:
            # Measured on the 8 x B200 box (profiles/r2_train_cfg5_ddp.jsonl): with the gradient all-reduces captured in
            # the graph, process-group teardown never returns (every rank had already printed / written its results).
            # Drop the graph, meet once more, and leave without the NCCL destructor.
            import sys
            graphed = None
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
