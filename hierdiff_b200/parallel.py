"""Multi-GPU plumbing: one process per GPU, the sample batch sharded over ranks.

Molecules are independent (masks, centre-of-gravity and schedule are per molecule), so the
path has no per-step exchange: ranks agree on the weights once (one broadcast of the flat parameter
buffer from rank 0 over NCCL) and rank 0 collects the per-rank result lists at the end.
"""
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class Context:
    rank: int
    world: int
    local_rank: int
    device: torch.device
    owns_group: bool


def init(backend=None):
    """Join the job ``torchrun`` started (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*); a no-op for one process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    owns = False
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if cuda and (backend or "nccl") == "nccl" else {}
        dist.init_process_group(backend or ("nccl" if cuda else "gloo"), rank=rank, world_size=world, **kw)
        owns = True
    return Context(rank, world, local, device, owns)


def broadcast_parameters(module, ctx, src=0):
    """All ranks adopt rank ``src``'s parameters and buffers: ONE collective on one flat fp32 buffer."""
    if ctx.world == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n
    if hasattr(module, "mark_weights_changed"):   # `.data` writes bump no autograd version: tell the packers
        module.mark_weights_changed()
    return flat.numel() * 4


def shard_count(total, ctx):
    """Number of the ``total`` batches this rank samples (contiguous blocks, remainder to the low ranks)."""
    base, rem = divmod(total, ctx.world)
    return base + (1 if ctx.rank < rem else 0)


def shard_range(total, ctx):
    base, rem = divmod(total, ctx.world)
    start = ctx.rank * base + min(ctx.rank, rem)
    return start, start + base + (1 if ctx.rank < rem else 0)


def gather_results(local, ctx, dst=0):
    """Rank ``dst`` receives every rank's ``(results, names)`` in rank order, merged; others get None."""
    if ctx.world == 1:
        return local
    out = [None] * ctx.world if ctx.rank == dst else None
    dist.gather_object(local, out, dst=dst)
    if ctx.rank != dst:
        return None
    results, names = [], []
    for r, n in out:
        results.extend(r)
        names.extend(n)
    return results, names


def max_over_ranks(value, ctx):
    """Max of a python float over ranks (timing: the job is as slow as its slowest rank)."""
    if ctx.world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(ctx):
    if ctx.world > 1:
        dist.barrier()


def finish(ctx):
    if ctx.owns_group and dist.is_initialized():
        dist.destroy_process_group()
