"""Data-parallel plumbing: patients shard across ranks, weights are replicated, and the only
exchange is one gradient all-reduce per step (NCCL over NVLink / NVSwitch on the GPU box; gloo in
the CPU tests).  The reference is single-process; this is the B200-native
equivalent of running it under DDP.

The fused backward returns every parameter gradient of a module as a view of ONE flat buffer, so
the all-reduce runs on a handful of large contiguous tensors (no per-parameter bucketing)."""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous patient range [lo, hi) of `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _flat_groups(params: Iterable[torch.nn.Parameter]) -> List[torch.Tensor]:
    """Groups .grad tensors by underlying storage and returns one 1-D covering view per storage."""
    by_storage = {}
    for p in params:
        g = p.grad
        if g is None:
            continue
        if not g.is_contiguous():
            raise RuntimeError("gradient all-reduce expects contiguous gradients")
        key = g.untyped_storage().data_ptr()
        lo = g.storage_offset()
        hi = lo + g.numel()
        ent = by_storage.get(key)
        if ent is None:
            by_storage[key] = [g, lo, hi]
        else:
            ent[1] = min(ent[1], lo)
            ent[2] = max(ent[2], hi)
    out = []
    for g, lo, hi in by_storage.values():
        out.append(torch.as_strided(g, (hi - lo,), (1,), lo))
    return out


def allreduce_gradients(modules: Iterable[torch.nn.Module], world_size: int = None, group=None) -> int:
    """Averages gradients over ranks in place.  Returns the number of collectives issued."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = world_size or dist.get_world_size(group)
    if world == 1:
        return 0
    params = [p for m in modules for p in m.parameters()]
    flats = _flat_groups(params)
    nccl = dist.get_backend(group) == "nccl"
    for f in flats:
        if nccl:                   # averaged inside the collective: no extra pass over the 79 MB buffer
            dist.all_reduce(f, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group)
            f.mul_(1.0 / world)
    return len(flats)
