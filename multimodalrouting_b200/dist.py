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
    if nccl and len(flats) > 1:
        # the flat buffers of the modules (route fusion: 19.7 M floats; projector + head: one more) go out as ONE NCCL group
        # (ncclGroupStart / End): one launch instead of one per module
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager
            cm = _coalescing_manager(group=group, device=flats[0].device, async_ops=False)
        except (ImportError, TypeError):
            cm = None              # older / different torch: one call per buffer below
        if cm is not None:
            with cm:               # (an error inside propagates: falling through would average twice)
                for f in flats:
                    dist.all_reduce(f, op=dist.ReduceOp.AVG, group=group)
            return 1
    for f in flats:
        if nccl:                   # averaged inside the collective: no extra pass over the 79 MB buffer
            dist.all_reduce(f, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group)
            f.mul_(1.0 / world)
    return len(flats)


class OverlappedGradReducer:
    """Gradient all-reduce overlapped with the backward of the route-fusion op.

    The fused backward lays the gradients out in completion order (ops.grad_layout): the block of layer l
    (out_proj / fc1 / fc2 / LayerNorm-1 of all six encoders, ~3.5 M floats) is final long before the backward
    ends.  The C library records a CUDA event after each block; this class waits for it on a side stream and
    all-reduces the block there (NCCL over NVLink) while layers l-1 .. 0 are still being differentiated.  Only
    the late gradients (in_proj / LayerNorm-0 / small tensors, ~25 %) and the other modules are reduced after
    the backward (`finish`).  Works eagerly and inside CUDA-graph capture (the side stream is forked from and
    joined back into the capturing stream)."""

    def __init__(self, mult_module: torch.nn.Module, other_modules: Iterable[torch.nn.Module], layers: int,
                 group=None):
        from . import ops
        self.mult, self.others = mult_module, list(other_modules)
        self.group = group
        self.world = dist.get_world_size(group)
        self.comm = torch.cuda.Stream()
        self.events = [torch.cuda.Event() for _ in range(layers)]
        for e in self.events:          # CUDA events are created lazily: force the handles to exist
            e.record()
        torch.cuda.synchronize()
        self._flat = None
        self._rest = None
        self._mult_params = [p for p in self.mult.parameters() if p.requires_grad]
        ops.set_grad_overlap(self._on_fusion_grads, self.events)     # raises if another reducer holds the hook

    def close(self):
        from . import ops
        ops.set_grad_overlap(None, None)

    def _reduce(self, t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)

    def _on_fusion_grads(self, flat: torch.Tensor, buckets):
        """Called by the autograd node right after the backward kernels were enqueued."""
        # The slices of `flat` are reduced IN PLACE on the side stream; that is only the gradient if autograd then installs
        # these very views as p.grad.  With a pre-existing .grad (gradient accumulation, zero_grad(set_to_none=False))
        # autograd would instead do p.grad += view on the main stream, racing with the collective and dropping the average.
        stale = [i for i, p in enumerate(self._mult_params) if p.grad is not None]
        if stale:
            raise RuntimeError(
                f"OverlappedGradReducer: {len(stale)} MULTModel parameters already hold a .grad when the backward runs; the "
                "overlapped all-reduce needs zero_grad(set_to_none=True) before every backward (no gradient accumulation). "
                "Use dist.allreduce_gradients(...) after the backward instead.")
        early, rest = buckets[:-1], buckets[-1]
        # bucket i holds layer (layers-1-i): completion order
        n = len(early)
        for i, (lo, hi) in enumerate(early):
            ev = self.events[n - 1 - i]
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                self._reduce(flat[lo:hi])
        self._flat, self._rest = flat, rest

    def finish(self):
        """Reduces what is left (late fusion gradients, projector, head) and joins the side stream."""
        cur = torch.cuda.current_stream()
        if self._flat is not None:
            base = self._flat.untyped_storage().data_ptr()
            foreign = [p for p in self._mult_params if p.grad is not None and p.grad.untyped_storage().data_ptr() != base]
            if foreign:
                raise RuntimeError(f"OverlappedGradReducer: {len(foreign)} gradients are not views of the reduced flat buffer "
                                   "(something replaced or accumulated into .grad); the averaged values did not reach them")
            lo, hi = self._rest
            if hi > lo:
                self._reduce(self._flat[lo:hi])
            self._flat = None
        else:                          # the fusion op did not run under the hook: plain path
            for f in _flat_groups(list(self.mult.parameters())):
                self._reduce(f)
        for f in _flat_groups([p for m in self.others for p in m.parameters()]):
            self._reduce(f)
        cur.wait_stream(self.comm)
