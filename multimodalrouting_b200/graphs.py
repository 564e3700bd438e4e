"""CUDA-graph capture of a whole route-fusion + routing training step.

The hot path is ~130 kernel launches per step; at the reference's batch sizes the Python / autograd /
launch overhead of issuing them (≈5 ms) is as long as the GPU work itself.  Shapes are static, every
buffer is caller-owned and the C ABI only enqueues work on the current stream, so the entire
forward + backward can be captured once and replayed with a single launch.

    step = GraphedStep(fn)      # fn() runs forward + backward on STATIC input tensors, returns the loss
    loss = step()               # replay; parameter .grad tensors and `loss` are updated in place

Copy new inputs into the static tensors before calling `step()`.  Gradients live in the graph's
private memory pool: read or all-reduce them after the replay, do not free them.
"""
from __future__ import annotations

import os
from typing import Callable

import torch


class GraphedStep:
    def __init__(self, fn: Callable[[], torch.Tensor], warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device")
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up off the default stream (lazy init, allocator, autotune-free)
            for _ in range(max(warmup, 1)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # capture on a high-priority stream (MMR_GRAPH_PRIORITY=0: default priority) so that the chain's kernels are preferred over
        # the (default-priority) weight-gradient side stream when both have CTAs waiting for an SM
        prio = os.environ.get("MMR_GRAPH_PRIORITY", "1") == "1"
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(priority=-1) if prio else None):
            self.out = fn()
        torch.cuda.synchronize()

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.out
