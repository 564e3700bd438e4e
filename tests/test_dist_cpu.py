"""CPU (gloo, world_size 2) tests of the data-parallel plumbing: patient sharding and the flat-buffer
gradient all-reduce used by bench.py --gpus N.  The arithmetic under test is the
host-side logic only; gradients come from the oracle, not from the CUDA path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from multimodalrouting_b200.dist import _flat_groups, allreduce_gradients, shard_range  # noqa: E402


def test_shard_range_partitions_every_patient_once():
    for n in (1, 7, 16, 512, 8191, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_flat_groups_cover_views_of_one_buffer():
    lin = torch.nn.Linear(4, 3)
    flat = torch.arange(20, dtype=torch.float32)
    lin.weight.grad = flat[:12].view(3, 4)
    lin.bias.grad = flat[12:15]
    other = torch.nn.Parameter(torch.zeros(5))
    other.grad = torch.ones(5)
    groups = _flat_groups([lin.weight, lin.bias, other])
    assert len(groups) == 2
    assert sorted(g.numel() for g in groups) == [5, 15]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import route_fusion_oracle as orc
        from oracle import synth
        torch.set_num_threads(1)
        K, B = 3, 4
        sdm, sdp, sdh = synth.make_state(K=K, seed=11)
        inp = synth.make_inputs(B=B, K=K, seed=12, TL=6, TN=4, TI=5)
        lo, hi = shard_range(B, rank, world)

        def grads(sl):
            a, b, h = [{k: v.clone().requires_grad_(True) for k, v in sd.items()} for sd in (sdm, sdp, sdh)]
            logits, _, _, _ = orc.full_forward(a, b, h, inp["x_l"][sl], inp["x_n"][sl], inp["x_i"][sl], inp["mL"][sl],
                                               inp["mN"][sl], inp["mI"][sl], variant="pheno",
                                               route_mask=inp["route_mask"][sl])
            synth.loss_fn(logits, inp["y"][sl], "pheno").backward()
            return a, b, h

        a, b, h = grads(slice(lo, hi))
        # emulate the fused backward: the gradients of a module are views of ONE flat buffer
        mods = []
        for sd in (a, b, h):
            m = torch.nn.Module()
            ps = [(k, v) for k, v in sd.items() if v.grad is not None]
            flat = torch.cat([v.grad.flatten() for _, v in ps])
            o = 0
            for i, (k, v) in enumerate(ps):
                p = torch.nn.Parameter(v.detach().clone())
                p.grad = flat[o:o + v.numel()].view(v.shape)
                o += v.numel()
                m.register_parameter(f"p{i}", p)
            mods.append((m, [k for k, _ in ps]))
        n = allreduce_gradients([m for m, _ in mods], world)
        assert n == 3, n      # one collective per module (flat buffer), not one per parameter
        if rank == 0:
            fa, fb, fh = grads(slice(0, B))      # single-process gradients of the whole batch
            worst = 0.0
            for (m, names), full in zip(mods, (fa, fb, fh)):
                for i, k in enumerate(names):
                    g = getattr(m, f"p{i}").grad
                    ref = full[k].grad
                    worst = max(worst, float((g - ref).abs().max() / (ref.abs().max() + 1e-12)))
            out.put(worst)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_matches_single_process():
    """Equal shards + mean-reduced loss: the averaged per-rank gradients equal the full-batch gradients."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    worst = out.get()
    assert worst < 1e-4, worst


def test_grad_layout_buckets_follow_backward_completion_order():
    """ops.grad_layout: the per-layer blocks that OverlappedGradReducer all-reduces during the backward are
    contiguous, disjoint, ordered last layer first, and hold exactly the 'early' parameter kinds."""
    from multimodalrouting_b200 import MULTModel
    from multimodalrouting_b200 import ops
    layers = 4
    mult = MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, layers, 0, 0., 0., 0., 0., 0., 0., 0., False)
    names = [n for n, _ in mult.named_parameters()]
    shapes = [tuple(p.shape) for _, p in mult.named_parameters()]
    offs, total, groups, buckets = ops.grad_layout(shapes, layers)
    numel = [int(torch.Size(s).numel()) for s in shapes]
    # every parameter has its own 16-byte aligned, non-overlapping slot
    spans = sorted((o, o + n) for o, n in zip(offs, numel))
    assert all(o % 4 == 0 for o, _ in spans)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= total
    assert len(buckets) == layers + 1
    assert buckets[0][0] == 0 and buckets[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
    early = (".self_attn.out_proj.", ".fc1.", ".fc2.", ".layer_norms.1.")
    for i, (lo, hi) in enumerate(buckets[:-1]):
        layer = layers - 1 - i                      # completion order of the backward
        inside = [names[j] for j in range(len(names)) if lo <= offs[j] < hi]
        assert len(inside) == 6 * 8
        assert all(f".layers.{layer}." in n and any(e in n for e in early) for n in inside), inside[:3]
    lo, hi = buckets[-1]
    late = [names[j] for j in range(len(names)) if lo <= offs[j] < hi]
    assert all(("in_proj" in n or "layer_norms.0" in n or ".layers." not in n) for n in late)
    # stacked groups are dense: cnt * numel floats from their start
    for start, idx, sh in groups:
        assert [offs[j] for j in idx] == [start + k * numel[idx[0]] for k in range(len(idx))]
