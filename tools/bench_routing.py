"""Capsule routing alone (route embeddings given): forward + backward time per call, CUDA events.
python tools/bench_routing.py [--B 512 --K 25 --variant pheno --iters 50]   (MMR_RT_SPLIT=0 selects the tile-of-4 kernels)"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalrouting_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=512)
    ap.add_argument("--K", type=int, default=25)
    ap.add_argument("--variant", default="pheno")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    if a.variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    _, sdp, sdh = synth.make_state(K=a.K, seed=5, sharp=1.0)
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=a.K)
    proj.load_state_dict(sdp); head.load_state_dict(sdh)
    proj, head = proj.cuda(), head.cuda()
    g = torch.Generator().manual_seed(1)
    stack = (0.5 * torch.randn(10, a.B, 256, generator=g)).cuda().requires_grad_(True)
    rm = (torch.rand(a.B, 10, generator=g) < 0.8).float().cuda()
    gl = torch.randn(a.B, a.K, generator=g).cuda()

    def step():
        ed = {r: stack[i] for i, r in enumerate(synth.ROUTES)}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not a.fp32):
            l, al, _, R = rh.forward_capsule_from_route_dict(ed, proj, head, route_mask=rm)
        t1 = torch.cuda.Event(enable_timing=True); t1.record()
        (l.float() * gl).sum().backward()
        return t1

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if a.graph:      # device time of the captured forward + backward (no launch gaps): the number that counts inside bench.py
        from multimodalrouting_b200.graphs import GraphedStep

        def fn():
            for p in list(proj.parameters()) + list(head.parameters()) + [stack]:
                p.grad = None
            ed = {r: stack[i] for i, r in enumerate(synth.ROUTES)}
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not a.fp32):
                l, al, _, R = rh.forward_capsule_from_route_dict(ed, proj, head, route_mask=rm)
            (l.float() * gl).sum().backward()
            return l
        gs = GraphedStep(fn)
        for _ in range(5):
            gs()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(a.iters):
            gs()
        t1.record()
        torch.cuda.synchronize()
        print(json.dumps({"B": a.B, "K": a.K, "variant": a.variant, "split": os.environ.get("MMR_RT_SPLIT", "1"),
                          "fwd_bwd_ms_graph": t0.elapsed_time(t1) / a.iters}))
        return
    fw = bw = 0.0
    for _ in range(a.iters):
        t0 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
        t0.record()
        t1 = step()
        t2.record()
        torch.cuda.synchronize()
        fw += t0.elapsed_time(t1); bw += t1.elapsed_time(t2)
    print(json.dumps({"B": a.B, "K": a.K, "variant": a.variant, "split": os.environ.get("MMR_RT_SPLIT", "1"),
                      "fwd_ms_eager": fw / a.iters, "bwd_ms_eager": bw / a.iters}))


if __name__ == "__main__":
    main()
