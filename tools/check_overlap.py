"""Checks the overlapped gradient all-reduce (dist.OverlappedGradReducer) against the plain one.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_overlap.py

Every rank runs the same model on its own shard; the averaged gradients produced (a) by the backward followed by
allreduce_gradients and (b) by the backward with per-layer-block all-reduces on a side stream (eagerly and inside
a CUDA graph) must agree.  Prints one line per mode; exits non-zero on a mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import bench
    from multimodalrouting_b200.dist import OverlappedGradReducer, allreduce_gradients
    from multimodalrouting_b200.graphs import GraphedStep
    from multimodalrouting_b200 import synth
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("B", "128"))
    rh, mult, proj, head, _ = bench.build_models(dev)
    modules = (mult, proj, head)
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in
           synth.make_inputs(B=B, K=bench.K_LABELS, seed=100 + rank).items()}
    adapter = rh.RouteDimAdapter(256, 256, 256, 256)
    lossf = torch.nn.BCEWithLogitsLoss()

    def fwd_bwd():
        for m in modules:
            m.zero_grad(set_to_none=True)
        xs = [inp[k].detach().requires_grad_(True) for k in ("x_l", "x_n", "x_i")]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, _, _, _ = rh.forward_capsule_from_multmodel(
                mult, xs[0], xs[1], xs[2], proj, head, mL=inp["mL"], mN=inp["mN"], mI=inp["mI"],
                route_adapter=adapter, route_mask=inp["route_mask"])
        loss = lossf(logits.float(), inp["y"])
        loss.backward()
        return loss

    def snapshot():
        torch.cuda.synchronize()
        return {f"{i}.{n}": p.grad.detach().clone() for i, m in enumerate(modules) for n, p in m.named_parameters()
                if p.grad is not None}

    fwd_bwd()
    allreduce_gradients(modules, world)
    ref = snapshot()

    red = OverlappedGradReducer(mult, (proj, head), bench.LAYERS)

    def step():
        loss = fwd_bwd()
        red.finish()
        return loss

    def compare(tag, got):
        worst, where = 0.0, ""
        assert got.keys() == ref.keys()
        for k, g in got.items():
            r = ref[k]
            e = float((g - r).abs().max() / (r.abs().max() + 1e-20))
            if e > worst:
                worst, where = e, k
        # split-K weight gradients accumulate with atomics: run-to-run differences of a few ulp are expected
        ok = worst < 1e-3
        if rank == 0:
            print(f"[check_overlap] {tag}: worst rel diff {worst:.3e} at {where} -> {'ok' if ok else 'MISMATCH'}",
                  flush=True)
        return ok

    step()
    ok = compare("eager", snapshot())
    g = GraphedStep(step, warmup=2)
    g()
    ok = compare("cuda-graph", snapshot()) and ok
    g()
    ok = compare("cuda-graph replay 2", snapshot()) and ok
    red.close()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    bad = int(flag.item())
    torch.cuda.synchronize()
    sys.stdout.flush()
    # collectives captured in a live CUDA graph: destroy_process_group would wait forever (see bench._teardown)
    os._exit(1 if bad else 0)


if __name__ == "__main__":
    main()
