"""Data-parallel check on real GPUs (BASELINE configs[2] protocol): the gradients after the
NCCL all-reduce over patient shards equal the single-process gradients of the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tools/check_dp.py

Mort variant (K=2, orig_d_n=768), global batch B (default 256) sharded with dist.shard_range; every rank also runs the
full batch alone as the comparison.  fp32 kernels: tolerance 1e-4 relative to the largest entry of each gradient;
bf16 kernels: 2e-2.  Prints one line per mode, exits non-zero on a mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from multimodalrouting_b200 import MULTModel
    from multimodalrouting_b200.MortModel import routing_and_heads as rh
    from multimodalrouting_b200.dist import allreduce_gradients, shard_range
    from multimodalrouting_b200 import synth
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("B", "256"))
    assert B % world == 0, "equal shards: the mean-reduced loss then averages exactly"
    sdm, sdp, sdh = synth.make_state(K=2, seed=42, sharp=2.0, orig_d_n=768)
    mult = MULTModel(256, 768, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    proj, head = rh.RoutePrimaryProjector(256, 32), rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=2)
    mult.load_state_dict(sdm); proj.load_state_dict(sdp); head.load_state_dict(sdh)
    modules = (mult.to(dev), proj.to(dev), head.to(dev))
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v)
           for k, v in synth.make_inputs(B=B, K=2, seed=3042, d_n=768, missing=True).items()}
    adapter = rh.RouteDimAdapter(256, 256, 256, 256)

    def grads(sl, autocast):
        for m in modules:
            m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            logits, _, _, _ = rh.forward_capsule_from_multmodel(
                modules[0], inp["x_l"][sl], inp["x_n"][sl], inp["x_i"][sl], modules[1], modules[2],
                mL=inp["mL"][sl], mN=inp["mN"][sl], mI=inp["mI"][sl], route_adapter=adapter,
                route_mask=inp["route_mask"][sl], act_temperature=1.2)
        synth.loss_fn(logits.float(), inp["y"][sl], "mort").backward()

    def snapshot():
        torch.cuda.synchronize()
        return {f"{i}.{n}": p.grad.detach().clone() for i, m in enumerate(modules) for n, p in m.named_parameters()
                if p.grad is not None}

    ok = True
    for name, autocast, tol in (("fp32", False, 1e-4), ("bf16", True, 2e-2)):
        grads(slice(0, B), autocast)
        full = snapshot()
        lo, hi = shard_range(B, rank, world)
        grads(slice(lo, hi), autocast)
        n = allreduce_gradients(modules, world)
        got = snapshot()
        worst, where = 0.0, ""
        assert got.keys() == full.keys()
        for k in full:
            e = float((got[k] - full[k]).abs().max() / (full[k].abs().max() + 1e-20))
            if e > worst:
                worst, where = e, k
        good = worst < tol
        ok = ok and good
        if rank == 0:
            print(f"[check_dp] {name}: world {world}, global batch {B}, {n} collectives, worst rel diff {worst:.3e} at {where} "
                  f"(tol {tol:g}) -> {'ok' if good else 'MISMATCH'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    bad = int(flag.item())
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
