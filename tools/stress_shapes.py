"""BASELINE configs[4] shape stress (L 512 / N 128 / I 196 tokens) and configs[2] (Mort, K=2) sanity + timing
on one GPU: finite outputs/gradients, R normalised over routes, patients/s of fwd+bwd (bf16)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodalrouting_b200 import synth  # noqa: E402
from multimodalrouting_b200 import MULTModel  # noqa: E402


def run(variant, K, B, TL, TN, TI, iters=5):
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    sdm, sdp, sdh = synth.make_state(K=K, seed=3)
    mult = MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
    mult.load_state_dict(sdm); proj.load_state_dict(sdp); head.load_state_dict(sdh)
    mult, proj, head = mult.cuda(), proj.cuda(), head.cuda()
    inp = synth.make_inputs(B=B, K=K, seed=4, TL=TL, TN=TN, TI=TI, missing=True)
    d = {k: v.cuda() for k, v in inp.items()}

    def step():
        for m in (mult, proj, head):
            m.zero_grad(set_to_none=True)
        xs = [d[k].detach().requires_grad_(True) for k in ("x_l", "x_n", "x_i")]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, alpha, routes, R = rh.forward_capsule_from_multmodel(
                mult, xs[0], xs[1], xs[2], proj, head, mL=d["mL"], mN=d["mN"], mI=d["mI"],
                route_adapter=rh.RouteDimAdapter(256, 256, 256, 256), route_mask=d["route_mask"])
        loss = synth.loss_fn(logits, d["y"], variant)
        loss.backward()
        return logits, R, xs
    logits, R, xs = step()
    torch.cuda.synchronize()
    assert torch.isfinite(logits).all() and torch.isfinite(R).all()
    kept = d["route_mask"].sum(1) > 0
    s = R.float().sum(1)[kept]
    assert (s - 1).abs().max() < 1e-3, float((s - 1).abs().max())       # M/main.py:319-338 guard
    for x in xs:
        assert torch.isfinite(x.grad).all()
    for m in (mult, proj, head):
        for n, p in m.named_parameters():
            assert p.grad is None or torch.isfinite(p.grad).all(), n
    t0 = time.time()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / iters
    print(f"{variant} K={K} B={B} T=({TL},{TN},{TI}): {dt*1e3:.2f} ms/step, {B/dt:.0f} patients/s, "
          f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")


if __name__ == "__main__":
    run("pheno", 25, 8, 512, 128, 196)
    run("pheno", 3, 256, 512, 128, 196)       # configs[4]: 2048 patients over 8 GPUs -> 256 per GPU
    run("mort", 2, 1024, 48, 16, 49)          # configs[2]: 8192 patients over 8 GPUs -> 1024 per GPU
