"""Helpers shared by the -m gpu parity tests: build the drop-in modules from a seeded state."""
import torch

from helpers import rebuild_case  # noqa: F401
from oracle import synth


def build_modules(c, sdm, sdp, sdh, device="cuda"):
    import multimodalrouting_b200 as mmr
    if c["variant"] == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    mult = mmr.MULTModel(c.get("orig_d_l", 0) or 256, c["orig_d_n"], c.get("orig_d_i", 0) or 256, 256, 256, 256, True, True, True, 8,
                         c.get("layers", 4), 0,
                         0., 0., 0., 0., 0., 0., 0., False)
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, c.get("iters", 3), 0.0, "EM", num_classes=c["K"])
    mult.load_state_dict(sdm, strict=True)
    proj.load_state_dict(sdp, strict=True)
    head.load_state_dict(sdh, strict=True)
    return rh, mult.to(device), proj.to(device), head.to(device)


def to_dev(inp, device="cuda"):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in inp.items()}


def run_case(c, sdm, sdp, sdh, inp, *, autocast=False, backward=True, r_probe=None):
    """Forward (+ loss + backward) of the drop-in path on the GPU.  Returns dict of results."""
    rh, mult, proj, head = build_modules(c, sdm, sdp, sdh)
    d = to_dev(inp)
    xs = {k: d[k].clone().requires_grad_(True) for k in ("x_l", "x_n", "x_i")}
    ao = d.get("acts_override")
    if ao is not None:
        ao = ao.detach().clone().requires_grad_(bool(c.get("override_grad")))
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
    with ctx:
        logits, alpha, routes, R = rh.forward_capsule_from_multmodel(
            mult, xs["x_l"], xs["x_n"], xs["x_i"], proj, head, mL=d["mL"], mN=d["mN"], mI=d["mI"],
            route_adapter=rh.RouteDimAdapter(256, 256, 256, 256), route_mask=d["route_mask"],
            act_temperature=c["temp"], detach_priors=c["detach"], acts_override=ao)
    out = {"logits": logits, "alpha": alpha, "R": R,
           "routes": torch.stack([routes[r] for r in synth.ROUTES], dim=1)}
    if backward:
        loss = synth.loss_fn(logits, d["y"], c["variant"])
        total = loss
        if r_probe is not None:
            total = total + 0.05 * (R * r_probe.to(R.device)).sum()
        total.backward()
        grads = {}
        for mod in (mult, proj, head):
            for n, p in mod.named_parameters():
                grads[n] = p.grad
        for k, v in xs.items():
            grads[k] = v.grad
        if ao is not None and ao.requires_grad:
            grads["acts_override"] = ao.grad
        out["loss"] = float(loss.detach())
        out["grads"] = grads
    return out
