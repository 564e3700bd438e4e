"""Pin the oracle restatement to golden vectors produced by the unmodified reference (CPU fp32)."""
import pytest
import torch

from helpers import GOLDEN_CASES, GOLDEN_LONG, check_grad_checksums, load_golden, max_rel, r_grad_probe, rebuild_case
from oracle import route_fusion_oracle as orc
from oracle import synth


@pytest.mark.parametrize("name", GOLDEN_CASES + GOLDEN_LONG)
def test_oracle_matches_reference_golden(name):
    gold = load_golden(name)
    c = gold["case"]
    sdm, sdp, sdh, inp = rebuild_case(c)
    for sd in (sdm, sdp, sdh):
        for k in sd:
            sd[k] = sd[k].clone().requires_grad_(True)
    xs = {k: inp[k].clone().requires_grad_(True) for k in ("x_l", "x_n", "x_i")}
    logits, alpha, routes, R = orc.full_forward(
        sdm, sdp, sdh, xs["x_l"], xs["x_n"], xs["x_i"], inp["mL"], inp["mN"], inp["mI"],
        variant=c["variant"], route_mask=inp["route_mask"], act_temperature=c["temp"],
        detach_priors=c["detach"], acts_override=inp.get("acts_override"), num_routing=c.get("iters", 3), layers=c.get("layers", 4))
    routes_t = torch.stack([routes[r] for r in synth.ROUTES], dim=1)
    assert max_rel(routes_t, gold["routes"]) < 2e-5
    assert max_rel(logits, gold["logits"]) < 2e-5
    assert max_rel(alpha, gold["alpha"]) < 2e-5
    assert max_rel(R, gold["R"]) < 2e-5
    loss = synth.loss_fn(logits, inp["y"], c["variant"])
    assert abs(float(loss) - gold["loss"]) < 1e-5
    total = loss + 0.05 * (R * r_grad_probe(c, R.shape)).sum()
    total.backward()
    grads = {}
    for sd in (sdm, sdp, sdh):
        for k, v in sd.items():
            grads[k] = v.grad
    for k, v in xs.items():
        grads[k] = v.grad
    if c.get("override_grad"):
        grads["acts_override"] = inp["acts_override"].grad
    none = sorted(k for k, g in grads.items() if g is None)
    assert none == gold["grad_none"], (none, gold["grad_none"])
    check_grad_checksums(grads, gold["grad_checksum"], 1e-4, name)
    for k, g in gold["grad_full"].items():
        assert max_rel(grads[k], g) < 1e-4, k
