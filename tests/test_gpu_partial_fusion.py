"""-m gpu: the attention-fusion route constructors of the Partial/ variant (multimodalrouting_b200/partial_fusion.py: the
Linear layers on the tensor-core GEMM op, the attention core on the hot path's attention kernels through mmr_attention_fwd/bwd)
against goldens produced by the unmodified reference modules (oracle/gen_golden_partial.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodalrouting_b200 import synth                      # noqa: E402
from test_partial_oracle import GOLD, compare, run_oracle      # noqa: E402

pytestmark = pytest.mark.gpu


def run_module(c, autocast):
    from multimodalrouting_b200 import partial_fusion as pf
    sd = synth.make_fusion_state(c["kind"], c["seed"])
    inp = synth.make_fusion_inputs(c["B"], c["TL"], c["TN"], c["TI"], c["seed"] + 1000)
    g = torch.Generator().manual_seed(c["seed"] + 2000)
    probe = torch.randn(c["B"], 256, generator=g).cuda()
    if c["kind"] == "cross":
        mod = pf.CrossAttentionFusion(256, 8, 0.0, c["pool"])
    else:
        mod = pf.TriTokenAttentionFusion(256, 8, 0.0)
    mod.load_state_dict(sd, strict=True)            # the reference's keys, nothing missing or unexpected
    mod = mod.cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        if c["kind"] == "cross":
            ins = {"A": inp[c["TA"]].cuda().requires_grad_(True), "B": inp[c["TB"]].cuda().requires_grad_(True)}
            out = mod(ins["A"], inp["m" + c["TA"]].cuda(), ins["B"], inp["m" + c["TB"]].cuda())
        else:
            ins = {k: inp[k].cuda().requires_grad_(True) for k in ("L", "N", "I")}
            out = mod(ins["L"], inp["mL"].cuda(), ins["N"], inp["mN"].cuda(), ins["I"], inp["mI"].cuda())
    (out.float() * probe).sum().backward()
    torch.cuda.synchronize()
    return out, {k: v.grad for k, v in ins.items()}, {k: p.grad for k, p in mod.named_parameters()}, inp


@pytest.mark.parametrize("name", sorted(GOLD))
def test_fp32_matches_reference_golden(name):
    c = GOLD[name]["case"]
    out, d_in, d_param, _ = run_module(c, autocast=False)
    compare(name, out, d_in, d_param, 1e-4, 5e-4)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_bf16_within_budget(name):
    """Under autocast: outputs within 2e-2 of the reference golden, every gradient within 8e-2 norm-wise of the fp64 oracle."""
    c = GOLD[name]["case"]
    out, d_in, d_param, _ = run_module(c, autocast=True)
    gold = GOLD[name]
    err = float((out.detach().double().cpu() - gold["out"].double()).abs().max() / gold["out"].abs().max())
    assert err < 2e-2, f"{name} out: {err:.2e}"
    o64, ins64, sd64 = run_oracle(c, torch.float64)
    for k, v in ins64.items():
        e = float((d_in[k].double().cpu() - v.grad).norm() / (v.grad.norm() + 1e-12))
        assert e < 8e-2, f"{name} d {k}: {e:.2e}"
    for k, v in sd64.items():
        e = float((d_param[k].double().cpu() - v.grad).norm() / (v.grad.norm() + 1e-12))
        assert e < 8e-2, f"{name} d {k}: {e:.2e}"


def test_sample_without_valid_keys_contributes_out_of_zero():
    """mN[1] = 0 in the seeded inputs: that sample's L<-N route is out(0) = Linear(LayerNorm(0)) and gets no input gradient."""
    name = "cross_mean"
    c = GOLD[name]["case"]
    out, d_in, _, inp = run_module(c, autocast=False)
    assert float(inp["mN"][1].sum()) == 0.0
    assert bool(torch.isfinite(out).all())
    assert float(d_in["A"][1].abs().max()) == 0.0 and float(d_in["B"][1].abs().max()) == 0.0
    assert float((out[1].cpu() - GOLD[name]["out"][1]).abs().max()) < 1e-5


def test_make_route_inputs_and_build_fusions():
    from multimodalrouting_b200 import partial_fusion as pf
    inp = synth.make_fusion_inputs(4, 12, 6, 9, 5)
    fus = pf.build_fusions(256, device="cuda")
    assert sorted(fus) == ["IL", "IN", "LI", "LN", "LNI", "NI", "NL"]
    z = {m: {"seq": inp[m].cuda(), "mask": inp["m" + m].cuda(), "pool": pf.masked_mean(inp[m].cuda(), inp["m" + m].cuda())}
         for m in ("L", "N", "I")}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        routes = pf.make_route_inputs(z, fus)
    assert list(routes) == ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]
    assert all(tuple(v.shape) == (4, 256) and bool(torch.isfinite(v).all()) for v in routes.values())
