"""CPU: the restatement of the Partial/ attention-fusion modules (oracle/partial_oracle.py) against goldens produced by the
unmodified reference modules (oracle/gen_golden_partial.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodalrouting_b200 import synth            # noqa: E402
from oracle import partial_oracle as po             # noqa: E402

GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "partial_fusion.pt"), weights_only=False)


def run_oracle(c, dtype=torch.float32):
    sd = {k: v.to(dtype).requires_grad_(True) for k, v in synth.make_fusion_state(c["kind"], c["seed"]).items()}
    inp = synth.make_fusion_inputs(c["B"], c["TL"], c["TN"], c["TI"], c["seed"] + 1000)
    g = torch.Generator().manual_seed(c["seed"] + 2000)
    probe = torch.randn(c["B"], 256, generator=g).to(dtype)
    if c["kind"] == "cross":
        ins = {"A": inp[c["TA"]].to(dtype).requires_grad_(True), "B": inp[c["TB"]].to(dtype).requires_grad_(True)}
        out = po.cross_attention_fusion(sd, ins["A"], inp["m" + c["TA"]], ins["B"], inp["m" + c["TB"]], c["pool"])
    else:
        ins = {k: inp[k].to(dtype).requires_grad_(True) for k in ("L", "N", "I")}
        out = po.tri_token_fusion(sd, ins["L"], inp["mL"], ins["N"], inp["mN"], ins["I"], inp["mI"])
    (out * probe).sum().backward()
    return out, ins, sd


def check_group(got: dict, ref: dict, tol: float, what: str):
    """Whole tensors where the golden holds them, (norm, projection) checksums for the large ones; the random directions are
    drawn in the generator's order."""
    gq = torch.Generator().manual_seed(99) if not hasattr(check_group, "_g") else check_group._g
    check_group._g = gq
    for k, r in ref.items():
        g = got[k]
        assert g is not None, f"{what} {k}: no gradient"
        if isinstance(r, dict):
            d = torch.randn(r["shape"], generator=gq)
            gd = g.detach().double().cpu()
            assert abs(float(gd.norm()) - r["norm"]) <= tol * max(r["norm"], 1e-12), f"{what} {k}: norm"
            assert abs(float((gd * d.double()).sum()) - r["proj"]) <= tol * max(r["norm"] * float(d.norm()), 1e-12), f"{what} {k}: proj"
        else:
            err = float((g.detach().double().cpu() - r.double()).abs().max() / (r.double().abs().max() + 1e-12))
            assert err < tol, f"{what} {k}: {err:.2e}"


def compare(name, out, d_in, d_param, tol_out, tol_grad):
    gold = GOLD[name]
    err = float((out.detach().double().cpu() - gold["out"].double()).abs().max() / gold["out"].abs().max())
    assert err < tol_out, f"{name} out: {err:.2e}"
    check_group._g = torch.Generator().manual_seed(99)
    check_group(d_in, gold["d_in"], tol_grad, f"{name} d_in")
    check_group(d_param, gold["d_param"], tol_grad, f"{name} d_param")


@pytest.mark.parametrize("name", sorted(GOLD))
def test_partial_oracle_matches_reference_golden(name):
    c = GOLD[name]["case"]
    out, ins, sd = run_oracle(c)
    compare(name, out, {k: v.grad for k, v in ins.items()}, {k: v.grad for k, v in sd.items()}, 2e-5, 2e-4)


def test_partial_fusion_modules_surface_and_loud_cpu_failure():
    """CPU: the drop-in modules expose exactly the reference's parameter names (the goldens' d_param keys come from the
    unmodified reference modules) and refuse to run without CUDA instead of falling back."""
    from multimodalrouting_b200 import partial_fusion as pf
    cross, tri = pf.CrossAttentionFusion(256, 8, 0.0, "mean"), pf.TriTokenAttentionFusion(256, 8, 0.0)
    assert sorted(dict(cross.named_parameters())) == sorted(GOLD["cross_mean"]["d_param"])
    assert sorted(dict(tri.named_parameters())) == sorted(GOLD["tri"]["d_param"])
    cross.load_state_dict(synth.make_fusion_state("cross", 1), strict=True)
    tri.load_state_dict(synth.make_fusion_state("tri", 2), strict=True)
    inp = synth.make_fusion_inputs(2, 5, 4, 3, 9)
    with pytest.raises(RuntimeError):
        cross(inp["L"], inp["mL"], inp["N"], inp["mN"])
    with pytest.raises(RuntimeError):
        tri(inp["L"], inp["mL"], inp["N"], inp["mN"], inp["I"], inp["mI"])
    with pytest.raises(NotImplementedError):
        pf.CrossAttentionFusion(128, 4)(inp["L"][..., :128], inp["mL"], inp["N"][..., :128], inp["mN"])
    f = pf.build_fusions(256, device="cpu")
    assert sorted(f) == ["IL", "IN", "LI", "LN", "LNI", "NI", "NL"]

