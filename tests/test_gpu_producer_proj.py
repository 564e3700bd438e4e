"""-m gpu: route-input projections (BERT chunk LayerNorm + Linear 768 -> 256, CXR token_proj
512 -> 256) against goldens produced by executing the reference's own module definitions (oracle/gen_golden_tail.py)."""
import os

import pytest
import torch

from helpers import GOLD, max_rel, rel_err

pytestmark = pytest.mark.gpu
CASES = ["chunk768", "chunk768_ragged", "token512"]


def _module(gold, key):
    from multimodalrouting_b200 import producers
    if key == "chunk":
        m = producers.NoteChunkProjector(768, 256)
        m.proj.load_state_dict(gold["state"][key])
    else:
        m = producers.ImageTokenProjector(512, 256)
        m.token_proj.load_state_dict(gold["state"][key])
    return m.cuda()


def _wgrad_check(got, ref, tol):
    pv = torch.randn(got.numel(), generator=torch.Generator().manual_seed(77)).double()
    g = got.detach().double().cpu()
    assert max_rel(g[:8], ref["head"]) < tol
    assert abs(float(g.norm()) - ref["norm"]) < tol * ref["norm"]
    assert abs(float(g.flatten() @ pv) - ref["proj"]) < 6 * tol * ref["norm"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16_in"])
def test_projection_matches_reference_golden(name, mode):
    gold = torch.load(os.path.join(GOLD, "tail_proj.pt"), weights_only=False)
    c = gold[name]
    mod = _module(gold, c["module"])
    x = c["x"].cuda()
    if mode == "bf16_in":                 # an encoder that already emits bf16 is consumed without a cast pass
        x = x.bfloat16()
    x.requires_grad_(True)
    tol = 1e-4 if mode == "fp32" else 2e-2
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode != "fp32"):
        y = mod(x)
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(c["y"].shape)
    assert max_rel(y, c["y"]) < tol
    y.backward(c["dy"].cuda())
    assert x.grad.dtype == x.dtype
    assert rel_err(x.grad.float(), c["dx"]) < (5e-4 if mode == "fp32" else 3e-2)
    params = dict(mod.proj.named_parameters()) if c["module"] == "chunk" else dict(mod.token_proj.named_parameters())
    for k, ref in c["grads"].items():
        if isinstance(ref, dict):
            _wgrad_check(params[k].grad, ref, 5e-4 if mode == "fp32" else 2e-2)
        else:
            assert rel_err(params[k].grad, ref) < (5e-4 if mode == "fp32" else 3e-2), k


def test_projection_feeds_the_hot_path():
    """chunk projection -> sanitize -> MULTModel with gradients flowing back into the projector."""
    import multimodalrouting_b200 as mmr
    from multimodalrouting_b200 import producers
    g = torch.Generator().manual_seed(3)
    B = 3
    proj = producers.NoteChunkProjector(768, 256).cuda()
    tok = producers.ImageTokenProjector(512, 256).cuda()
    mult = mmr.MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, 2, 0, 0., 0., 0., 0., 0., 0., 0., False).cuda()
    chunks = torch.randn(B, 6, 768, generator=g).cuda()
    fmap = torch.randn(B, 9, 512, generator=g).cuda()
    x_l = torch.randn(B, 7, 256, generator=g).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        x_n = producers._sanitize_encoder_out({"seq": proj(chunks)}, "N")["seq"]
        x_i = producers._sanitize_encoder_out({"seq": tok(fmap)}, "I")["seq"]
        routes = mult(x_l, x_n, x_i)
    sum(r.float().square().mean() for r in routes.values()).backward()
    for p in list(proj.parameters()) + list(tok.parameters()):
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()) and float(p.grad.abs().max()) > 0
