"""-m gpu: the packed-weight cache of MULTModel (compute-type weight copies rebuilt only when a parameter changed)."""
import pytest
import torch

from gpu_common import build_modules, to_dev
from helpers import max_rel, rebuild_case

pytestmark = pytest.mark.gpu

CASE = dict(variant="pheno", K=5, orig_d_n=256, B=4, seed=31, sharp=1.0, temp=1.0, detach=False, missing=True,
            mask_mode="full")


def _fwd(mult, d, autocast=True):
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = mult(d["x_l"], d["x_n"], d["x_i"], mL=d["mL"], mN=d["mN"], mI=d["mI"])
    return torch.stack([out[r] for r in ("L", "LN", "IN", "LNI")])


@pytest.mark.parametrize("autocast", [True, False])
def test_repacks_exactly_when_a_parameter_changes(autocast):
    from multimodalrouting_b200 import _lib, optim
    lib = _lib.load()
    sdm, sdp, sdh, inp = rebuild_case(CASE)
    _, mult, proj, head = build_modules(CASE, sdm, sdp, sdh)
    d = to_dev(inp)
    with torch.no_grad():
        a = _fwd(mult, d, autocast)
        n0 = lib.mmr_launch_count()
        b = _fwd(mult, d, autocast)
        n_cached = lib.mmr_launch_count() - n0
        assert torch.equal(a, b)
        # an in-place update bumps the tensor version: the next forward re-packs and sees the new weights
        mult.trans_l_with_n.layers[0].fc1.weight.mul_(1.5)
        n0 = lib.mmr_launch_count()
        c = _fwd(mult, d, autocast)
        n_repacked = lib.mmr_launch_count() - n0
    assert n_repacked > n_cached, (n_repacked, n_cached)          # pack / bias-fold kernels ran again
    _, fresh, _, _ = build_modules(CASE, {k: v.clone() for k, v in mult.state_dict().items()}, sdp, sdh)
    with torch.no_grad():
        ref = _fwd(fresh.cuda(), d, autocast)
    assert torch.equal(c, ref) and not torch.equal(c, a)
    # the fused optimizer writes through raw pointers and bumps the versions itself
    out = _fwd(mult, d, autocast)
    out.square().mean().backward()
    opt = optim.FusedAdamW([p for p in mult.parameters() if p.grad is not None], lr=1e-2)
    opt.step()
    with torch.no_grad():
        e = _fwd(mult, d, autocast)
        _, fresh2, _, _ = build_modules(CASE, {k: v.clone() for k, v in mult.state_dict().items()}, sdp, sdh)
        ref2 = _fwd(fresh2.cuda(), d, autocast)
    assert torch.equal(e, ref2) and not torch.equal(e, c)


def test_graph_with_static_weights_and_refresh():
    """A captured forward with static_weights reads the cached buffer; after an update outside the graph
    refresh_packed_weights() re-packs into the same buffer and the replay sees the new weights."""
    sdm, sdp, sdh, inp = rebuild_case(CASE)
    _, mult, _, _ = build_modules(CASE, sdm, sdp, sdh)
    d = to_dev(inp)
    mult.static_weights = True
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            eager = _fwd(mult, d)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = _fwd(mult, d)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)
        mult.trans_i_with_l.layers[3].fc2.weight.mul_(0.5)
        mult.refresh_packed_weights()
        g.replay()
        torch.cuda.synchronize()
        mult.static_weights = False
        assert max_rel(out, _fwd(mult, d)) == 0.0 and not torch.equal(out, eager)
