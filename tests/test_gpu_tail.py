"""-m gpu: the producer epilogue and the training tail through the C ABI against
fixtures produced by the reference's own definitions (oracle/gen_golden_tail.py) and against the CPU oracle."""
import os

import pytest
import torch

from helpers import ROOT
from oracle import synth
from oracle import tail_oracle as to

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30))


# ----------------------------------------------------------------------------- producer epilogue ---
@pytest.mark.parametrize("case", ["seq256", "seq768_bf16", "pool512_f16", "seq12"])
def test_sanitize_matches_reference_golden(case):
    from multimodalrouting_b200 import producers
    g = _load("tail_sanitize.pt")[case]
    x = g["x"].cuda().requires_grad_(True)
    y = producers._sanitize_encoder_out({"seq": x, "mask": torch.ones(x.shape[:-1], dtype=torch.int64, device="cuda")}, "t")
    assert y["seq"].dtype == torch.float32 and y["mask"].dtype == torch.float32
    # fp32 tolerance: only the summation order of the row norm differs from ATen's
    assert torch.allclose(y["seq"].detach().cpu(), g["y"], rtol=2e-6, atol=1e-30)
    (dx,) = torch.autograd.grad(y["seq"], x, g["dy"].cuda())
    assert dx.dtype == x.dtype
    tol = 1e-5 if x.dtype == torch.float32 else 1e-2      # dx is rounded to the encoder's dtype, like autograd does
    assert _rel(dx.float(), g["dx"].float()) < tol
    # non-finite entries: Mort and Pheno flavours, exact zeros / +-1e4 where the reference puts them
    yb, cnt = producers.count_nonfinite(g["x_bad"].cuda(), producers.MODE_CLAMP_NORM, 20.0)
    gold = g["y_bad_mort"]
    assert torch.equal(yb.cpu() == 0, gold == 0)
    assert torch.allclose(yb.cpu(), gold, rtol=2e-6, atol=1e-30)
    # what nan_to_num has to rewrite: the whole row of the NaN (NaN norm) and the two inf entries (inf * 0)
    expect = int((~torch.isfinite(to.clamp_norm(g["x_bad"].float(), 20.0))).sum())
    assert expect == g["x_bad"].shape[-1] + 2 and int(cnt) == expect
    yp = producers._sanitize_encoder_out({"seq": g["x_bad"].cuda()}, "t", variant="pheno")["seq"]
    assert yp.dtype == g["y_bad_pheno"].dtype
    assert torch.equal(yp.cpu(), g["y_bad_pheno"])


def test_sanitize_full_size_properties_and_oracle():
    """BASELINE configs[1] producer shapes: idempotence (a clamped tensor passes through unchanged), norm bound, and
    the CPU oracle on a slice."""
    from multimodalrouting_b200 import producers
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(512, 48, 256, generator=gen) * torch.rand(512, 48, 1, generator=gen) * 4.0
    xd = x.cuda()
    y = producers._clamp_norm(xd, 20.0)
    n = y.norm(dim=-1)
    assert float(n.max()) <= 20.0 * (1 + 1e-5)
    small = x.norm(dim=-1) < 19.9
    assert torch.equal(y.cpu()[small], x[small])                       # rows under the bound are bit-identical
    y2 = producers._clamp_norm(y, 20.0)
    assert torch.allclose(y2, y, rtol=1e-6, atol=0)
    ref = to.sanitize_mort(x[:64])
    assert torch.allclose(y[:64].cpu(), ref, rtol=2e-6, atol=1e-30)
    assert producers._clamp_norm(torch.ones(4, device="cuda")).shape == (4,)     # other ranks pass through


# --------------------------------------------------------------------------------- training tail ---
def test_fused_adamw_ema_matches_reference_golden():
    from multimodalrouting_b200 import optim
    g = _load("tail_adamw_ema.pt")
    model = torch.nn.Module()
    for i, t in enumerate(g["init"]):
        model.register_parameter(f"p{i}", torch.nn.Parameter(t.clone()))
    model.cuda()
    params = list(model.parameters())
    opt = optim.FusedAdamW([{"params": params[:3], "lr": g["lr"], "weight_decay": g["weight_decay"], "name": "enc"},
                            {"params": params[3:], "lr": g["lr"], "weight_decay": g["weight_decay"], "name": "head"}])
    ema = optim.EMA([model], decay=g["ema_decay"])
    for s, grads in enumerate(g["grads"]):
        for p, gr in zip(params, grads):
            p.grad = gr.cuda()
        opt.step(max_norm=g["max_norm"], ema=ema)
        assert optim.grads_are_finite(opt) == (not g["skipped"][s])
        if not g["skipped"][s]:
            assert abs(float(opt.total_norm) - g["norms"][s]) <= 1e-5 * g["norms"][s]
        opt.zero_grad(set_to_none=True)
    assert int(opt.step_count) == g["step"]
    for key, got in (("params", [p.detach() for p in params]),
                     ("exp_avg", [opt.state[p]["exp_avg"] for p in params]),
                     ("exp_avg_sq", [opt.state[p]["exp_avg_sq"] for p in params]),
                     ("ema", [ema.shadow[0][f"p{i}"] for i in range(len(params))])):
        for a, b in zip(got, g[key]):
            assert torch.allclose(a.cpu(), b, rtol=2e-5, atol=1e-9), key
    # checkpoint layout = torch.optim.AdamW's: a torch AdamW can load it and vice versa
    sd = opt.state_dict()
    ref_opt = torch.optim.AdamW([{"params": params[:3]}, {"params": params[3:]}], lr=g["lr"])
    ref_opt.load_state_dict(sd)
    assert float(ref_opt.state[params[0]]["step"]) == g["step"]
    opt2 = optim.FusedAdamW([{"params": params[:3]}, {"params": params[3:]}], lr=g["lr"])
    opt2.load_state_dict(ref_opt.state_dict())
    assert int(opt2.step_count) == g["step"]
    # EMA apply_to / restore round trip
    before = [p.detach().clone() for p in params]
    ema.apply_to()
    assert all(torch.equal(p.detach(), ema.shadow[0][f"p{i}"]) for i, p in enumerate(params))
    ema.restore()
    assert all(torch.equal(p.detach(), b) for p, b in zip(params, before))


def test_fused_tail_on_the_real_modules_matches_oracle_and_graph_capture():
    """Gradients of the real drop-in modules (views of the fused backward's flat buffers), 3 steps of
    clip(0.3) + AdamW + EMA against the CPU oracle; then the same step captured into a CUDA graph."""
    from gpu_common import build_modules, to_dev
    from multimodalrouting_b200 import optim
    c = dict(variant="pheno", K=25, orig_d_n=256, temp=1.0, detach=False)
    sdm, sdp, sdh = synth.make_state(K=25, seed=21, sharp=2.0)
    rh, mult, proj, head = build_modules(c, sdm, sdp, sdh)
    modules = (mult, proj, head)
    inp = to_dev(synth.make_inputs(B=8, K=25, seed=22))
    adapter = rh.RouteDimAdapter(256, 256, 256, 256)

    def fwd_bwd():
        for m in modules:
            m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, _, _, _ = rh.forward_capsule_from_multmodel(
                mult, inp["x_l"], inp["x_n"], inp["x_i"], proj, head, mL=inp["mL"], mN=inp["mN"], mI=inp["mI"],
                route_adapter=adapter, route_mask=inp["route_mask"])
        synth.loss_fn(logits.float(), inp["y"], "pheno").backward()

    params = [p for m in modules for p in m.parameters()]
    opt = optim.FusedAdamW(params, lr=2e-4, weight_decay=1e-4)
    ema = optim.EMA(modules, decay=0.99)
    init = [p.detach().cpu().clone() for p in params]
    used, grads_per_step = None, []
    for _ in range(3):
        fwd_bwd()
        used = [i for i, p in enumerate(params) if p.grad is not None]
        grads_per_step.append([params[i].grad.detach().cpu().clone() for i in used])
        opt.step(max_norm=0.3, ema=ema)
    torch.cuda.synchronize()
    ref = to.train_tail([init[i] for i in used], grads_per_step, 2e-4, (0.9, 0.999), 1e-8, 1e-4, 0.3, 0.99)
    assert abs(float(opt.total_norm) - ref["norms"][-1]) <= 1e-5 * ref["norms"][-1]
    sh = ema.shadow_by_storage()
    for j, i in enumerate(used):
        p = params[i]
        assert torch.allclose(p.detach().cpu(), ref["params"][j], rtol=2e-5, atol=1e-8)
        assert torch.allclose(opt.state[p]["exp_avg_sq"].cpu(), ref["exp_avg_sq"][j], rtol=2e-5, atol=1e-12)
        assert torch.allclose(sh[p.data_ptr()].cpu(), ref["ema"][j], rtol=2e-5, atol=1e-8)
    for i, p in enumerate(params):                       # parameters without gradient stay untouched
        if i not in used:
            assert torch.equal(p.detach().cpu(), init[i])
    # CUDA-graph capture: the whole tail enqueues on the capturing stream and never syncs
    for p in params:
        if p.grad is not None:
            p.grad = p.grad.clone()
    before = [p.detach().clone() for p in params]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step(max_norm=0.3, ema=ema)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step(max_norm=0.3, ema=ema)
    n0 = int(opt.step_count)
    graph.replay()
    torch.cuda.synchronize()
    assert int(opt.step_count) == n0 + 1
    assert any(not torch.equal(p.detach(), b) for p, b in zip(params, before))


# ------------------------------------------------------------------------- route mask from presence ---
def test_route_mask_from_presence_matches_reference_golden():
    from multimodalrouting_b200 import producers
    g = _load("tail_route_mask.pt")
    has = [g[k].cuda() for k in ("hasL", "hasN", "hasI")]
    m = producers.build_route_mask_from_presence(*has)
    assert m.dtype == torch.float32 and torch.equal(m.cpu(), g["mask"])
    assert torch.equal(producers.build_route_mask_from_modalities(*[h.bool() for h in has]).cpu(), g["mask"])
    d = producers.build_route_mask_from_presence(*has, drop_routes=(3, 9))
    ref = g["mask"].clone()
    ref[:, [3, 9]] = 0
    assert torch.equal(d.cpu(), ref)
    with pytest.raises(ValueError):
        producers.build_route_mask_from_presence(has[0], has[1][:5], has[2])
    with pytest.raises(ValueError):
        producers.build_route_mask_from_presence(*has, drop_routes=(10,))
