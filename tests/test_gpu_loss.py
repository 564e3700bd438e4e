"""-m gpu: the loss tail (csrc/loss.cuh) through the C ABI against fixtures produced by the
reference's own statements / definitions (oracle/gen_golden_tail.py: loss_cases) and against the CPU oracle at full size."""
import os

import pytest
import torch

from helpers import ROOT
from oracle import tail_oracle as to

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _load():
    return torch.load(os.path.join(GOLD, "tail_loss.pt"), weights_only=False)


def _close(a, b, tol=3e-6):
    a, b = float(torch.as_tensor(a).detach()), float(torch.as_tensor(b).detach())
    return abs(a - b) <= tol * max(1.0, abs(b))


def _cu(t):
    return None if t is None else t.cuda()


@pytest.mark.parametrize("idx", range(4))
def test_mort_loss_matches_reference_golden(idx):
    from multimodalrouting_b200 import losses
    g = _load()["mort"][idx]
    lg = g["logits"].cuda().requires_grad_(True)
    out = losses.mort_train_loss(lg, g["y"].cuda(), g["prim_acts"].cuda(), label_smoothing=g["label_smoothing"],
                                 route_entropy_lambda=g["lam_ent"], route_entropy_warmup_epochs=g["warm_ent"],
                                 route_uniform_lambda=g["lam_uni"], route_uniform_warmup_epochs=g["warm_uni"],
                                 cur_epoch=g["cur_epoch"])
    st = out.state
    assert _close(out.loss, g["loss"]) and _close(st.base, g["base"]), g["name"]
    assert _close(st.ent, g["ent"]) and _close(st.uni, g["uni"]), g["name"]
    out.loss.backward()
    assert torch.allclose(lg.grad.cpu(), g["dlogits"], rtol=2e-5, atol=1e-8)
    assert int(st.nonfinite_logits) == int((~torch.isfinite(g["logits"])).sum())
    assert int(st.info) == 0
    # a non-unit upstream gradient just scales d loss / d logits
    lg2 = g["logits"].cuda().requires_grad_(True)
    (3.0 * losses.mort_train_loss(lg2, g["y"].cuda(), g["prim_acts"].cuda(), label_smoothing=g["label_smoothing"]).loss).backward()
    assert torch.allclose(lg2.grad.cpu(), 3.0 * g["dlogits"], rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize("idx", range(8))
def test_pheno_loss_matches_reference_golden(idx):
    from multimodalrouting_b200 import losses
    g = _load()["pheno"][idx]
    lg = g["logits"].cuda().requires_grad_(True)
    out = losses.pheno_train_loss(lg, g["y"].cuda(), g["rc_raw"].cuda(), g["prim_acts"].cuda(), _cu(g["route_mask"]),
                                  pos_weight=g["pos_weight"].cuda(), route_entropy_lambda=g["lam_ent"],
                                  route_entropy_warmup_epochs=g["warm_ent"], route_uniform_lambda=g["lam_uni"],
                                  route_uniform_warmup_epochs=g["warm_uni"], cur_epoch=g["cur_epoch"])
    st = out.state
    if g["raises"] is not None:
        exc = TypeError if g["raises"][0] == "TypeError" else AssertionError
        with pytest.raises(exc):
            st.check()
        with pytest.raises(exc):                                         # the drop-in (synchronising) entry point too
            rep, info = losses.coerce_rc_to_report(g["rc_raw"].cuda(), g["prim_acts"].cuda(), _cu(g["route_mask"]))
            losses.assert_routing_over_routes(rep)
        if exc is TypeError:
            return
    else:
        assert st.check() == g["info"]
    assert losses.INFO_TEXT[int(st.info)] == g["info"]
    assert out.rc_report.dtype == torch.float32
    assert torch.allclose(out.rc_report.cpu(), g["rc_report"], rtol=2e-6, atol=1e-9)
    assert _close(out.loss, g["loss"]), g["name"]
    out.loss.backward()
    assert torch.allclose(lg.grad.cpu(), g["dlogits"], rtol=2e-5, atol=1e-8)
    rep, info = (None, None)
    if g["raises"] is None:
        rep, info = losses.coerce_rc_to_report(g["rc_raw"].cuda(), g["prim_acts"].cuda(), _cu(g["route_mask"]))
        assert info == g["info"] and torch.equal(rep, out.rc_report)


def test_loss_full_size_vs_oracle_and_reproducible():
    """BASELINE configs[1] (Pheno, B=512, K=25, bf16 R as the bf16 path returns it) and configs[2] per-GPU size (Mort,
    B=4096) against the CPU oracle; two launches give bit-identical results (fixed reduction order)."""
    from multimodalrouting_b200 import losses
    gen = torch.Generator().manual_seed(77)
    B, K = 512, 25
    mask = (torch.rand(B, 10, generator=gen) < 0.8).float()
    mask[:, 0] = 1.0
    q = (torch.rand(B, 10, K, generator=gen) + 1e-3) * mask.unsqueeze(-1)
    for rc in (q / q.sum(1, keepdim=True), (q / q.sum(1, keepdim=True)).to(torch.bfloat16)):
        logits = torch.randn(B, K, generator=gen) * 3
        y = (torch.rand(B, K, generator=gen) < 0.2).float()
        pw = torch.rand(K, generator=gen) * 4 + 0.5
        pa = torch.rand(B, 10, generator=gen)
        ref = to.pheno_train_loss(logits.clone().requires_grad_(True), y, rc, pa, mask, pw, 0.01, 0, 0.1, 0, 1.0)
        outs = []
        for _ in range(2):
            lg = logits.cuda().requires_grad_(True)
            o = losses.pheno_train_loss(lg, y.cuda(), rc.cuda(), pa.cuda(), mask.cuda(), pos_weight=pw.cuda(),
                                        route_entropy_lambda=0.01, route_uniform_lambda=0.1)
            o.loss.backward()
            outs.append((o.loss.detach().clone(), o.rc_report.clone(), lg.grad.clone(), o.state.buf.clone()))
        assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
        assert _close(outs[0][0], ref["loss"]) and int(o.state.info) == ref["info"]
        assert _close(o.state.ent, ref["ent"]) and _close(o.state.uni, ref["uni"])
        assert torch.allclose(outs[0][1].cpu(), ref["rc_report"], rtol=2e-6, atol=1e-9)
        s = outs[0][1].sum(1)
        assert torch.allclose(s, torch.ones_like(s), atol=1e-5)          # rc_report is a distribution over routes
        assert float(o.state.max_route_sum_err) <= 1e-5
    B = 4096
    logits = torch.randn(B, 2, generator=gen) * 2
    y = (torch.rand(B, generator=gen) < 0.15).long()
    pa = torch.rand(B, 10, generator=gen)
    lgc = logits.clone().requires_grad_(True)
    ref = to.mort_train_loss(lgc, y, pa, 0.02, 0.01, 0, 0.1, 0, 1)
    (dref,) = torch.autograd.grad(ref["loss"], lgc)
    lg = logits.cuda().requires_grad_(True)
    o = losses.mort_train_loss(lg, y.cuda(), pa.cuda(), label_smoothing=0.02, route_entropy_lambda=0.01, route_uniform_lambda=0.1)
    o.loss.backward()
    assert _close(o.loss, ref["loss"]) and _close(o.state.ent, ref["ent"]) and _close(o.state.uni, ref["uni"])
    assert torch.allclose(lg.grad.cpu(), dref, rtol=2e-5, atol=1e-10)
    # the death logit only sees logits[:,1] - logits[:,0]: the two gradient columns are exact negatives
    assert torch.equal(lg.grad[:, 0], -lg.grad[:, 1])


def test_loss_is_graph_capturable():
    """No host sync anywhere in the loss tail: capture forward + backward of the Pheno loss, replay on new inputs."""
    from multimodalrouting_b200 import losses
    gen = torch.Generator().manual_seed(78)
    B, K = 64, 25
    q = torch.rand(B, 10, K, generator=gen) + 1e-3
    rc = (q / q.sum(1, keepdim=True)).cuda()
    y = (torch.rand(B, K, generator=gen) < 0.2).float().cuda()
    pa = torch.rand(B, 10, generator=gen).cuda()
    static_logits = torch.zeros(B, K, device="cuda")
    st = losses.LossState("cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):                                               # warm-up outside capture
            lg = static_logits.clone().requires_grad_(True)
            losses.pheno_train_loss(lg, y, rc, pa, None, route_entropy_lambda=0.01, state=st).loss.backward()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        lg = static_logits.clone().requires_grad_(True)
        out = losses.pheno_train_loss(lg, y, rc, pa, None, route_entropy_lambda=0.01, state=st)
        out.loss.backward()
        static_loss, static_grad = out.loss.detach(), lg.grad
    for seed in (1, 2):
        x = torch.randn(B, K, generator=torch.Generator().manual_seed(seed)) * 2
        static_logits.copy_(x.cuda())
        graph.replay()
        ref = to.pheno_train_loss(x.clone().requires_grad_(True), y.cpu(), rc.cpu(), pa.cpu(), None, None, 0.01, 0, 0.0, 0, 1.0)
        assert _close(static_loss, ref["loss"])
        xr = x.clone().requires_grad_(True)
        (dref,) = torch.autograd.grad(to.pheno_train_loss(xr, y.cpu(), rc.cpu(), pa.cpu(), None, None, 0.01)["loss"], xr)
        assert torch.allclose(static_grad.cpu(), dref, rtol=2e-5, atol=1e-9)
