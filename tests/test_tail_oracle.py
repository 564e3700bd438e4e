"""CPU: the tail oracle (oracle/tail_oracle.py) against fixtures produced by the reference's own definitions
(oracle/gen_golden_tail.py), and the host-side contracts of the producer-epilogue / training-tail bindings."""
import os

import pytest
import torch

from helpers import ROOT, max_rel
from oracle import tail_oracle as to

GOLD = os.path.join(ROOT, "tests", "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


@pytest.mark.parametrize("case", ["seq256", "seq768_bf16", "pool512_f16", "seq12"])
def test_sanitize_oracle_matches_reference(case):
    g = _load("tail_sanitize.pt")[case]
    x = g["x"].clone().requires_grad_(True)
    y = to.sanitize_mort(x)
    assert y.dtype == torch.float32 and g["mask_dtype"] == "torch.float32"
    assert torch.equal(y.detach(), g["y"])
    (dx,) = torch.autograd.grad(y, x, g["dy"])
    assert torch.equal(dx, g["dx"])
    assert torch.equal(to.sanitize_mort(g["x_bad"]), g["y_bad_mort"])
    assert torch.equal(to.sanitize_pheno(g["x_bad"]), g["y_bad_pheno"])
    # the norm clamp really is exercised: some rows were scaled, some were not
    n_in = g["x"].float().norm(dim=-1)
    assert (n_in > 20).any() and (n_in < 20).any()
    assert float(g["y"].norm(dim=-1).max()) <= 20.0 + 1e-3


def test_train_tail_oracle_matches_reference():
    g = _load("tail_adamw_ema.pt")
    out = to.train_tail(g["init"], g["grads"], g["lr"], g["betas"], g["eps"], g["weight_decay"], g["max_norm"],
                        g["ema_decay"])
    assert out["skipped"] == g["skipped"] and sum(g["skipped"]) == 1
    assert sum(1 for n in g["norms"] if n == n and n > g["max_norm"]) >= 2       # clipping active on some steps
    assert sum(1 for n in g["norms"] if n < g["max_norm"]) >= 2                   # and inactive on others
    for a, b in zip(out["norms"], g["norms"]):
        assert (a != a and b != b) or abs(a - b) <= 1e-6 * abs(b)
    for key in ("params", "exp_avg", "exp_avg_sq", "ema"):
        for a, b in zip(out[key], g[key]):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-9), key
    assert g["step"] == len(g["skipped"]) - 1


def test_tail_bindings_fail_loudly_on_cpu():
    from multimodalrouting_b200 import optim, producers
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        producers.sanitize_rows(torch.randn(2, 3, 256))
    p = torch.nn.Parameter(torch.randn(8))
    p.grad = torch.randn(8)
    opt = optim.FusedAdamW([p], lr=1e-3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step(max_norm=1.0)
    with pytest.raises(ValueError):
        optim.FusedAdamW([p], lr=-1.0)


def test_tail_entry_points_reject_bad_arguments_without_launch():
    import ctypes as C
    from multimodalrouting_b200 import _lib
    lib = _lib.load()
    assert lib.mmr_sanitize_rows_fwd(None, 0, None, 4, 256, 0, 20.0, None, None) != 0        # null pointers
    assert lib.mmr_sanitize_rows_fwd(1, 0, 1, 4, 250, 0, 20.0, None, None) != 0              # width not % 4
    assert lib.mmr_sanitize_rows_fwd(1, 0, 1, 4, 2048, 0, 20.0, None, None) != 0             # width > 1024
    assert lib.mmr_sanitize_rows_fwd(1, 9, 1, 4, 256, 0, 20.0, None, None) != 0              # dtype
    assert lib.mmr_sanitize_rows_fwd(1, 0, 1, 0, 256, 0, 20.0, None, None) == 0              # no rows: no launch
    hp = _lib.OptHyper(1e-3, 1.5, 0.999, 1e-8, 0.0, 0.0, 0.0, 0, 0)
    assert lib.mmr_opt_prepare(C.byref(hp), 1, None) != 0                                    # beta1 out of range
    assert C.sizeof(_lib.OptTensor) == 48 and C.sizeof(_lib.OptHyper) == 64
    assert lib.mmr_grad_sqnorm(None, 0, 1, None) == 0                                        # empty table
    t = (_lib.OptTensor * 1)()
    t[0].n = 5
    assert lib.mmr_grad_sqnorm(t, 1, 1, None) != 0                                           # null tensor pointers


def test_route_mask_oracle_matches_reference_and_synth():
    g = _load("tail_route_mask.pt")
    m = to.route_mask_from_presence(g["hasL"], g["hasN"], g["hasI"])
    assert torch.equal(m, g["mask"])
    assert 0 < float(g["mask"].mean()) < 1
    # the rule oracle/synth.py uses for BASELINE config 4 is the same one
    from oracle import synth
    for mod, idx in synth.NEEDS.items():
        need = [i for i, r in enumerate(to.ROUTES) if mod in r]
        assert sorted(idx) == need


# ------------------------------------------------------------------------------------- loss tail ---
def _close(a, b, tol=1e-6):
    a, b = float(torch.as_tensor(a).detach()), float(torch.as_tensor(b).detach())
    return abs(a - b) <= tol * max(1.0, abs(b))


@pytest.mark.parametrize("idx", range(4))
def test_mort_loss_oracle_matches_reference(idx):
    """Fixtures: the reference's own statements main.py:3092-3126 executed by oracle/gen_golden_tail.py."""
    g = _load("tail_loss.pt")["mort"][idx]
    lg = g["logits"].clone().requires_grad_(True)
    o = to.mort_train_loss(lg, g["y"], g["prim_acts"], g["label_smoothing"], g["lam_ent"], g["warm_ent"], g["lam_uni"],
                           g["warm_uni"], g["cur_epoch"])
    for k in ("loss", "base", "ent", "uni"):
        assert _close(o[k], g[k]), (g["name"], k)
    (dl,) = torch.autograd.grad(o["loss"], lg)
    assert torch.allclose(dl, g["dlogits"], rtol=1e-5, atol=1e-9)
    if g["name"] == "gated":
        assert float(g["ent"]) == 0.0 and float(g["uni"]) > 0.0        # `>=` gate: warm-up 3 off, warm-up 2 on at epoch 2
    if g["name"] == "nonfinite":
        bad = ~torch.isfinite(g["logits"])
        assert bad.sum() == 3 and bool((g["dlogits"][bad] == 0).all())   # nan_to_num blocks the gradient there


@pytest.mark.parametrize("idx", range(8))
def test_pheno_loss_oracle_matches_reference(idx):
    """Fixtures: coerce_rc_to_report / assert_routing_over_routes definitions and the statements main.py:2793-2812 of
    the reference executed as they are."""
    g = _load("tail_loss.pt")["pheno"][idx]
    args = (g["y"], g["rc_raw"], g["prim_acts"], g["route_mask"], g["pos_weight"], g["lam_ent"], g["warm_ent"],
            g["lam_uni"], g["warm_uni"], g["cur_epoch"])
    if g["raises"] and g["raises"][0] == "TypeError":
        with pytest.raises(TypeError, match="multiple values"):
            to.pheno_train_loss(g["logits"], *args)
        return
    lg = g["logits"].clone().requires_grad_(True)
    o = to.pheno_train_loss(lg, *args)
    assert to_info(o["info"]) == g["info"]
    assert torch.allclose(o["rc_report"], g["rc_report"], rtol=1e-6, atol=1e-9)
    assert _close(o["loss"], g["loss"])
    (dl,) = torch.autograd.grad(o["loss"], lg)
    assert torch.allclose(dl, g["dlogits"], rtol=1e-5, atol=1e-9)
    s = o["rc_report"].sum(dim=1)
    ok = bool(torch.allclose(s, torch.ones_like(s), atol=1e-3, rtol=0.0))
    assert ok == (g["raises"] is None)                                  # all_masked: the reference's assertion fires
    if g["name"] == "gated":
        assert _close(o["ent"], 0.0) and _close(o["uni"], 0.0)          # strict `>` gate at cur_epoch == 1.0


def to_info(code):
    from multimodalrouting_b200.losses import INFO_TEXT
    return INFO_TEXT[code]


def test_loss_bindings_fail_loudly_on_cpu_and_validate():
    import ctypes as C
    from multimodalrouting_b200 import _lib, losses
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.mort_train_loss(torch.randn(4, 2), torch.zeros(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.coerce_rc_to_report(torch.rand(2, 10, 3), None, None)
    lib = _lib.load()
    assert lib.mmr_loss_scratch_bytes(512) == 16 * 336 * 8 and lib.mmr_loss_scratch_bytes(1) == 336 * 8
    a = _lib.LossArgs()
    a.variant, a.B, a.K, a.atol = 0, 4, 3, 1e-3                          # Mort needs [B,2]
    assert lib.mmr_loss_fwd_bwd(C.byref(a), None) == 1
    assert b"2 logits" in lib.mmr_last_error_string()
    a.K = 2                                                              # null pointers: rejected before any launch
    assert lib.mmr_loss_fwd_bwd(C.byref(a), None) == 1
    assert b"null pointer" in lib.mmr_last_error_string()
    a.variant, a.K = 1, 33
    assert lib.mmr_loss_fwd_bwd(C.byref(a), None) == 1
    assert lib.mmr_loss_fwd_bwd(None, None) == 1
    # warm-up gates: Mort `>=`, Pheno `>`  (main.py:3116 vs PhenoModel main.py:2799)
    assert losses._gate(0.1, 2, 2, False) == 0.1 and losses._gate(0.1, 2, 2.0, True) == 0.0
    assert losses._gate(0.1, 0, 1, True) == 0.1 and losses._gate(0.0, 0, 5, False) == 0.0


def test_loss_argument_errors_match_the_reference_drivers():
    """Host-side checks of losses.py fire before any device work and raise what the reference's statements raise:
    torch's BCE shape error (ValueError, same text), the drivers' asserts on routing_coef (PhenoModel main.py:2761-2762),
    death_logit_from_logits2's assert (MortModel main.py:1754)."""
    from multimodalrouting_b200 import losses
    lg, y = torch.randn(4, 25), torch.zeros(4, 25)
    with pytest.raises(ValueError) as ours:
        losses.pheno_train_loss(lg, torch.zeros(4, 24))
    with pytest.raises(ValueError) as ref:
        torch.nn.BCEWithLogitsLoss()(lg, torch.zeros(4, 24))
    assert str(ours.value) == str(ref.value)
    with pytest.raises(AssertionError, match=r"routing_coef must be \[B,R,K\]"):
        losses.pheno_train_loss(lg, y, torch.rand(4, 250))
    with pytest.raises(AssertionError, match="Expected R=10"):
        losses.pheno_train_loss(lg, y, torch.rand(4, 9, 25))
    with pytest.raises(ValueError, match="pos_weight"):
        losses.pheno_train_loss(lg, y, pos_weight=torch.ones(3))
    with pytest.raises(ValueError, match="route_mask"):
        losses.pheno_train_loss(lg, y, torch.rand(4, 10, 25), None, torch.ones(4, 9))
    with pytest.raises(AssertionError, match=r"expected \[B,2\]"):
        losses.mort_train_loss(torch.randn(4, 3), torch.zeros(4))
    with pytest.raises(AssertionError, match=r"expected \[B,2\]"):
        losses.death_logit_from_logits2(torch.randn(4, 3))
    assert torch.equal(losses.death_logit_from_logits2(torch.tensor([[1.0, 3.0]])), torch.tensor([[2.0]]))
    with pytest.raises(AssertionError):
        losses.coerce_rc_to_report(torch.rand(4, 250), None, None)
    # well-formed CPU arguments pass validation and are refused by the device check
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.pheno_train_loss(lg, y, torch.rand(4, 10, 25), torch.rand(4, 10), torch.ones(10))


@pytest.mark.parametrize("name", ["chunk768", "chunk768_ragged", "token512"])
def test_projection_oracle_matches_reference_modules(name):
    """oracle chunk / token projections vs the outputs of the reference's own module definitions (tail_proj.pt)."""
    gold = torch.load(os.path.join(GOLD, "tail_proj.pt"), weights_only=False)
    c = gold[name]
    st = gold["state"][c["module"]]
    x = c["x"].clone().requires_grad_(True)
    if c["module"] == "chunk":
        y = to.chunk_projection(x, st["0.weight"], st["0.bias"], st["1.weight"])
    else:
        y = to.token_projection(x, st["weight"])
    assert max_rel(y, c["y"]) < 1e-6
    y.backward(c["dy"])
    assert max_rel(x.grad, c["dx"]) < 1e-6
