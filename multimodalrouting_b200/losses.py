"""Loss tail of the hot path: what the drivers do with `logits, prim_acts, routing_coef`
right after `capsule_forward_from_encoded`, as one or two device-side launches (csrc/loss.cuh) with no host sync.

Reference semantics reproduced (same names where the reference has a function):

    death_logit_from_logits2            MortModel/Paired_Cross_Attention/main.py:1753-1755
    Mort training loss                  main.py:3084-3126   (_safe_tensor, label smoothing, BCEWithLogitsLoss,
                                                             route-entropy bonus, route-uniformity penalty on alpha)
    coerce_rc_to_report                 PhenoModel/Paired_Cross_Attention/main.py:1472-1564  (two .item() syncs there)
    assert_routing_over_routes          main.py:261-276
    Pheno training loss                 main.py:2755-2812   (BCEWithLogitsLoss(pos_weight), batch-mean routing
                                                             entropy / uniformity terms)

Only the BCE term carries gradient in the reference: `prim_acts` comes back detached (routing_and_heads.py:363) and
`coerce_rc_to_report` detaches the routing coefficients (main.py:1483), so the regularisers move the reported loss but
not the weights.  The kernels keep that: d loss / d logits is produced in the forward launch and the autograd node
just scales it.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import LOSS_STATE_BYTES, LossArgs
from .ops import _ptr, _require_cuda, _stream

MORT, PHENO = 0, 1
INFO_TEXT = {
    0: "no routing coefficients",
    1: "rc_raw is already p(route|phenotype) (nan-safe normalize)",
    2: "converted p(pheno|route)->p(route|pheno) (nan-safe + repair)",
    3: "WARNING: rc_raw not normalized; forced nan-safe normalize over routes",
}


class LossState:
    """Device-resident mmr_loss_state (48 bytes) exposed as 0-dim views; reading a view on the host is the only sync."""

    def __init__(self, device):
        self.buf = torch.zeros(LOSS_STATE_BYTES, dtype=torch.uint8, device=device)

    def _f(self, i) -> Tensor:
        return self.buf[4 * i:4 * i + 4].view(torch.float32)[0]

    def _i(self, i) -> Tensor:
        return self.buf[4 * i:4 * i + 4].view(torch.int32)[0]

    loss = property(lambda s: s._f(0))
    base = property(lambda s: s._f(1))
    ent = property(lambda s: s._f(2))
    uni = property(lambda s: s._f(3))
    err_routes = property(lambda s: s._f(4))
    err_k = property(lambda s: s._f(5))
    max_route_sum_err = property(lambda s: s._f(6))
    info = property(lambda s: s._i(7))
    nonfinite_logits = property(lambda s: s._i(8))

    def check(self, atol: float = 1e-3, name: str = "routing_coef") -> str:
        """The reference's host-side checks, on demand (this synchronises): returns coerce_rc_to_report's info string;
        raises what the reference raises -- TypeError for coefficients that sum to one over labels (main.py:1519 calls
        route_given_pheno with route_mask twice), AssertionError when rc_report is not normalised over routes."""
        info = int(self.info)
        if info == 2:
            raise TypeError("route_given_pheno() got multiple values for argument 'route_mask'")
        err = float(self.max_route_sum_err)
        if info != 0 and not err <= atol:
            raise AssertionError(f"{name}: NOT normalized over routes dim=1. abs_err max={err:.3e}")
        return INFO_TEXT[info]


def _f32(t: Optional[Tensor]) -> Optional[Tensor]:
    return None if t is None else t.detach().to(torch.float32).contiguous()


def _launch(variant: int, logits: Tensor, y: Tensor, *, pos_weight=None, prim_acts=None, rc_raw=None, route_mask=None,
            label_smoothing=0.0, lam_ent=0.0, lam_uni=0.0, atol=1e-3, want_grad=True, want_report=True,
            state: Optional[LossState] = None):
    # shape checks first (they mirror the reference's assertions / torch's BCE error and need no device), then the device check
    if logits.ndim != 2:
        raise ValueError(f"logits must be [B,K], got {tuple(logits.shape)}")
    B, K = logits.shape
    if variant == MORT:
        assert K == 2, f"expected [B,2], got {tuple(logits.shape)}"          # death_logit_from_logits2, main.py:1754
        if y.numel() != B:
            raise ValueError(f"y must hold one label per patient, got {tuple(y.shape)}")
    elif tuple(y.shape) != (B, K):
        raise ValueError(f"Target size ({y.size()}) must be the same as input size ({logits.size()})")   # torch's BCE text
    lg, yf = _f32(logits), _f32(y)
    pw, pa = _f32(pos_weight), _f32(prim_acts)
    if pw is not None and pw.numel() != K:
        raise ValueError(f"pos_weight must have {K} entries")
    if pa is not None and tuple(pa.shape) != (B, 10):
        raise ValueError(f"prim_acts must be [B,10], got {tuple(pa.shape)}")
    rc = None
    if rc_raw is not None:
        assert rc_raw.ndim == 3, f"routing_coef must be [B,R,K], got {tuple(rc_raw.shape)}"      # main.py:2761
        assert int(rc_raw.shape[1]) == 10, f"Expected R=10, got {tuple(rc_raw.shape)}"
        if tuple(rc_raw.shape) != (B, 10, K):
            raise ValueError(f"routing_coef must be [{B},10,{K}], got {tuple(rc_raw.shape)}")
        rc = rc_raw.detach()
        rc = rc.contiguous() if rc.dtype in (torch.float32, torch.bfloat16) else rc.float().contiguous()
    rm = None
    if route_mask is not None:
        rm = _f32(route_mask)
        if rm.ndim == 1:
            rm = rm.view(1, -1).expand(B, -1).contiguous()
        if tuple(rm.shape) != (B, 10):
            raise ValueError(f"route_mask must be [B,10] or [10], got {tuple(route_mask.shape)}")
    _require_cuda(logits, y, pos_weight, prim_acts, rc_raw, route_mask)
    dev = logits.device
    st = state if state is not None else LossState(dev)
    lib = _lib.load()
    scratch = torch.empty(lib.mmr_loss_scratch_bytes(B) // 8, dtype=torch.float64, device=dev)
    dlog = torch.empty(B, K, dtype=torch.float32, device=dev) if want_grad else None
    rep = torch.empty(B, 10, K, dtype=torch.float32, device=dev) if (rc is not None and want_report) else None
    a = LossArgs()
    a.variant, a.B, a.K = variant, B, K
    a.rc_dtype = 1 if (rc is not None and rc.dtype == torch.bfloat16) else 0
    a.logits, a.y, a.pos_weight, a.prim_acts = _ptr(lg), _ptr(yf), _ptr(pw), _ptr(pa)
    a.rc_raw, a.route_mask = _ptr(rc), _ptr(rm)
    a.label_smoothing, a.route_entropy_lambda, a.route_uniform_lambda, a.atol = label_smoothing, lam_ent, lam_uni, atol
    a.dlogits, a.rc_report, a.state, a.scratch = _ptr(dlog), _ptr(rep), st.buf.data_ptr(), scratch.data_ptr()
    _lib.check(lib.mmr_loss_fwd_bwd(C.byref(a), _stream()), "mmr_loss_fwd_bwd")
    return st, dlog, rep


class _LossFn(torch.autograd.Function):
    """loss (0-dim) as a differentiable function of logits; everything else rides along as constants."""

    @staticmethod
    def forward(ctx, logits, variant, y, kw):
        st, dlog, rep = _launch(variant, logits, y, **kw)
        ctx.save_for_backward(dlog)
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(*([rep] if rep is not None else []))
        loss = st.loss.clone()
        return (loss, rep) if rep is not None else (loss,)

    @staticmethod
    def backward(ctx, g, *_):
        (dlog,) = ctx.saved_tensors
        return (dlog * g).to(ctx.in_dtype), None, None, None


class LossParts(NamedTuple):
    loss: Tensor            # 0-dim, differentiable w.r.t. logits
    state: LossState        # base / ent / uni / diagnostics, device resident
    rc_report: Optional[Tensor]


def _gate(lam: float, warmup: float, cur_epoch: float, strict: bool) -> float:
    """`lambda > 0 and (warmup <= 0 or cur_epoch >= warmup)`; Pheno uses `>` (main.py:2799 vs Mort main.py:3116)."""
    if lam <= 0.0:
        return 0.0
    if warmup <= 0:
        return float(lam)
    return float(lam) if (cur_epoch > warmup if strict else cur_epoch >= warmup) else 0.0


def death_logit_from_logits2(logits2: Tensor) -> Tensor:
    """main.py:1753-1755 (a view op; kept in torch)."""
    assert logits2.ndim == 2 and logits2.size(1) == 2, f"expected [B,2], got {tuple(logits2.shape)}"
    return (logits2[:, 1] - logits2[:, 0]).unsqueeze(1)


def mort_train_loss(logits: Tensor, y: Tensor, prim_acts: Optional[Tensor] = None, *, label_smoothing: float = 0.02,
                    route_entropy_lambda: float = 0.0, route_entropy_warmup_epochs: int = 0,
                    route_uniform_lambda: float = 0.0, route_uniform_warmup_epochs: int = 0, cur_epoch: int = 1,
                    state: Optional[LossState] = None) -> LossParts:
    """loss = BCEWithLogits(logits[:,1]-logits[:,0], smooth(y)) - ent_bonus + uniform_pen    (main.py:3092-3126)."""
    st = state if state is not None else LossState(logits.device)
    kw = dict(prim_acts=prim_acts, label_smoothing=float(label_smoothing),
              lam_ent=_gate(route_entropy_lambda, route_entropy_warmup_epochs, cur_epoch, False),
              lam_uni=_gate(route_uniform_lambda, route_uniform_warmup_epochs, cur_epoch, False), state=st)
    (loss,) = _LossFn.apply(logits, MORT, y, kw)
    return LossParts(loss, st, None)


def pheno_train_loss(logits: Tensor, y: Tensor, routing_coef: Optional[Tensor] = None, prim_acts: Optional[Tensor] = None,
                     route_mask: Optional[Tensor] = None, *, pos_weight: Optional[Tensor] = None,
                     route_entropy_lambda: float = 0.0, route_entropy_warmup_epochs: int = 0,
                     route_uniform_lambda: float = 0.0, route_uniform_warmup_epochs: int = 0, cur_epoch: float = 1.0,
                     atol: float = 1e-3, state: Optional[LossState] = None) -> LossParts:
    """rc_report = coerce_rc_to_report(routing_coef, prim_acts, route_mask); loss = BCEWithLogits(logits, y, pos_weight)
    - lambda_e * H(mean_b rc) + lambda_u * U(mean_b rc)                                          (main.py:2755-2812).
    `state.check()` performs the reference's host-side assertions when the caller wants them."""
    st = state if state is not None else LossState(logits.device)
    kw = dict(pos_weight=pos_weight, prim_acts=prim_acts, rc_raw=routing_coef, route_mask=route_mask, atol=float(atol),
              lam_ent=_gate(route_entropy_lambda, route_entropy_warmup_epochs, cur_epoch, True),
              lam_uni=_gate(route_uniform_lambda, route_uniform_warmup_epochs, cur_epoch, True), state=st)
    out = _LossFn.apply(logits, PHENO, y, kw)
    return LossParts(out[0], st, out[1] if len(out) > 1 else None)


def coerce_rc_to_report(rc_raw: Tensor, prim_acts: Optional[Tensor], route_mask: Optional[Tensor], split_name: str = "",
                        atol: float = 1e-3) -> Tuple[Tensor, str]:
    """Drop-in for the drivers' coerce_rc_to_report (main.py:1472-1564): (rc_report [B,R,K] fp32, info string).
    Returning the string synchronises, as the reference's `.item()` calls do; training code that must not sync uses
    pheno_train_loss / `coerce_rc_to_report_async` and reads `state.info` later."""
    rep, st = coerce_rc_to_report_async(rc_raw, prim_acts, route_mask, atol=atol)
    info = int(st.info)
    if info == 2:
        raise TypeError("route_given_pheno() got multiple values for argument 'route_mask'")
    return rep, INFO_TEXT[info]


def coerce_rc_to_report_async(rc_raw: Tensor, prim_acts: Optional[Tensor], route_mask: Optional[Tensor], *,
                              atol: float = 1e-3, state: Optional[LossState] = None) -> Tuple[Tensor, LossState]:
    assert rc_raw.ndim == 3, f"rc_raw must be [B,R,K], got {tuple(rc_raw.shape)}"
    B, _, K = rc_raw.shape
    zeros = torch.zeros(B, K, dtype=torch.float32, device=rc_raw.device)
    st, _, rep = _launch(PHENO, zeros, zeros, prim_acts=None, rc_raw=rc_raw, route_mask=route_mask, atol=float(atol),
                         want_grad=False, state=state)
    return rep, st


def assert_routing_over_routes(rc: Tensor, routes_dim: int = 1, atol: float = 1e-3, name: str = "routing_coef"):
    """main.py:261-276 (host-side check, kept in torch: it is a debugging assertion, not part of the step)."""
    if rc.ndim != 3:
        raise AssertionError(f"{name}: expected [B,R,K], got {tuple(rc.shape)}")
    s = rc.sum(dim=routes_dim)
    if not torch.allclose(s, torch.ones_like(s), atol=atol, rtol=0.0):
        max_err = float((s - 1.0).abs().max())
        raise AssertionError(f"{name}: NOT normalized over routes dim={routes_dim}. abs_err max={max_err:.3e}")
