"""CPU restatement of the steps either side of the hot path (TEST INFRASTRUCTURE -- only tests/, smoke() and
bench.py's cpu_baseline may import this; the product path never does).

Pinned: tests/golden/tail_*.pt are produced by oracle/gen_golden_tail.py, which executes the reference's OWN
function/class definitions (extracted from MortModel/PhenoModel main.py, which cannot be imported as a module
because it needs matplotlib) together with torch.nn.utils.clip_grad_norm_ and torch.optim.AdamW;
tests/test_tail_oracle.py checks this restatement against them.

Third-party arithmetic: torch.optim.AdamW and torch.nn.utils.clip_grad_norm_ (PyTorch 2.11.0, the reference pins
"PyTorch >= 2.1").  Their published algorithm (Loshchilov & Hutter, decoupled weight decay; torch/optim/adamw.py
`_single_tensor_adam` with decoupled_weight_decay=True) is restated in `adamw_step` below.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch


# ---- producer epilogue --------------------------------------------------------------------------------
def clamp_norm(x: torch.Tensor, max_norm: float = 20.0) -> torch.Tensor:
    """MortModel/Paired_Cross_Attention/main.py:1772-1779 (_clamp_norm)."""
    if x.ndim in (2, 3):
        n = x.norm(dim=x.ndim - 1, keepdim=True) + 1e-6
        return x * torch.clamp(max_norm / n, max=1.0)
    return x


def safe_tensor(x: torch.Tensor) -> torch.Tensor:
    """main.py:1781-1786 (_safe_tensor): rewrite only when a non-finite entry exists."""
    if not torch.isfinite(x).all():
        x = torch.nan_to_num(x, nan=0.0, posinf=1e4, neginf=-1e4)
    return x


def sanitize_mort(x: torch.Tensor) -> torch.Tensor:
    """main.py:1788-1796 (_sanitize_encoder_out, "seq"/"pool" entries)."""
    return safe_tensor(clamp_norm(x.float(), 20.0)).float()


def sanitize_pheno(x: torch.Tensor) -> torch.Tensor:
    """PhenoModel/Paired_Cross_Attention/main.py:1452-1460."""
    return torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)


def chunk_projection(x: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """`BioClinBERTEncoder.proj` = Sequential(LayerNorm(hidden), Linear(hidden, d, bias=False)),
    MortModel/Paired_Cross_Attention/encoders.py:289-293, applied at :472-475."""
    return torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (x.shape[-1],), ln_w, ln_b), W)


def token_projection(x: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """`token_proj` = Linear(token_in_dim, d, bias=False), encoders.py:620, applied at :747-749."""
    return torch.nn.functional.linear(x, W)


# ---- training tail ------------------------------------------------------------------------------------
def clip_grad_norm(grads: List[torch.Tensor], max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_ (norm_type 2): total = ||(||g_i||)_i||, coef = clamp(max_norm/(total+1e-6), max=1),
    every gradient scaled in place (main.py:3147,3156).  Returns the total norm."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return float(total)


def grads_are_finite(grads: List[torch.Tensor]) -> bool:
    """main.py:46-57."""
    return all(bool(torch.isfinite(g).all()) for g in grads)


def adamw_step(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float) -> None:
    """torch.optim.AdamW single-tensor algorithm (amsgrad=False), `step` already incremented (main.py:2886-2890, 3161)."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        p.mul_(1.0 - lr * weight_decay)
        m.lerp_(g, 1.0 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def ema_update(shadow: List[torch.Tensor], params: List[torch.Tensor], decay: float) -> None:
    """EMA.update (main.py:69-90)."""
    for s, p in zip(shadow, params):
        s.mul_(decay).add_(p, alpha=1.0 - decay)


def train_tail(params, grads_per_step, lr, betas, eps, weight_decay, max_norm, ema_decay):
    """The reference's per-step sequence (main.py:3143-3165): clip, finite guard (skip), AdamW, EMA.
    Returns dict(params, exp_avg, exp_avg_sq, ema, norms, skipped)."""
    params = [p.clone() for p in params]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    ema = [p.clone() for p in params]
    norms, skipped, step = [], [], 0
    for grads in grads_per_step:
        grads = [g.clone() for g in grads]
        norms.append(clip_grad_norm(grads, max_norm))
        if not grads_are_finite(grads):
            skipped.append(True)
            continue
        skipped.append(False)
        step += 1
        adamw_step(params, grads, m, v, step, lr, betas[0], betas[1], eps, weight_decay)
        ema_update(ema, params, ema_decay)
    return dict(params=params, exp_avg=m, exp_avg_sq=v, ema=ema, norms=norms, skipped=skipped)


# ---- route mask from presence -------------------------------------------------------------------------
ROUTES = ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]


def route_mask_from_presence(hasL, hasN, hasI):
    """PhenoModel/Partial/Cross_Attention/routing_and_heads.py:10-64: a route is allowed iff every modality in its
    name is present."""
    has = {"L": hasL.float(), "N": hasN.float(), "I": hasI.float()}
    cols = []
    for r in ROUTES:
        m = torch.ones_like(has["L"])
        for ch in r:
            m = m * has[ch]
        cols.append(m)
    return torch.stack(cols, dim=1)


# ---- loss tail ----------------------------------------------------------------------------------------
# Third-party arithmetic: torch.nn.functional.binary_cross_entropy_with_logits (ATen; the reference builds
# nn.BCEWithLogitsLoss objects, MortModel main.py:2717, PhenoModel main.py:2467).  Its published formula,
# loss = (1 - y) x - (1 + (pos_weight - 1) y) log(sigmoid(x)), mean-reduced, is what `bce_with_logits` restates.
def bce_with_logits(x: torch.Tensor, y: torch.Tensor, pos_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    lw = 1.0 if pos_weight is None else (pos_weight - 1.0) * y + 1.0
    return ((1.0 - y) * x - lw * torch.nn.functional.logsigmoid(x)).mean()


def mort_train_loss(logits, y, prim_acts, label_smoothing=0.02, route_entropy_lambda=0.0,
                    route_entropy_warmup_epochs=0, route_uniform_lambda=0.0, route_uniform_warmup_epochs=0, cur_epoch=1):
    """MortModel/Paired_Cross_Attention/main.py:3084-3126.  Returns dict(loss, base, ent, uni)."""
    logits = safe_tensor(logits.float())                                   # :3084
    prim_acts = safe_tensor(prim_acts.float())                             # :3085
    y_f = y.float().view(-1, 1)                                            # :3092
    death_logit = (logits[:, 1] - logits[:, 0]).unsqueeze(1)               # :1753-1755
    if label_smoothing > 0.0:
        y_f = y_f * (1.0 - label_smoothing) + 0.5 * label_smoothing        # :3104-3105
    base = bce_with_logits(death_logit, y_f)                               # :3107
    ent = base.new_zeros(())
    uni = base.new_zeros(())
    pa = prim_acts.clamp_min(1e-6)
    pa = pa / pa.sum(dim=1, keepdim=True).clamp_min(1e-6)                  # :3113-3114
    if route_entropy_lambda > 0.0 and (route_entropy_warmup_epochs <= 0 or cur_epoch >= route_entropy_warmup_epochs):
        p = pa.clamp_min(1e-12)
        ent = -(p * p.log()).sum(dim=1).mean() * route_entropy_lambda       # :3116-3119
    if route_uniform_lambda > 0.0 and (route_uniform_warmup_epochs <= 0 or cur_epoch >= route_uniform_warmup_epochs):
        p_mean = pa.mean(dim=0)
        uni = ((p_mean - 1.0 / p_mean.numel()).pow(2)).sum() * route_uniform_lambda   # :3120-3123
    return dict(loss=base - ent + uni, base=base, ent=ent, uni=uni)


def coerce_rc_to_report(rc_raw, prim_acts, route_mask, atol=1e-3):
    """PhenoModel/Paired_Cross_Attention/main.py:1472-1564.  Returns (rc_report, info code 1 | 3); info 2 (the
    coefficients sum to one over labels) raises TypeError exactly as the reference's call at :1519 does
    (`route_given_pheno(rc_raw_f, pa_f, route_mask=rm_f)` binds route_mask twice)."""
    rc = torch.nan_to_num(rc_raw.detach().float(), nan=0.0, posinf=0.0, neginf=0.0)
    rm = None if route_mask is None else torch.nan_to_num(route_mask.detach().float(), nan=0.0, posinf=0.0, neginf=0.0)
    err_routes = float((rc.sum(dim=1) - 1.0).abs().max())
    err_k = float((rc.sum(dim=2) - 1.0).abs().max())
    if err_routes < atol:
        info = 1
    elif err_k < atol:
        raise TypeError("route_given_pheno() got multiple values for argument 'route_mask'")
    else:
        info = 3
        rc = rc.clamp(min=0.0)
    if rm is not None:
        rc = rc * rm.unsqueeze(-1)
    denom = rc.sum(dim=1, keepdim=True)
    bad = (~torch.isfinite(denom)) | (denom < 1e-8)
    if bad.any():
        if rm is not None:
            avail = rm.unsqueeze(-1)
            uniform = avail / avail.sum(dim=1, keepdim=True).clamp(min=1.0)
            rc = torch.where(bad, uniform.expand_as(rc), rc)
        else:
            rc = torch.where(bad, torch.full_like(rc, 1.0 / float(rc.shape[1])), rc)
    rc = rc / rc.sum(dim=1, keepdim=True).clamp(min=1e-8)
    return rc, info


def pheno_train_loss(logits, y, routing_coef, prim_acts, route_mask, pos_weight=None, route_entropy_lambda=0.0,
                     route_entropy_warmup_epochs=0, route_uniform_lambda=0.0, route_uniform_warmup_epochs=0,
                     cur_epoch=1.0, atol=1e-3):
    """PhenoModel/Paired_Cross_Attention/main.py:2755-2812.  Returns dict(loss, base, ent, uni, rc_report, info)."""
    rc_report, info = (None, 0)
    if routing_coef is not None:
        rc_report, info = coerce_rc_to_report(routing_coef, prim_acts, route_mask, atol)        # :2765
    logits = safe_tensor(logits.float())                                                        # :2784
    base = bce_with_logits(logits, y.float(), pos_weight)                                       # :2793
    ent = base.new_zeros(())
    uni = base.new_zeros(())
    if rc_report is not None:
        rc = rc_report.float().clamp(1e-6, 1.0)                                                 # :2797
        R = rc.shape[1]
        if route_entropy_lambda > 0.0 and (route_entropy_warmup_epochs <= 0 or cur_epoch > route_entropy_warmup_epochs):
            m = rc.mean(dim=0)
            ent = route_entropy_lambda * (-(m * m.log()).sum(dim=0)).mean()                     # :2800-2804
        if route_uniform_lambda > 0.0 and (route_uniform_warmup_epochs <= 0 or cur_epoch > route_uniform_warmup_epochs):
            m = rc.mean(dim=0)
            uni = route_uniform_lambda * (((m.transpose(0, 1) - 1.0 / R) ** 2).sum(dim=1)).mean()   # :2806-2812
    return dict(loss=base - ent + uni, base=base, ent=ent, uni=uni, rc_report=rc_report, info=info)
