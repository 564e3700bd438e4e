"""Producer epilogue of the route inputs: the reference's sanitising of the encoder
outputs right before the hot path, as one fused sm_100a row kernel (csrc/tail.cuh) instead of ~8 elementwise passes.

Same function names and semantics as the reference drivers:
    _clamp_norm, _safe_tensor, _sanitize_encoder_out     MIMIC-IV/MortModel/Paired_Cross_Attention/main.py:1772-1796
    _sanitize_encoder_out (nan_to_num only)              MIMIC-IV/PhenoModel/Paired_Cross_Attention/main.py:1452-1460

There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib
from .ops import _ptr, _require_cuda, _stream

MODE_CLAMP_NORM, MODE_NAN_TO_NUM = 0, 1
_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _rows(x: Tensor):
    if x.dim() < 1:
        raise ValueError("sanitize expects at least one dimension")
    D = x.shape[-1]
    return x.numel() // max(D, 1), D


@torch.library.custom_op("mmr_b200::sanitize_rows_fwd", mutates_args=())
def sanitize_rows_fwd(x: Tensor, mode: int, max_norm: float) -> Tensor:
    _require_cuda(x)
    if x.dtype not in _DTYPES:
        raise ValueError(f"sanitize: unsupported dtype {x.dtype}")
    x = x.contiguous()
    rows, D = _rows(x)
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    rc = _lib.load().mmr_sanitize_rows_fwd(_ptr(x), _DTYPES[x.dtype], _ptr(y), rows, D, mode, max_norm, None, _stream())
    _lib.check(rc, "mmr_sanitize_rows_fwd")
    return y


@sanitize_rows_fwd.register_fake
def _(x, mode, max_norm):
    return x.new_empty(x.shape, dtype=torch.float32)


@torch.library.custom_op("mmr_b200::sanitize_rows_bwd", mutates_args=())
def sanitize_rows_bwd(x: Tensor, dy: Tensor, mode: int, max_norm: float) -> Tensor:
    _require_cuda(x, dy)
    x = x.contiguous()
    dy = dy.contiguous().float()
    rows, D = _rows(x)
    dx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    rc = _lib.load().mmr_sanitize_rows_bwd(_ptr(x), _DTYPES[x.dtype], _ptr(dy), _ptr(dx), rows, D, mode, max_norm, _stream())
    _lib.check(rc, "mmr_sanitize_rows_bwd")
    return dx


@sanitize_rows_bwd.register_fake
def _(x, dy, mode, max_norm):
    return x.new_empty(x.shape, dtype=torch.float32)


class _SanitizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode, max_norm):
        ctx.save_for_backward(x)
        ctx.cfg = (mode, max_norm)
        return sanitize_rows_fwd(x.detach(), mode, max_norm)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        mode, max_norm = ctx.cfg
        dx = sanitize_rows_bwd(x.detach(), dy, mode, max_norm)
        return dx.to(x.dtype), None, None


def sanitize_rows(x: Tensor, mode: int = MODE_CLAMP_NORM, max_norm: float = 20.0) -> Tensor:
    """fp32 [.., D] = nan_to_num(clamp_norm(x.float())) (mode 0) / nan_to_num(x, 0, 0, 0) (mode 1), differentiable."""
    _require_cuda(x)
    return _SanitizeFn.apply(x, int(mode), float(max_norm))


def count_nonfinite(x: Tensor, mode: int = MODE_CLAMP_NORM, max_norm: float = 20.0):
    """(sanitised tensor, device int64 counter of replaced entries) -- the number the reference's
    '[NaN/Inf GUARD]' message prints (main.py:1781-1786) without a host sync."""
    _require_cuda(x)
    x = x.contiguous()
    rows, D = _rows(x)
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=x.device)
    rc = _lib.load().mmr_sanitize_rows_fwd(_ptr(x), _DTYPES[x.dtype], _ptr(y), rows, D, mode, max_norm, _ptr(cnt), _stream())
    _lib.check(rc, "mmr_sanitize_rows_fwd")
    return y, cnt


# ---- reference-named surface (MortModel/Paired_Cross_Attention/main.py:1772-1796) -------------------------
def _clamp_norm(x: Tensor, max_norm: float = 20.0) -> Tensor:
    """x * clamp(max_norm / (||x|| + 1e-6), max=1) over the last dim for [B,D] / [B,T,D]; other ranks pass through."""
    if x.ndim in (2, 3):
        return sanitize_rows(x, MODE_CLAMP_NORM, max_norm)
    return x


def _safe_tensor(x: Tensor, name: str = "") -> Tensor:
    """nan_to_num(x, nan=0, posinf=1e4, neginf=-1e4); finite tensors come back unchanged (as the reference, which
    only rewrites when a non-finite entry exists)."""
    return torch.nan_to_num(x, nan=0.0, posinf=1e4, neginf=-1e4)


def _sanitize_encoder_out(out: Dict[str, Optional[Tensor]], name: str, variant: str = "mort") -> Dict[str, Optional[Tensor]]:
    """Drop-in for the drivers' `_sanitize_encoder_out`: "seq" [B,T,D] and "pool" [B,D] are clamped to token norm
    <= 20 and made finite in ONE kernel each (variant="mort"), or only nan_to_num'ed (variant="pheno");
    "mask" becomes float."""
    mode = MODE_CLAMP_NORM if variant == "mort" else MODE_NAN_TO_NUM
    out2 = dict(out)
    for key in ("seq", "pool"):
        t = out2.get(key)
        if t is not None:
            y = sanitize_rows(t, mode, 20.0)
            out2[key] = y if variant == "mort" else y.to(t.dtype)
    if out2.get("mask") is not None:
        out2["mask"] = out2["mask"].float()
    return out2


# ---- route mask of the missing-modality protocol (PhenoModel/Partial/Cross_Attention) -----------------------
def build_route_mask_from_presence(hasL: Tensor, hasN: Tensor, hasI: Tensor, *, device: Optional[torch.device] = None,
                                   dtype: torch.dtype = torch.float32, drop_routes=()) -> Tensor:
    """[B, 10] route mask in ROUTES order, 1 = allowed: a route needs every modality in its name
    (Partial/Cross_Attention/routing_and_heads.py:10-64).  `drop_routes`: route indices zeroed for the whole batch
    (the training-time route dropout of MortModel/Paired_Cross_Attention/main.py:3027-3033).  One launch."""
    if device is None:
        device = hasL.device
    hs = [h.to(device=device).float().contiguous() for h in (hasL, hasN, hasI)]
    _require_cuda(*hs)
    B = hs[0].shape[0]
    if any(h.dim() != 1 or h.shape[0] != B for h in hs):
        raise ValueError("hasL, hasN, hasI must be [B] vectors of equal length")
    bits = 0
    for r in drop_routes:
        if not 0 <= int(r) < 10:
            raise ValueError(f"drop_routes: route index {r} out of range")
        bits |= 1 << int(r)
    out = torch.empty(B, 10, dtype=torch.float32, device=device)
    rc = _lib.load().mmr_route_mask_from_presence(_ptr(hs[0]), _ptr(hs[1]), _ptr(hs[2]), B, bits, _ptr(out), _stream())
    _lib.check(rc, "mmr_route_mask_from_presence")
    return out if dtype == torch.float32 else out.to(dtype)


def build_route_mask_from_modalities(hasL: Tensor, hasN: Tensor, hasI: Tensor) -> Tensor:
    """Partial/Cross_Attention/main.py:109-132 -- same rule, float32 on the inputs' device."""
    return build_route_mask_from_presence(hasL, hasN, hasI)
