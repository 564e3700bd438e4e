"""Producer epilogue of the route inputs: the reference's sanitising of the encoder
outputs right before the hot path, as one fused sm_100a row kernel (csrc/tail.cuh) instead of ~8 elementwise passes.

Same function names and semantics as the reference drivers:
    _clamp_norm, _safe_tensor, _sanitize_encoder_out     MIMIC-IV/MortModel/Paired_Cross_Attention/main.py:1772-1796
    _sanitize_encoder_out (nan_to_num only)              MIMIC-IV/PhenoModel/Paired_Cross_Attention/main.py:1452-1460

There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

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


# ---- route-input projections between the encoders and the hot path ---------------------------------------
#   BioClinicalBERT chunk projection  Sequential(LayerNorm(768), Linear(768 -> 256, bias=False))   encoders.py:289-293,472-475
#   CXR token projection              Linear(512 -> 256, bias=False)                               encoders.py:620,747-749
def _proj_dims(rows, d_in, d_out, has_ln, has_bias, x_dtype, dtype, engine):
    return _lib.ProjDims(int(rows), int(d_in), int(d_out), int(has_ln), int(has_bias), int(x_dtype), int(dtype), int(engine), 0)


def _proj_sizes(dims):
    s = [C.c_size_t() for _ in range(3)]
    _lib.check(_lib.load().mmr_producer_proj_sizes(C.byref(dims), *[C.byref(v) for v in s]), "mmr_producer_proj_sizes")
    return tuple(int(v.value) for v in s)


@torch.library.custom_op("mmr_b200::producer_proj_fwd", mutates_args=())
def producer_proj_fwd(x: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor], W: Tensor, bias: Optional[Tensor],
                      dtype: int, engine: int) -> Tuple[Tensor, Tensor]:
    """x [rows, d_in] fp32 | bf16 -> (y fp32 [rows, d_out], saved-for-backward bytes)."""
    _require_cuda(x, W)
    rows, d_in = x.shape
    dims = _proj_dims(rows, d_in, W.shape[0], ln_w is not None, bias is not None, _DTYPES[x.dtype], dtype, engine)
    saved_b, sf_b, _ = _proj_sizes(dims)
    y = torch.empty(rows, W.shape[0], dtype=torch.float32, device=x.device)
    saved = torch.empty(saved_b, dtype=torch.uint8, device=x.device)
    scratch = torch.empty(max(sf_b, 16), dtype=torch.uint8, device=x.device)
    rc = _lib.load().mmr_producer_proj_fwd(C.byref(dims), _ptr(x), _ptr(ln_w), _ptr(ln_b), _ptr(W), _ptr(bias), _ptr(y),
                                           _ptr(saved), _ptr(scratch), _stream())
    _lib.check(rc, "mmr_producer_proj_fwd")
    return y, saved


@producer_proj_fwd.register_fake
def _(x, ln_w, ln_b, W, bias, dtype, engine):
    dims = _proj_dims(x.shape[0], x.shape[1], W.shape[0], ln_w is not None, bias is not None, _DTYPES[x.dtype], dtype, engine)
    return x.new_empty(x.shape[0], W.shape[0], dtype=torch.float32), x.new_empty(_proj_sizes(dims)[0], dtype=torch.uint8)


@torch.library.custom_op("mmr_b200::producer_proj_bwd", mutates_args=())
def producer_proj_bwd(x: Tensor, ln_w: Optional[Tensor], W: Tensor, has_bias: bool, dy: Tensor, saved: Tensor,
                      need_dx: bool, dtype: int, engine: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Returns (dx fp32 [rows, d_in] | empty, d_ln_w, d_ln_b ([d_in] | empty), dW [d_out, d_in], dbias [d_out] | empty)."""
    rows, d_in = x.shape
    d_out = W.shape[0]
    dev = x.device
    dims = _proj_dims(rows, d_in, d_out, ln_w is not None, has_bias, _DTYPES[x.dtype], dtype, engine)
    _, _, sb_b = _proj_sizes(dims)
    scratch = torch.empty(sb_b, dtype=torch.uint8, device=dev)
    dx = torch.empty((rows, d_in) if need_dx else (0,), dtype=torch.float32, device=dev)
    dlw = torch.zeros(d_in if ln_w is not None else 0, dtype=torch.float32, device=dev)
    dlb = torch.zeros(d_in if ln_w is not None else 0, dtype=torch.float32, device=dev)
    dW = torch.zeros(d_out, d_in, dtype=torch.float32, device=dev)
    db = torch.zeros(d_out if has_bias else 0, dtype=torch.float32, device=dev)
    rc = _lib.load().mmr_producer_proj_bwd(C.byref(dims), _ptr(x), _ptr(ln_w), _ptr(W), _ptr(dy), _ptr(saved), _ptr(scratch),
                                           _ptr(dx) if need_dx else None, _ptr(dlw) if ln_w is not None else None,
                                           _ptr(dlb) if ln_w is not None else None, _ptr(dW),
                                           _ptr(db) if has_bias else None, _stream())
    _lib.check(rc, "mmr_producer_proj_bwd")
    return dx, dlw, dlb, dW, db


@producer_proj_bwd.register_fake
def _(x, ln_w, W, has_bias, dy, saved, need_dx, dtype, engine):
    d_in, d_out = x.shape[1], W.shape[0]
    f = lambda *sh: x.new_empty(*sh, dtype=torch.float32)
    return (f(x.shape[0], d_in) if need_dx else f(0), f(d_in if ln_w is not None else 0),
            f(d_in if ln_w is not None else 0), f(d_out, d_in), f(d_out if has_bias else 0))


class _ProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, W, bias, dtype, engine):
        x2 = x.detach()
        det = lambda t: None if t is None else t.detach().float().contiguous()
        y, saved = producer_proj_fwd(x2, det(ln_w), det(ln_b), det(W), det(bias), dtype, engine)
        ctx.save_for_backward(x2, saved, *[t for t in (ln_w, W) if t is not None])
        ctx.cfg = (ln_w is not None, bias is not None, dtype, engine)
        return y

    @staticmethod
    def backward(ctx, dy):
        has_ln, has_bias, dtype, engine = ctx.cfg
        sv = list(ctx.saved_tensors)
        x, saved = sv[0], sv[1]
        ln_w = sv[2].detach().float().contiguous() if has_ln else None
        W = sv[-1].detach().float().contiguous()
        dx, dlw, dlb, dW, db = producer_proj_bwd(x, ln_w, W, has_bias, dy.contiguous().float(), saved,
                                                 bool(ctx.needs_input_grad[0]), dtype, engine)
        return (dx.to(x.dtype) if ctx.needs_input_grad[0] else None, dlw if has_ln else None, dlb if has_ln else None,
                dW, db if has_bias else None, None, None)


def fused_ln_linear(x: Tensor, ln_weight: Optional[Tensor], ln_bias: Optional[Tensor], weight: Tensor,
                    bias: Optional[Tensor] = None) -> Tensor:
    """fp32 [.., d_out] = Linear(LayerNorm(x)) (LayerNorm optional) as one row kernel + one tensor-core GEMM.  bf16 operands
    with fp32 LayerNorm / accumulation under torch.autocast (the reference's mixed-precision flow), fp32 kernels otherwise."""
    from . import ops
    _require_cuda(x, weight)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    lead, d_in = x.shape[:-1], x.shape[-1]
    if weight.dim() != 2 or weight.shape[1] != d_in:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x.numel() // max(d_in, 1)}x{d_in} and "
                           f"{weight.shape[1]}x{weight.shape[0]})")
    x2 = x.reshape(-1, d_in).contiguous()
    if x2.shape[0] == 0:
        return x.new_zeros(*lead, weight.shape[0], dtype=torch.float32)
    y = _ProjFn.apply(x2, ln_weight, ln_bias, weight, bias, ops.resolve_dtype(), ops.resolve_engine())
    return y.reshape(*lead, weight.shape[0])


class NoteChunkProjector(torch.nn.Module):
    """`BioClinBERTEncoder.proj` (encoders.py:289-293): Sequential(LayerNorm(hidden), Linear(hidden, d, bias=False)) applied to
    the per-chunk [CLS] / masked-mean embeddings (encoders.py:472-475).  Same state_dict keys (proj.0.*, proj.1.weight)."""

    def __init__(self, hidden: int = 768, d: int = 256):
        super().__init__()
        self.hidden, self.out_dim = int(hidden), int(d)
        self.proj = torch.nn.Sequential(torch.nn.LayerNorm(self.hidden), torch.nn.Linear(self.hidden, self.out_dim, bias=False))

    def forward(self, chunk_emb: Tensor) -> Tensor:
        ln, lin = self.proj[0], self.proj[1]
        return fused_ln_linear(chunk_emb, ln.weight, ln.bias, lin.weight, None)


class ImageTokenProjector(torch.nn.Module):
    """`token_proj` of the CXR encoder (encoders.py:620, 747-749): Linear(layer4 channels -> d, bias=False) over the
    [B, H*W, C] feature-map tokens."""

    def __init__(self, token_in_dim: int = 512, d: int = 256):
        super().__init__()
        self.token_proj = torch.nn.Linear(int(token_in_dim), int(d), bias=False)

    def forward(self, tokens: Tensor) -> Tensor:
        return fused_ln_linear(tokens, None, None, self.token_proj.weight, None)


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
