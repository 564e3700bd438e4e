"""Attention-fusion route constructors of the missing-modality ("Partial") model variant:

    CrossAttentionFusion, TriTokenAttentionFusion, build_fusions, make_route_inputs, masked_mean
    /root/reference/MIMIC-IV/PhenoModel/Partial/Cross_Attention/routing_and_heads.py:97-251

Same class names, constructor / forward signatures and state_dict keys as the reference (``attn.in_proj_weight``,
``attn.out_proj.*``, ``ln1``, ``ff.0``, ``ff.2``, ``ln2``, ``out.0``, ``out.1``, ``q``, ``ln_kv``), so
``new.load_state_dict(ref.state_dict())`` works.  ``self.attn`` is a real ``nn.MultiheadAttention`` kept as the parameter
container; its forward is never called.  What runs instead:

  * every Linear (Q / K|V in-projections with the head_dim^-1/2 query scaling folded into the weight, out-projection, the
    feed-forward pair, the LayerNorm + Linear output head, ``ln_kv`` + K|V projection of the tri-token block) on the
    tensor-core GEMM op of the producer projections (``producers.fused_ln_linear`` -> ``mmr_producer_proj_fwd/bwd``:
    tcgen05 under autocast, the fp32 engine otherwise), LayerNorm fused into the GEMM's row kernel where one precedes it;
  * the attention core on the hot path's per-patient attention kernels (``ops.AttentionFn`` -> ``mmr_attention_fwd/bwd``);
  * residual LayerNorms, ReLU and the masked pooling as PyTorch element-wise glue.

Key masking: the reference hands ``key_padding_mask`` to ``nn.MultiheadAttention`` (-inf fill) and zeroes the pooled vector of
a sample without any valid key; the attention kernels fill padded keys with finfo(bf16).min (a sample without valid keys
attends uniformly instead of dividing by zero) and the same zeroing follows, so both produce ``out(0)`` for such a sample and
no gradient flows into it.  Dropout inside the attention is not implemented (the reference's ``cross_attn_dropout`` is 0).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .producers import fused_ln_linear

HEAD_DIM = 32


def masked_mean(x: Tensor, m: Tensor) -> Tensor:
    """x [B,T,D], m [B,T] (1 = keep) -> [B,D]   (routing_and_heads.py:97-101)"""
    m = m.float()
    denom = m.sum(dim=1, keepdim=True).clamp_min(1.0)
    return (x * m.unsqueeze(-1)).sum(dim=1) / denom


def _check_attn(attn: nn.MultiheadAttention, d: int, training: bool) -> None:
    if d != 256 or attn.num_heads * HEAD_DIM != d:
        raise NotImplementedError("the B200 attention kernels are specialised for d = 256, 8 heads of 32")
    if attn.dropout > 0.0 and training:
        raise NotImplementedError("attention dropout > 0 is not implemented (the reference config uses 0.0)")


def _qkv(attn: nn.MultiheadAttention, q_in: Tensor, kv_in: Tensor, ln_kv: nn.LayerNorm = None):
    """(scaled Q [.., 256], K|V [.., 512]) of nn.MultiheadAttention's packed in-projection."""
    d = attn.embed_dim
    w, b = attn.in_proj_weight, attn.in_proj_bias
    s = float(HEAD_DIM) ** -0.5
    q = fused_ln_linear(q_in, None, None, w[:d] * s, b[:d] * s)
    if ln_kv is not None:
        kv = fused_ln_linear(kv_in, ln_kv.weight, ln_kv.bias, w[d:], b[d:])
    else:
        kv = fused_ln_linear(kv_in, None, None, w[d:], b[d:])
    return q, kv


class CrossAttentionFusion(nn.Module):
    """Directional cross-attention: A attends to B (Q = A, K / V = B), post-LN block, pooled over A -> [B, D]
    (routing_and_heads.py:103-172)."""

    def __init__(self, d: int, n_heads: int = 8, attn_dropout: float = 0.0, pool: str = "mean", ff_mult: int = 4):
        super().__init__()
        self.pool = pool
        self.attn = nn.MultiheadAttention(d, n_heads, dropout=attn_dropout, batch_first=True)
        self.ln1 = nn.LayerNorm(d)
        self.ff = nn.Sequential(nn.Linear(d, ff_mult * d), nn.ReLU(), nn.Linear(ff_mult * d, d))
        self.ln2 = nn.LayerNorm(d)
        self.out = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d))

    def forward(self, A: Tensor, mA: Tensor, B: Tensor, mB: Tensor) -> Tensor:
        d = self.attn.embed_dim
        _check_attn(self.attn, d, self.training)
        mA = mA.float()
        mB = mB.float()
        validB = (mB > 0.5).any(dim=1)
        q, kv = _qkv(self.attn, A, B)
        o = ops.AttentionFn.apply(q, kv, mB)
        A2B = fused_ln_linear(o, None, None, self.attn.out_proj.weight, self.attn.out_proj.bias)
        X = F.layer_norm(A.float() + A2B, (d,), self.ln1.weight, self.ln1.bias, self.ln1.eps)
        h = F.relu(fused_ln_linear(X, None, None, self.ff[0].weight, self.ff[0].bias))
        X = F.layer_norm(X + fused_ln_linear(h, None, None, self.ff[2].weight, self.ff[2].bias), (d,), self.ln2.weight,
                         self.ln2.bias, self.ln2.eps)
        if self.pool == "first":
            has_any = (mA > 0.5).any(dim=1)
            idx = torch.where(has_any, (mA > 0.5).float().argmax(dim=1), torch.zeros_like(has_any, dtype=torch.long))
            z = X[torch.arange(X.size(0), device=X.device), idx]
        else:
            z = masked_mean(X, mA)
        z = z * validB.float().unsqueeze(-1)      # a sample whose B stream holds no valid token contributes out(0)
        return fused_ln_linear(z, self.out[0].weight, self.out[0].bias, self.out[1].weight, self.out[1].bias)


class TriTokenAttentionFusion(nn.Module):
    """A learned query token attends to concat([L_seq, N_seq, I_seq]) -> [B, D]   (routing_and_heads.py:175-206)."""

    def __init__(self, d: int, n_heads: int = 8, attn_dropout: float = 0.0):
        super().__init__()
        self.q = nn.Parameter(torch.zeros(1, 1, d))
        nn.init.normal_(self.q, std=0.02)
        self.attn = nn.MultiheadAttention(d, n_heads, dropout=attn_dropout, batch_first=True)
        self.ln_kv = nn.LayerNorm(d)
        self.out = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d))

    def forward(self, L_seq, mL, N_seq, mN, I_seq, mI) -> Tensor:
        d = self.attn.embed_dim
        _check_attn(self.attn, d, self.training)
        B = L_seq.size(0)
        kv_in = torch.cat([L_seq, N_seq, I_seq], dim=1)
        m = torch.cat([mL, mN, mI], dim=1).float()
        validKV = (m > 0.5).any(dim=1)
        q1, kv = _qkv(self.attn, self.q.view(1, d), kv_in, self.ln_kv)      # LayerNorm(kv) fused into the K|V projection
        o = ops.AttentionFn.apply(q1.view(1, 1, d).expand(B, 1, d), kv, m)
        z = fused_ln_linear(o[:, 0, :], None, None, self.attn.out_proj.weight, self.attn.out_proj.bias)
        z = z * validKV.float().unsqueeze(-1)
        return fused_ln_linear(z, self.out[0].weight, self.out[0].bias, self.out[1].weight, self.out[1].bias)


def build_fusions(d: int, feature_mode: str = "seq", p_drop: float = 0.0, *, heads: int = 8, pool: str = "mean",
                  device=None) -> Dict[str, nn.Module]:
    """The seven fusion blocks keyed like the reference (routing_and_heads.py:209-231).  The reference reads heads / dropout /
    pooling from its global CFG; here they are keyword arguments with the reference's defaults."""
    pool = str(pool).lower().strip()
    if pool not in {"mean", "first"}:
        pool = "mean"
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    f = {k: CrossAttentionFusion(d, n_heads=heads, attn_dropout=p_drop, pool=pool).to(dev)
         for k in ("LN", "NL", "LI", "IL", "NI", "IN")}
    f["LNI"] = TriTokenAttentionFusion(d, n_heads=heads, attn_dropout=p_drop).to(dev)
    return f


def make_route_inputs(z, fusion):
    """The 10 route embeddings from the unimodal encoder outputs (routing_and_heads.py:237-251)."""
    Ls, Ns, Is = z["L"]["seq"], z["N"]["seq"], z["I"]["seq"]
    Lm, Nm, Im = z["L"]["mask"], z["N"]["mask"], z["I"]["mask"]
    return {
        "L": z["L"]["pool"], "N": z["N"]["pool"], "I": z["I"]["pool"],
        "LN": fusion["LN"](Ls, Lm, Ns, Nm), "NL": fusion["NL"](Ns, Nm, Ls, Lm),
        "LI": fusion["LI"](Ls, Lm, Is, Im), "IL": fusion["IL"](Is, Im, Ls, Lm),
        "NI": fusion["NI"](Ns, Nm, Is, Im), "IN": fusion["IN"](Is, Im, Ns, Nm),
        "LNI": fusion["LNI"](Ls, Lm, Ns, Nm, Is, Im),
    }
