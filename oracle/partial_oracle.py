"""CPU restatement of the attention-fusion route constructors of the Partial/ model variant (TEST INFRASTRUCTURE ONLY:
imported by tests/ and the golden generator, never by the product path).

    masked_mean                routing_and_heads.py:97-101
    CrossAttentionFusion       routing_and_heads.py:103-172   (nn.MultiheadAttention: torch/nn/functional.py
                               multi_head_attention_forward, PyTorch 2.11 -- packed in-projection, q scaled by head_dim^-1/2,
                               -inf key padding, softmax, out_proj)
    TriTokenAttentionFusion    routing_and_heads.py:175-206
of /root/reference/MIMIC-IV/PhenoModel/Partial/Cross_Attention/.  Functional form over a state_dict; pinned against the
unmodified reference modules by oracle/gen_golden_partial.py -> tests/golden/partial_fusion.pt (tests/test_partial_oracle.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

HEADS, HEAD_DIM = 8, 32


def masked_mean(x, m):
    m = m.to(x.dtype)
    return (x * m.unsqueeze(-1)).sum(dim=1) / m.sum(dim=1, keepdim=True).clamp_min(1.0)


def _mha(sd, pfx, q_in, kv_in, key_pad):
    """nn.MultiheadAttention(batch_first=True, need_weights=False) with key_padding_mask (True = pad)."""
    d = q_in.shape[-1]
    W, b = sd[pfx + "in_proj_weight"], sd[pfx + "in_proj_bias"]
    B, Tq, Tk = q_in.shape[0], q_in.shape[1], kv_in.shape[1]
    q = F.linear(q_in, W[:d], b[:d]).view(B, Tq, HEADS, HEAD_DIM).transpose(1, 2) * (HEAD_DIM ** -0.5)
    k = F.linear(kv_in, W[d:2 * d], b[d:2 * d]).view(B, Tk, HEADS, HEAD_DIM).transpose(1, 2)
    v = F.linear(kv_in, W[2 * d:], b[2 * d:]).view(B, Tk, HEADS, HEAD_DIM).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)     # a fully padded sample: zeros (the module discards it anyway)
    o = (p @ v).transpose(1, 2).reshape(B, Tq, d)
    return F.linear(o, sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"])


def cross_attention_fusion(sd, A, mA, B, mB, pool: str = "mean"):
    d = A.shape[-1]
    mA, mB = mA.to(A.dtype), mB.to(A.dtype)
    validB = (mB > 0.5).any(dim=1)
    A2B = _mha(sd, "attn.", A, B, mB < 0.5)
    X = F.layer_norm(A + A2B, (d,), sd["ln1.weight"], sd["ln1.bias"])
    ff = F.linear(F.relu(F.linear(X, sd["ff.0.weight"], sd["ff.0.bias"])), sd["ff.2.weight"], sd["ff.2.bias"])
    X = F.layer_norm(X + ff, (d,), sd["ln2.weight"], sd["ln2.bias"])
    if pool == "first":
        has_any = (mA > 0.5).any(dim=1)
        idx = torch.where(has_any, (mA > 0.5).to(A.dtype).argmax(dim=1), torch.zeros_like(has_any, dtype=torch.long))
        z = X[torch.arange(X.size(0)), idx]
    else:
        z = masked_mean(X, mA)
    z = z * validB.to(A.dtype).unsqueeze(-1)
    return F.linear(F.layer_norm(z, (d,), sd["out.0.weight"], sd["out.0.bias"]), sd["out.1.weight"], sd["out.1.bias"])


def tri_token_fusion(sd, L, mL, N, mN, I, mI):
    d = L.shape[-1]
    B = L.shape[0]
    kv = F.layer_norm(torch.cat([L, N, I], dim=1), (d,), sd["ln_kv.weight"], sd["ln_kv.bias"])
    m = torch.cat([mL, mN, mI], dim=1).to(L.dtype)
    valid = (m > 0.5).any(dim=1)
    out = _mha(sd, "attn.", sd["q"].expand(B, 1, d), kv, m < 0.5)
    z = out[:, 0, :] * valid.to(L.dtype).unsqueeze(-1)
    return F.linear(F.layer_norm(z, (d,), sd["out.0.weight"], sd["out.0.bias"]), sd["out.1.weight"], sd["out.1.bias"])
