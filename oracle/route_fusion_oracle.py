"""CPU/PyTorch restatement of the reference route-fusion + capsule-routing path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``multimodalrouting_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the timed CPU baseline -- never as the product path.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the unmodified reference
modules from ``/root/reference`` in the build container, runs them on seeded
weights/inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
(fp32, CPU) on every run.  (The reference itself ships no tests or golden
vectors)

The restatement is functional and batch-first ([B,T,C]); the reference is a tree
of nn.Modules working in [T,B,C].  Parameters are addressed by the reference's
own state_dict key names so either implementation's weights can be fed in.  The
same torch op families are used (F.linear / bmm / einsum / layer_norm / fp32
softmax) so that running it under ``torch.autocast(bfloat16)`` reproduces the
reference's mixed-precision dtype flow ("autocast dtype flow").

Reference files followed (relative to /root/reference/MIMIC-IV):
  M = MortModel/Paired_Cross_Attention, P = PhenoModel/Paired_Cross_Attention
  M/mult_model.py:116-193      -> mult_forward
  M/transformer.py:56-115      -> encoder_forward
  M/transformer.py:149-216     -> _encoder_layer
  P/multihead_attention.py:48-148 -> _attention
  P/position_embedding.py:68-117  -> positional_table
  M/capsule_layers.py:75-117   -> _capsule_fc
  M|P/routing_and_heads.py:101-121,194-272,271-369 -> projector_forward,
      capsule_head_forward, routing_forward
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

ROUTES = ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]  # M/env_config.py:53
# (query modality, key/value modality, reference encoder attribute) M/mult_model.py:38-45,151-158
DIRECTIONS = [
    ("LN", "l", "n", "trans_l_with_n"),
    ("LI", "l", "i", "trans_l_with_i"),
    ("NL", "n", "l", "trans_n_with_l"),
    ("NI", "n", "i", "trans_n_with_i"),
    ("IL", "i", "l", "trans_i_with_l"),
    ("IN", "i", "n", "trans_i_with_n"),
]
PRIOR_FLOOR = 0.02    # M/env_config.py:157 (default always wins)
PRIOR_CEILING = 0.98  # M/env_config.py:158


def positional_table(T: int, dim: int, device=None) -> torch.Tensor:
    """Rows 1..T of the sinusoid table, truncated through int64 (quirk).

    P/position_embedding.py:68-93 builds [sin | cos] with frequencies
    exp(-i*log(1e4)/(half-1)); the caller feeds a LongTensor so the table is
    cast ``.to(int64)`` (P/position_embedding.py:111, M/transformer.py:69-71),
    i.e. truncated toward zero, before being added.  Positions are 1..T
    (padding_idx=0, P/position_embedding.py:12-41).
    """
    half = dim // 2
    step = math.log(10000.0) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * (-step))
    pos = torch.arange(T + 1, dtype=torch.float32).unsqueeze(1)
    ang = pos * freq.unsqueeze(0)
    tab = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if dim % 2 == 1:
        tab = torch.cat([tab, torch.zeros(T + 1, 1)], dim=1)
    tab[0, :] = 0.0
    tab = tab.to(torch.int64)           # the quirk: truncation toward zero
    out = tab[1:T + 1]
    return out.to(device) if device is not None else out


def _embed(x_btc: torch.Tensor, dim: int) -> torch.Tensor:
    """embed_scale * x + pos  (M/transformer.py:63-72)."""
    x = math.sqrt(dim) * x_btc
    pos = positional_table(x_btc.shape[1], dim, x_btc.device).to(dtype=x.dtype)
    return x + pos.unsqueeze(0)


def _attention(sd, pfx, q_in, k_in, v_in, key_pad, heads):
    """P/multihead_attention.py:48-148 in batch-first form.  q_in [B,Tq,C]; k_in/v_in [B,Tk,C]."""
    B, Tq, C = q_in.shape
    Tk = k_in.shape[1]
    hd = C // heads
    w = sd[pfx + "in_proj_weight"]
    b = sd[pfx + "in_proj_bias"]
    q = F.linear(q_in, w[:C], b[:C])
    k = F.linear(k_in, w[C:2 * C], b[C:2 * C])
    v = F.linear(v_in, w[2 * C:], b[2 * C:])
    q = q * (hd ** -0.5)
    q = q.reshape(B, Tq, heads, hd).permute(0, 2, 1, 3).reshape(B * heads, Tq, hd)
    k = k.reshape(B, Tk, heads, hd).permute(0, 2, 1, 3).reshape(B * heads, Tk, hd)
    v = v.reshape(B, Tk, heads, hd).permute(0, 2, 1, 3).reshape(B * heads, Tk, hd)
    s = torch.bmm(q, k.transpose(1, 2))                      # [B*H,Tq,Tk]
    if key_pad is not None:
        s = s.view(B, heads, Tq, Tk).masked_fill(
            key_pad.view(B, 1, 1, Tk), torch.finfo(s.dtype).min).view(B * heads, Tq, Tk)
    sf = s if s.dtype == torch.float64 else s.float()        # fp32 softmax (fp64 when evaluating the exact answer)
    p = F.softmax(sf, dim=-1).to(dtype=s.dtype)              # cast back
    o = torch.bmm(p, v)                                      # [B*H,Tq,hd]
    o = o.view(B, heads, Tq, hd).permute(0, 2, 1, 3).reshape(B, Tq, C)
    return F.linear(o, sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"])


def _encoder_layer(sd, pfx, x, xk, xv, q_keep, key_pad, heads):
    """M/transformer.py:149-216 (pre-LN, normalize_before=True)."""
    C = x.shape[-1]
    ln0 = (sd[pfx + "layer_norms.0.weight"], sd[pfx + "layer_norms.0.bias"])
    ln1 = (sd[pfx + "layer_norms.1.weight"], sd[pfx + "layer_norms.1.bias"])
    res = x
    h = F.layer_norm(x, (C,), *ln0)
    if q_keep is not None:
        h = h * q_keep
    if xk is None:
        a = _attention(sd, pfx + "self_attn.", h, h, h, key_pad, heads)
    else:
        kk = F.layer_norm(xk, (C,), *ln0)
        vv = F.layer_norm(xv, (C,), *ln0)
        a = _attention(sd, pfx + "self_attn.", h, kk, vv, key_pad, heads)
    x = res + a
    if q_keep is not None:
        x = x * q_keep
    res = x
    h = F.layer_norm(x, (C,), *ln1)
    if q_keep is not None:
        h = h * q_keep
    h = F.relu(F.linear(h, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"]))
    h = F.linear(h, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"])
    x = res + h
    if q_keep is not None:
        x = x * q_keep
    return x


def encoder_forward(sd, pfx, n_layers, x_q, x_kv, q_mask, kv_mask, heads):
    """M/transformer.py:56-115.  Inputs/outputs [B,T,C]; masks [B,T] float (1 = keep) or None."""
    C = x_q.shape[-1]
    x = _embed(x_q, C)
    q_keep = None
    if q_mask is not None:
        q_keep = q_mask.to(dtype=x.dtype).unsqueeze(-1)
        x = x * q_keep
    if x_kv is not None:
        xk = _embed(x_kv, C)
        xv = _embed(x_kv, C)
    else:
        xk = xv = None
    for l in range(n_layers):
        if xk is not None:
            kp = (kv_mask < 0.5) if kv_mask is not None else None
            x = _encoder_layer(sd, f"{pfx}layers.{l}.", x, xk, xv, q_keep, kp, heads)
        else:
            kp = (q_mask < 0.5) if q_mask is not None else None
            x = _encoder_layer(sd, f"{pfx}layers.{l}.", x, None, None, q_keep, kp, heads)
    x = F.layer_norm(x, (C,), sd[pfx + "layer_norm.weight"], sd[pfx + "layer_norm.bias"])
    if q_keep is not None:
        x = x * q_keep
    return x


def _masked_mean(h_btc, m_bt):
    """M/mult_model.py:84-90."""
    if m_bt is None:
        return h_btc.mean(dim=1)
    m = m_bt.float()
    denom = m.sum(dim=1, keepdim=True).clamp_min(1.0)
    return (h_btc * m.unsqueeze(-1)).sum(dim=1) / denom


def _float_mask(m, B):
    if m is None:
        return None
    if m.dim() == 1:
        m = m.unsqueeze(0).expand(B, -1)
    return m.float()


def mult_forward(sd: Dict[str, torch.Tensor], x_l, x_n, x_i, mL=None, mN=None, mI=None,
                 *, d: int = 256, heads: int = 8, layers: int = 4, self_layers: int = 0,
                 out_dtype: Optional[torch.dtype] = None) -> Dict[str, torch.Tensor]:
    """M/mult_model.py:116-193.  ``sd`` uses MULTModel's state_dict names."""
    B = x_l.shape[0]
    assert x_n.shape[0] == B and x_i.shape[0] == B
    mL, mN, mI = _float_mask(mL, B), _float_mask(mN, B), _float_mask(mI, B)

    def proj(x, name):
        if x.shape[-1] == d:
            return x
        # Conv1d(k=1, bias=False) over [B,C,T]  (M/mult_model.py:30-32,134-136)
        return F.conv1d(x.transpose(1, 2), sd[name + ".weight"]).transpose(1, 2)

    p = {"l": proj(x_l, "proj_l"), "n": proj(x_n, "proj_n"), "i": proj(x_i, "proj_i")}
    m = {"l": mL, "n": mN, "i": mI}

    out = {}
    for r, mod in (("L", "l"), ("N", "n"), ("I", "i")):
        h = encoder_forward(sd, f"trans_{mod}.", self_layers, p[mod], None, m[mod], m[mod], heads)
        out[r] = _masked_mean(h, m[mod])
    for r, qm, km, attr in DIRECTIONS:
        h = encoder_forward(sd, attr + ".", layers, p[qm], p[km], m[qm], m[km], heads)
        out[r] = _masked_mean(h, m[qm])

    def lin(name, x):
        return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])

    e_ln = lin("proj_pair_ln", torch.cat([out["LN"], out["NL"]], dim=1))
    e_li = lin("proj_pair_li", torch.cat([out["LI"], out["IL"]], dim=1))
    e_ni = lin("proj_pair_ni", torch.cat([out["NI"], out["IN"]], dim=1))
    out["LNI"] = lin("final_lni", torch.cat([e_ln, e_li, e_ni], dim=1))
    tgt = out_dtype if out_dtype is not None else sd["final_lni.weight"].dtype
    return {r: out[r].to(dtype=tgt) for r in ROUTES}


def projector_forward(sd_proj, route_embs: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """M/routing_and_heads.py:111-121.  Returns poses [B,10,pc], acts [B,10,1]."""
    pdt = sd_proj["proj.L.weight"].dtype
    pcs = [F.linear(route_embs[r].to(pdt), sd_proj[f"proj.{r}.weight"], sd_proj[f"proj.{r}.bias"])
           for r in ROUTES]
    pc = torch.stack(pcs, dim=1)
    pcd = pc.shape[-1] - 1
    return pc[:, :, :pcd], torch.sigmoid(pc[:, :, pcd:])


def _capsule_fc(w, x, act, v_prev, next_act):
    """M/capsule_layers.py:75-117 (act_type EM/ONES, p_drop=0, no nonlinearity)."""
    B, N, _A = x.shape
    M, D = w.shape[2], w.shape[3]
    act = act.reshape(B, N)
    if v_prev is None:
        q = F.softmax(torch.zeros(B, N, M).type_as(x), dim=2)
        v = torch.einsum("bnm,bna,namd->bmd", q, x, w)          # NB: no act factor at it=0
    else:
        s = torch.einsum("bna,namd,bmd->bnm", x, w, v_prev)
        s = s * (1.0 / (D ** 0.5))
        q = F.softmax(s, dim=2)
        q = torch.einsum("bnm,bm->bnm", q, next_act)
        q = q / (torch.sum(q, dim=2, keepdim=True) + 1e-10)
        v = torch.einsum("bnm,bna,namd,bn->bmd", q, x, w, act)
    return v, q


def capsule_head_forward(sd_head, prim_pose, prim_act, route_mask=None, *, variant: str,
                         num_routing: int = 3):
    """CapsuleMortalityHead.forward.  variant 'mort': M/routing_and_heads.py:194-268;
    variant 'pheno': P/routing_and_heads.py:194-272."""
    assert variant in ("mort", "pheno")
    if prim_act.dim() == 2:
        prim_act = prim_act.unsqueeze(-1)
    elif not (prim_act.dim() == 3 and prim_act.size(-1) == 1):
        raise ValueError(f"prim_act must be [B,R] or [B,R,1], got {tuple(prim_act.shape)}")
    B = prim_pose.shape[0]
    act_route = torch.ones_like(prim_act) if variant == "mort" else None
    if route_mask is not None:
        rm = route_mask
        if rm.ndim == 1:
            rm = rm.view(1, -1, 1).expand(B, -1, 1)
        elif rm.ndim == 2:
            rm = rm.unsqueeze(-1)
        else:
            raise ValueError(f"route_mask must be [R] or [B,R], got {tuple(rm.shape)}")
        rm = rm.to(device=prim_pose.device, dtype=prim_pose.dtype)
        prim_pose = prim_pose * rm
        prim_act = prim_act * rm
        if variant == "mort":
            act_route = act_route * rm
    if variant == "pheno":
        act_route = prim_act
    w = sd_head["capsule.w"]
    v, q = None, None
    for _ in range(num_routing):
        na = None
        if v is not None:
            na = torch.ones(B, v.shape[1], device=v.device, dtype=prim_act.dtype)
        v, q = _capsule_fc(w, prim_pose, act_route, v, na)
    alpha = prim_act.squeeze(-1)
    # route_given_pheno, M/routing_and_heads.py:39-48
    resp = q
    if route_mask is not None:
        mm = route_mask
        mm = mm.view(1, -1, 1) if mm.ndim == 1 else mm.unsqueeze(-1)
        resp = resp * mm.to(device=resp.device, dtype=resp.dtype)
    R = resp / resp.sum(dim=1, keepdim=True).clamp_min(1e-10)
    if variant == "mort":
        d_bkp = torch.einsum("brk,brp->bkp", R, prim_pose)
    else:
        d_bkp = torch.einsum("brk,br,brp->bkp", R, alpha, prim_pose)
    d_bkm = F.linear(d_bkp, sd_head["pose_to_mc.weight"])
    logits = torch.einsum("bkm,km->bk", d_bkm, sd_head["embedding"]) + sd_head["bias"]
    return logits, alpha, R


def routing_forward(sd_proj, sd_head, route_embs: Dict[str, torch.Tensor], *, variant: str,
                    route_mask=None, act_temperature: float = 1.0, detach_priors: bool = False,
                    acts_override=None, num_routing: int = 3):
    """forward_capsule_from_route_dict, M/routing_and_heads.py:271-369 (P: 276-375).
    Returns (logits [B,K], alpha [B,10] detached, R [B,10,K])."""
    pdt = sd_proj["proj.L.weight"].dtype
    embs = {r: route_embs[r].to(pdt) for r in ROUTES}
    poses, acts = projector_forward(sd_proj, embs)
    prior = acts if acts_override is None else acts_override.to(device=acts.device, dtype=acts.dtype)
    keep = None
    if route_mask is not None:
        rm = route_mask
        if rm.ndim == 1:
            rm = rm.view(1, -1).expand(prior.size(0), -1)
        elif rm.ndim != 2:
            raise ValueError(f"route_mask must be [R] or [B,R], got {tuple(rm.shape)}")
        rm = rm.to(device=prior.device, dtype=prior.dtype)
        keep = rm.unsqueeze(-1).bool()
        prior = prior * rm.unsqueeze(-1)
    if act_temperature != 1.0 and keep is not None:
        x32 = torch.clamp(prior[keep].to(torch.float32), 1e-6, 1.0 - 1e-6)
        y32 = torch.sigmoid((torch.log(x32) - torch.log1p(-x32)) / float(act_temperature))
        prior = prior.clone()
        prior[keep] = y32.to(dtype=prior.dtype)
    if keep is None:
        prior = torch.clamp(prior, min=PRIOR_FLOOR, max=PRIOR_CEILING)
    else:
        prior = prior.clone()
        prior[keep] = torch.clamp(prior[keep], min=PRIOR_FLOOR, max=PRIOR_CEILING)
    acts_for_caps = prior.detach() if detach_priors else prior
    logits, alpha, R = capsule_head_forward(sd_head, poses, acts_for_caps.squeeze(-1), route_mask,
                                            variant=variant, num_routing=num_routing)
    return logits, alpha.detach(), R


def full_forward(sd_mult, sd_proj, sd_head, x_l, x_n, x_i, mL, mN, mI, *, variant: str,
                 route_mask=None, act_temperature: float = 1.0, detach_priors: bool = False,
                 heads: int = 8, layers: int = 4, acts_override=None, num_routing: int = 3):
    """forward_capsule_from_multmodel, M/routing_and_heads.py:372-409 (adapter = Identity)."""
    routes = mult_forward(sd_mult, x_l, x_n, x_i, mL, mN, mI, heads=heads, layers=layers)
    logits, alpha, R = routing_forward(sd_proj, sd_head, routes, variant=variant,
                                       route_mask=route_mask, act_temperature=act_temperature,
                                       detach_priors=detach_priors, acts_override=acts_override,
                                       num_routing=num_routing)
    return logits, alpha, routes, R
