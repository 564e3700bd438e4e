"""Seeded synthetic weights and MIMIC-IV-shaped inputs: the workload definition shared by bench.py, the tools and the tests
(`oracle/synth.py` re-exports it).  Pure data generation -- no arithmetic of the path lives here.

Weights are keyed by the reference's state_dict names and are
generated from a CPU ``torch.Generator`` so the same tensors can be rebuilt on the
GPU box without the reference present.  Unlike the reference's default init
(zero biases, unit LayerNorm, zero decision embedding -- SURVEY.md section 0.6) every
tensor is perturbed, otherwise bias/affine/gradient paths would be tested
vacuously.  Inputs follow SURVEY.md section 8d.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

ROUTES = ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]
CROSS = ["trans_l_with_n", "trans_l_with_i", "trans_n_with_l", "trans_n_with_i",
         "trans_i_with_l", "trans_i_with_n"]


def mult_param_spec(orig_d_l=256, orig_d_n=256, orig_d_i=256, d=256, layers=4) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) in MULTModel registration order (M/mult_model.py:30-57)."""
    spec = [("proj_l.weight", (d, orig_d_l, 1), "conv"),
            ("proj_n.weight", (d, orig_d_n, 1), "conv"),
            ("proj_i.weight", (d, orig_d_i, 1), "conv")]
    for m in "lni":
        spec += [(f"trans_{m}.layer_norm.weight", (d,), "ln_w"), (f"trans_{m}.layer_norm.bias", (d,), "ln_b")]
    for enc in CROSS:
        for l in range(layers):
            p = f"{enc}.layers.{l}."
            spec += [(p + "self_attn.in_proj_weight", (3 * d, d), "xavier"),
                     (p + "self_attn.in_proj_bias", (3 * d,), "bias"),
                     (p + "self_attn.out_proj.weight", (d, d), "xavier"),
                     (p + "self_attn.out_proj.bias", (d,), "bias"),
                     (p + "fc1.weight", (4 * d, d), "xavier"), (p + "fc1.bias", (4 * d,), "bias"),
                     (p + "fc2.weight", (d, 4 * d), "xavier"), (p + "fc2.bias", (d,), "bias"),
                     (p + "layer_norms.0.weight", (d,), "ln_w"), (p + "layer_norms.0.bias", (d,), "ln_b"),
                     (p + "layer_norms.1.weight", (d,), "ln_w"), (p + "layer_norms.1.bias", (d,), "ln_b")]
        spec += [(f"{enc}.layer_norm.weight", (d,), "ln_w"), (f"{enc}.layer_norm.bias", (d,), "ln_b")]
    for n in ("ln", "li", "ni"):
        spec += [(f"proj_pair_{n}.weight", (d, 2 * d), "linear"), (f"proj_pair_{n}.bias", (d,), "bias")]
    spec += [("final_lni.weight", (d, 3 * d), "linear"), ("final_lni.bias", (d,), "bias")]
    return spec


def _fill(shape, kind, g) -> torch.Tensor:
    if kind == "xavier":
        a = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=g) * 2 - 1) * a
    if kind in ("linear", "conv"):
        fan_in = shape[1]
        a = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * a
    if kind == "bias":
        return 0.02 * torch.randn(shape, generator=g)
    if kind == "ln_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if kind == "ln_b":
        return 0.1 * torch.randn(shape, generator=g)
    raise ValueError(kind)


def make_state(*, K: int, orig_d_n: int = 256, seed: int = 42, sharp: float = 1.0, d: int = 256,
               layers: int = 4, pc_dim: int = 32, mc_dim: int = 64, orig_d_l: int = 0, orig_d_i: int = 0):
    """Returns (sd_mult, sd_proj, sd_head) on CPU fp32.  orig_d_l / orig_d_i default to d (no input projection)."""
    g = torch.Generator().manual_seed(seed)
    sd_mult = {n: _fill(s, k, g) for n, s, k in mult_param_spec(orig_d_l or d, orig_d_n, orig_d_i or d, d, layers)}
    sd_proj = {}
    for r in ROUTES:
        sd_proj[f"proj.{r}.weight"] = sharp * _fill((pc_dim + 1, d), "linear", g)
        sd_proj[f"proj.{r}.bias"] = sharp * _fill((pc_dim + 1,), "bias", g)
    R = len(ROUTES)
    sd_head = {
        "embedding": 0.1 * torch.randn(K, mc_dim, generator=g),
        "bias": 0.1 * torch.randn(K, generator=g),
        "capsule.w": sharp * math.sqrt(K / (pc_dim * R)) * torch.randn(R, pc_dim, K, mc_dim, generator=g),
        "capsule.beta_u": torch.randn(K, generator=g),
        "capsule.beta_a": torch.randn(K, generator=g),
        "pose_to_mc.weight": _fill((mc_dim, pc_dim), "linear", g),
    }
    return sd_mult, sd_proj, sd_head


def _clamp_norm(x, max_norm=20.0):
    n = x.norm(dim=-1, keepdim=True) + 1e-6       # M/main.py:1772-1779
    return x * torch.clamp(max_norm / n, max=1.0)


# indices of ROUTES that need each modality (PX/main.py:109-132)
NEEDS = {"L": [0, 3, 4, 5, 6, 9], "N": [1, 3, 4, 7, 8, 9], "I": [2, 5, 6, 7, 8, 9]}


def make_inputs(*, B: int, TL: int = 48, TN: int = 16, TI: int = 49, d_l: int = 256, d_n: int = 256,
                d_i: int = 256, K: int = 2, seed: int = 42, missing: bool = False,
                p_label: float = 0.2) -> Dict[str, torch.Tensor]:
    """SURVEY.md section 8d synthetic batch.  ``missing`` adds per-patient N/I dropout (config 4)."""
    g = torch.Generator().manual_seed(seed)
    x_l = _clamp_norm(torch.randn(B, TL, d_l, generator=g))
    x_n = _clamp_norm(torch.randn(B, TN, d_n, generator=g))
    x_i = _clamp_norm(torch.randn(B, TI, d_i, generator=g))
    len_l = torch.randint(max(1, TL // 2), TL + 1, (B,), generator=g)
    len_n = torch.randint(1, TN + 1, (B,), generator=g)
    mL = (torch.arange(TL).unsqueeze(0) < len_l.unsqueeze(1)).float()
    mN = (torch.arange(TN).unsqueeze(0) < len_n.unsqueeze(1)).float()
    mI = torch.ones(B, TI)
    route_mask = torch.ones(B, len(ROUTES))
    if missing:
        has_n = torch.rand(B, generator=g) < 0.7
        has_i = torch.rand(B, generator=g) < 0.7
        mN = mN * has_n.float().unsqueeze(1)
        # dropped image: half the time zero tokens with mI=1 (reference keeps I_mask=1), else mI=0
        zero_tok = torch.rand(B, generator=g) < 0.5
        drop_i = ~has_i
        x_i = torch.where((drop_i & zero_tok).view(B, 1, 1), torch.zeros_like(x_i), x_i)
        mI = mI * (~(drop_i & ~zero_tok)).float().unsqueeze(1)
        for b in range(B):
            if not has_n[b]:
                route_mask[b, NEEDS["N"]] = 0.0
            if not has_i[b]:
                route_mask[b, NEEDS["I"]] = 0.0
    y = (torch.rand(B, K, generator=g) < p_label).float()
    return {"x_l": x_l, "x_n": x_n, "x_i": x_i, "mL": mL, "mN": mN, "mI": mI,
            "route_mask": route_mask, "y": y}


def loss_fn(logits: torch.Tensor, y: torch.Tensor, variant: str) -> torch.Tensor:
    """Mort: BCE on death_logit = logits[:,1]-logits[:,0] (M/main.py:1753-1755,3104-3107);
    Pheno: BCEWithLogits over [B,K] (P/main.py:2467,2793)."""
    import torch.nn.functional as F
    if variant == "mort":
        dl = (logits[:, 1] - logits[:, 0]).unsqueeze(1)
        return F.binary_cross_entropy_with_logits(dl.float(), y[:, :1].float())
    return F.binary_cross_entropy_with_logits(logits.float(), y.float())


# ---- attention-fusion modules of the Partial/ variant (partial_fusion.py): seeded parameters -----------------------
def make_fusion_state(kind: str, seed: int, d: int = 256, ff_mult: int = 4):
    """state_dict (CPU fp32) for CrossAttentionFusion (kind = "cross") or TriTokenAttentionFusion (kind = "tri") with the
    reference's key names; weights ~ N(0, 1/fan_in), non-trivial biases and LayerNorm affines."""
    g = torch.Generator().manual_seed(seed)

    def w(o, i):
        return torch.randn(o, i, generator=g) / (i ** 0.5)

    def b(n, s=0.1):
        return s * torch.randn(n, generator=g)

    def ln(pfx, sd):
        sd[pfx + ".weight"] = 1.0 + 0.2 * torch.randn(d, generator=g)
        sd[pfx + ".bias"] = b(d)

    sd = {}
    if kind == "tri":
        sd["q"] = 0.5 * torch.randn(1, 1, d, generator=g)
    sd["attn.in_proj_weight"] = w(3 * d, d)
    sd["attn.in_proj_bias"] = b(3 * d)
    sd["attn.out_proj.weight"] = w(d, d)
    sd["attn.out_proj.bias"] = b(d)
    if kind == "cross":
        ln("ln1", sd)
        sd["ff.0.weight"] = w(ff_mult * d, d); sd["ff.0.bias"] = b(ff_mult * d)
        sd["ff.2.weight"] = w(d, ff_mult * d); sd["ff.2.bias"] = b(d)
        ln("ln2", sd)
    else:
        ln("ln_kv", sd)
    ln("out.0", sd)
    sd["out.1.weight"] = w(d, d)
    sd["out.1.bias"] = b(d)
    return sd


def make_fusion_inputs(B: int, TL: int, TN: int, TI: int, seed: int, d: int = 256):
    """Three encoder sequences with padding masks: ragged valid lengths, one sample without any N token, one without I."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, T in (("L", TL), ("N", TN), ("I", TI)):
        out[name] = torch.randn(B, T, d, generator=g)
        n = torch.randint(1, T + 1, (B,), generator=g)
        out["m" + name] = (torch.arange(T)[None, :] < n[:, None]).float()
    if B > 1:
        out["mN"][1] = 0.0
    if B > 2:
        out["mI"][2] = 0.0
    return out

