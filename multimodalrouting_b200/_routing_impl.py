"""RoutePrimaryProjector / RouteDimAdapter / CapsuleMortalityHead / forward_capsule_* for the B200 path.

Mirrors MIMIC-IV/{MortModel,PhenoModel}/Paired_Cross_Attention/routing_and_heads.py: same names,
argument meaning, return structure, error conditions and state_dict keys.  The Mort and Pheno files
of the reference differ only inside CapsuleMortalityHead.forward (routing activation and the final
aggregation, Mort :208-265 vs Pheno :208-265); here that is the ``variant`` of the head, fixed by
which sub-package is imported (multimodalrouting_b200.MortModel / .PhenoModel).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .capsule_layers import CapsuleFC
from .env_config import CFG, ROUTES
from .mult_model import MULTModel


def route_given_pheno(q_brk: torch.Tensor, route_mask=None, eps=1e-10):
    """routing_and_heads.py:39-48 (tiny post-processing helper; plain tensor ops by design)."""
    resp = q_brk
    if route_mask is not None:
        m = route_mask
        if m.ndim == 1:
            m = m.view(1, -1, 1)
        elif m.ndim == 2:
            m = m.unsqueeze(-1)
        resp = resp * m.to(device=resp.device, dtype=resp.dtype)
    denom = resp.sum(dim=1, keepdim=True).clamp_min(eps)
    return resp / denom


def make_route_inputs_mult(z, multmodel: MULTModel):
    """routing_and_heads.py:82-98."""
    Ls, Ns, Is = z["L"]["seq"], z["N"]["seq"], z["I"]["seq"]
    Lm, Nm, Im = z["L"].get("mask", None), z["N"].get("mask", None), z["I"].get("mask", None)
    routes = multmodel(x_l=Ls, x_n=Ns, x_i=Is, mL=Lm, mN=Nm, mI=Im)
    expected, got = set(ROUTES), set(routes.keys())
    if expected != got:
        raise RuntimeError(f"[make_route_inputs_mult] Route key mismatch. missing={expected - got}, extra={got - expected}")
    return routes


class RoutePrimaryProjector(nn.Module):
    """routing_and_heads.py:101-121: 10 independent Linear(d_in -> pc_dim+1); pose | sigmoid(act)."""

    def __init__(self, d_in: int, pc_dim: int):
        super().__init__()
        self.d_in = int(d_in)
        self.pc_dim = int(pc_dim)
        if self.d_in != 256 or self.pc_dim != 32:
            raise NotImplementedError("B200 routing is specialised for d_in=256, pc_dim=32")
        self.proj = nn.ModuleDict({r: nn.Linear(self.d_in, self.pc_dim + 1, bias=True) for r in ROUTES})

    def _weights(self):
        return [self.proj[r].weight for r in ROUTES], [self.proj[r].bias for r in ROUTES]

    def forward(self, route_embs):
        """routing_and_heads.py:111-121: (poses [B,10,pc_dim], acts [B,10,1]) of the projector alone.  The hot path
        (forward_capsule_from_route_dict) runs the same projection inside the routing kernel; this entry serves callers
        that use the projector by itself (csrc/projector.cuh, fp32)."""
        embs = [route_embs[r] for r in ROUTES]                 # KeyError on a missing route, like the reference
        _check_route_embs(embs, self.d_in)
        pw, pb = self._weights()
        return ops.ProjectorFn.apply(*embs, *pw, *pb)


def _check_route_embs(embs, d_in):
    """Shapes must be settled before raw pointers go to the C ABI; the reference fails in F.linear / torch.stack
    (RuntimeError) for the same inputs."""
    B = embs[0].shape[0]
    dev = embs[0].device
    for r, x in zip(ROUTES, embs):
        if x.dim() != 2 or x.shape[1] != d_in:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({'x'.join(map(str, x.shape))} and {d_in}x33): "
                               f"route '{r}' must be [B,{d_in}]")
        if x.shape[0] != B:
            raise RuntimeError(f"stack expects each tensor to be equal size, but route '{r}' has batch {x.shape[0]}, "
                               f"route '{ROUTES[0]}' has {B}")
        if x.device != dev:
            raise RuntimeError(f"Expected all tensors to be on the same device, but route '{r}' is on {x.device} and "
                               f"route '{ROUTES[0]}' on {dev}")


class RouteDimAdapter(nn.Module):
    """routing_and_heads.py:124-155.  Identity for every route when the dims are equal (always, in
    the reference configuration: zero parameters)."""

    def __init__(self, d_in: int, d_l: int, d_n: int, d_i: int):
        super().__init__()
        d_in, d_l, d_n, d_i = int(d_in), int(d_l), int(d_n), int(d_i)
        if not (d_in == d_l == d_n == d_i):
            raise NotImplementedError("RouteDimAdapter with differing dims is not used by the reference drivers")
        self.adapt = nn.ModuleDict({r: nn.Identity() for r in
                                    ["L", "LN", "LI", "LNI", "N", "NL", "NI", "I", "IL", "IN"]})

    def forward(self, route_embs_in: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {r: self.adapt[r](route_embs_in[r]) for r in ROUTES}


def _expand_route_mask(route_mask, B, device):
    if route_mask is None:
        return None
    rm = route_mask
    if rm.ndim == 1:
        rm = rm.view(1, -1).expand(B, -1)
    elif rm.ndim != 2:
        raise ValueError(f"route_mask must be [R] or [B,R], got {tuple(rm.shape)}")
    if rm.shape[1] != len(ROUTES) or rm.shape[0] != B:
        # the kernels index route_mask[b*10 + r]; the reference fails in the broadcast `acts * mask` with this error
        d = 1 if rm.shape[1] != len(ROUTES) else 0
        raise RuntimeError(f"The size of tensor a ({(B, len(ROUTES))[d]}) must match the size of tensor b ({rm.shape[d]}) at "
                           f"non-singleton dimension {d}: route_mask must be [{len(ROUTES)}] or [{B},{len(ROUTES)}]")
    return rm.to(device=device, dtype=torch.float32).contiguous()


class _CapsuleHeadBase(nn.Module):
    VARIANT = "mort"

    def __init__(self, pc_dim: int, mc_caps_dim: int, num_routing: int, dp: float = 0.0, act_type: str = "ONES",
                 layer_norm: bool = False, dim_pose_to_vote: int = 0, num_classes: int = 25):
        super().__init__()
        if pc_dim != 32 or mc_caps_dim != 64:
            raise NotImplementedError("B200 routing is specialised for pc_dim=32, mc_caps_dim=64")
        if not 1 <= num_classes <= 32:
            raise NotImplementedError("num_classes must be in [1, 32]")
        if not 1 <= int(num_routing) <= 4:
            raise NotImplementedError("num_routing must be in [1, 4]")
        if dp != 0.0:
            raise NotImplementedError("capsule dropout > 0 is not implemented (reference uses 0.0)")
        self.in_n_capsules = len(ROUTES)
        self.in_d_capsules = pc_dim
        self.out_n_capsules = num_classes
        self.out_d_capsules = mc_caps_dim
        self.num_routing = int(num_routing)
        self.capsule = CapsuleFC(in_n_capsules=self.in_n_capsules, in_d_capsules=self.in_d_capsules,
                                 out_n_capsules=self.out_n_capsules, out_d_capsules=self.out_d_capsules, n_rank=0,
                                 dp=dp, dim_pose_to_vote=dim_pose_to_vote, uniform_routing_coefficient=False,
                                 act_type=act_type, small_std=True)
        self.pose_to_mc = nn.Linear(self.in_d_capsules, self.out_d_capsules, bias=False)
        self.embedding = nn.Parameter(torch.zeros(self.out_n_capsules, self.out_d_capsules))
        self.bias = nn.Parameter(torch.zeros(self.out_n_capsules))
        self.nonlinear_act = nn.Sequential()

    def forward(self, prim_pose: torch.Tensor, prim_act: torch.Tensor, uniform_routing: bool = False,
                route_mask: Optional[torch.Tensor] = None):
        """Mort routing_and_heads.py:194-268 / Pheno :194-272.  Returns (logits, alpha, R_brk)."""
        if prim_act.dim() == 2:
            pass
        elif prim_act.dim() == 3 and prim_act.size(-1) == 1:
            prim_act = prim_act.squeeze(-1)
        else:
            raise ValueError(f"prim_act must be [B,len(ROUTES)] or [B,len(ROUTES),1], got {prim_act.shape}")
        if uniform_routing:
            raise NotImplementedError("uniform_routing=True is never used by the reference drivers")
        if route_mask is not None and route_mask.ndim not in (1, 2):
            raise ValueError(f"route_mask must be [R] or [B,R], got {tuple(route_mask.shape)}")
        B = prim_pose.shape[0]
        # the kernels read poses_in[b*320 + j] and acts_in[b*10 + r]: a wrong shape must never reach them (the reference
        # fails in CapsuleFC's einsum with a RuntimeError)
        if prim_pose.dim() != 3 or tuple(prim_pose.shape[1:]) != (len(ROUTES), self.in_d_capsules):
            raise RuntimeError(f"einsum(): prim_pose must be [B,{len(ROUTES)},{self.in_d_capsules}], got {tuple(prim_pose.shape)}")
        if tuple(prim_act.shape) != (B, len(ROUTES)):
            raise RuntimeError(f"einsum(): prim_act must be [{B},{len(ROUTES)}] to broadcast with prim_pose, got {tuple(prim_act.shape)}")
        if prim_act.device != prim_pose.device:
            raise RuntimeError("Expected all tensors to be on the same device")
        rm = _expand_route_mask(route_mask, B, prim_pose.device)
        cfg = (ops.VARIANT[self.VARIANT], self.num_routing, False, 1.0, 0.0, 1.0, True)
        logits, alpha, R, _, _ = ops.RoutingFn.apply(cfg, None, rm, self.capsule.w, self.pose_to_mc.weight,
                                                      self.embedding, self.bias, prim_pose, prim_act)
        # alpha = prim_act * mask is differentiable in the reference; rebuild it with tensor ops
        alpha_out = prim_act if rm is None else prim_act * rm
        return logits, alpha_out, R


def forward_capsule_from_route_dict(route_embs_in: Dict[str, torch.Tensor], projector: RoutePrimaryProjector,
                                    capsule_head: _CapsuleHeadBase, *, acts_override: Optional[torch.Tensor] = None,
                                    route_mask: Optional[torch.Tensor] = None, act_temperature: float = 1.0,
                                    detach_priors: bool = False, return_routing: bool = True
                                    ) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor], Optional[torch.Tensor]]:
    """routing_and_heads.py:271-369: projector + activation priors + capsule routing + head in one
    persistent kernel launch.  Returns (logits [B,K], prim_acts [B,10] detached, route_embs, R [B,10,K])."""
    expected, got = set(ROUTES), set(route_embs_in.keys())
    if expected != got:
        raise RuntimeError(f"Route key mismatch. missing={expected - got}, extra={got - expected}")
    route_embs: Dict[str, torch.Tensor] = {}
    for r in ROUTES:
        x = route_embs_in[r]
        if not torch.is_tensor(x):
            raise TypeError(f"route_embs_in['{r}'] must be a Tensor, got {type(x)}")
        if x.dim() == 3 and x.size(1) == 1:
            x = x.squeeze(1)
        if x.dim() != 2:
            raise ValueError(f"route_embs_in['{r}'] must be [B,d] (or [B,1,d]), got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            x = x.float()
        route_embs[r] = x
    B = route_embs[ROUTES[0]].shape[0]
    dev = route_embs[ROUTES[0]].device
    _check_route_embs([route_embs[r] for r in ROUTES], projector.d_in)
    if route_mask is not None and route_mask.ndim not in (1, 2):
        raise ValueError(f"route_mask must be [R] or [B,R], got {tuple(route_mask.shape)}")
    rm = _expand_route_mask(route_mask, B, dev)
    ao = None
    if acts_override is not None:
        if acts_override.numel() != B * len(ROUTES) or acts_override.shape[0] != B:
            raise RuntimeError(f"einsum(): acts_override must be [{B},{len(ROUTES)},1] (or [{B},{len(ROUTES)}]), got "
                               f"{tuple(acts_override.shape)}")
        ao = acts_override.to(device=dev, dtype=torch.float32)     # [B,10,1] | [B,10]; gradients flow back (csrc/routing.cuh)
    floor = float(getattr(CFG, "route_prior_floor", 1e-3))
    ceil = float(getattr(CFG, "route_prior_ceiling", 0.999))
    lo = floor if floor > 0.0 else 0.0
    hi = ceil if ceil > 0.0 else 1.0
    pw, pb = projector._weights()
    cfg = (ops.VARIANT[capsule_head.VARIANT], capsule_head.num_routing, bool(detach_priors), float(act_temperature),
           lo, hi, False)
    logits, alpha, R, _, _ = ops.RoutingFn.apply(cfg, ao, rm, capsule_head.capsule.w, capsule_head.pose_to_mc.weight,
                                                  capsule_head.embedding, capsule_head.bias,
                                                  *[route_embs[r] for r in ROUTES], *pw, *pb)
    prim_acts = alpha.detach()
    if not return_routing:
        R = None
    return logits, prim_acts, route_embs, R


def forward_capsule_from_multmodel(multmodel: nn.Module, x_l: torch.Tensor, x_n: torch.Tensor, x_i: torch.Tensor,
                                   projector: RoutePrimaryProjector, capsule_head: _CapsuleHeadBase, *,
                                   mL: Optional[torch.Tensor] = None, mN: Optional[torch.Tensor] = None,
                                   mI: Optional[torch.Tensor] = None, route_adapter: Optional[RouteDimAdapter] = None,
                                   acts_override: Optional[torch.Tensor] = None,
                                   route_mask: Optional[torch.Tensor] = None, act_temperature: float = 1.0,
                                   detach_priors: bool = False, return_routing: bool = True):
    """routing_and_heads.py:372-409."""
    route_embs_in = multmodel(x_l, x_n, x_i, mL=mL, mN=mN, mI=mI)
    expected, got = set(ROUTES), set(route_embs_in.keys())
    if expected != got:
        raise RuntimeError(f"[mult->caps] Route key mismatch. missing={expected - got}, extra={got - expected}")
    if route_adapter is not None:
        route_embs_in = route_adapter(route_embs_in)
    return forward_capsule_from_route_dict(route_embs_in=route_embs_in, projector=projector, capsule_head=capsule_head,
                                           acts_override=acts_override, route_mask=route_mask,
                                           act_temperature=act_temperature, detach_priors=detach_priors,
                                           return_routing=return_routing)
