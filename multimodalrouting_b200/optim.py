"""Training tail of the hot path: gradient clipping, the finite-gradient guard, AdamW and
the EMA of the weights as three device-side stages (csrc/tail.cuh) that never synchronise the host.

Replaces, with the same semantics,

    torch.nn.utils.clip_grad_norm_(trainable_params, max_norm=grad_clip)      MortModel/.../main.py:3147,3156
    if not grads_are_finite(trainable_params): skip                           main.py:3148-3151, 3157-3160
    optimizer.step()            # torch.optim.AdamW(lr, weight_decay)         main.py:2886-2890, 3161
    ema.update()                # class EMA                                   main.py:58-108, 3162-3163

`FusedAdamW` is a torch.optim.Optimizer (param_groups / state / state_dict keep torch.optim.AdamW's layout, so LR
schedulers and checkpoints work unchanged); `EMA` has the reference class's interface.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import OPT_STATE_BYTES, OptHyper, OptTensor
from .ops import _require_cuda, _stream


def _table(entries):
    arr = (OptTensor * len(entries))()
    for i, (p, g, m, v, e, n) in enumerate(entries):
        arr[i].p, arr[i].g, arr[i].m, arr[i].v, arr[i].ema, arr[i].n = p, g, m, v, e, n
    return arr


class EMA:
    """Exponential moving average of module state_dicts; interface of the reference's EMA (main.py:58-108)."""

    def __init__(self, model_list, decay: float = 0.999):
        self.decay = float(decay)
        self.model_list = list(model_list)
        self.shadow = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in self.model_list]
        self.backup = None
        self._rebinds = 0          # bumped whenever a shadow entry is replaced by a new tensor (FusedAdamW caches their addresses)

    def shadow_by_storage(self) -> Dict[int, Tensor]:
        """live tensor data_ptr -> its shadow tensor (how FusedAdamW finds the EMA slot of a parameter)."""
        out = {}
        for m, sh in zip(self.model_list, self.shadow):
            for k, v in m.state_dict().items():
                if k in sh and torch.is_floating_point(v):
                    out[v.data_ptr()] = sh[k]
        return out

    @torch.no_grad()
    def update(self, _exclude: Optional[set] = None):
        """shadow = decay * shadow + (1 - decay) * value for every float entry (one multi-tensor launch per 384
        tensors); non-float entries are copied.  `_exclude`: data_ptrs already updated by FusedAdamW.step(ema=...)."""
        entries = []
        for m, sh in zip(self.model_list, self.shadow):
            for k, v in m.state_dict().items():
                if k not in sh:
                    sh[k] = v.detach().clone()
                    self._rebinds += 1
                    continue
                if not torch.is_floating_point(v):
                    sh[k].copy_(v)
                    continue
                if _exclude is not None and v.data_ptr() in _exclude:
                    continue
                if sh[k].device != v.device or sh[k].dtype != v.dtype:
                    sh[k] = sh[k].to(device=v.device, dtype=v.dtype)
                    self._rebinds += 1
                _require_cuda(v)
                if v.dtype != torch.float32 or not v.is_contiguous() or not sh[k].is_contiguous():
                    raise ValueError("EMA.update expects contiguous fp32 tensors")
                entries.append((v.data_ptr(), None, None, None, sh[k].data_ptr(), v.numel()))
        if entries:
            rc = _lib.load().mmr_ema_update(_table(entries), len(entries), self.decay, _stream())
            _lib.check(rc, "mmr_ema_update")

    @torch.no_grad()
    def apply_to(self):
        self.backup = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in self.model_list]
        for m, sh in zip(self.model_list, self.shadow):
            m.load_state_dict(sh, strict=True)

    @torch.no_grad()
    def restore(self):
        if self.backup is None:
            return
        for m, bk in zip(self.model_list, self.backup):
            m.load_state_dict(bk, strict=True)
        self.backup = None


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (amsgrad=False, maximize=False) with the gradient clipping, the finite-gradient guard and the
    EMA update fused into its step:

        opt.step(max_norm=0.3, ema=ema)

    = clip_grad_norm_(params, 0.3); if grads finite: AdamW step; ema.update() -- in one norm pass, one scalar kernel
    and one update pass over (param, grad, exp_avg, exp_avg_sq, shadow).  The total norm, the clip coefficient, the
    skip flag and the step count live on the device (`total_norm`, `skipped`, `step_count` return 0-dim views), so
    the step neither syncs the host nor breaks CUDA-graph capture.

    Two differences from the reference's sequence (main.py:3143-3165), both invisible to the parameters and the EMA of the
    trained tensors: the clip coefficient is applied inside the update, so `p.grad` still holds the UNCLIPPED gradient after
    `step()` (read `clip_coef` / `total_norm` instead of re-measuring); and on a skipped (non-finite) step the EMA entries
    the optimizer does not own -- buffers, frozen tensors -- still take one averaging step towards their (unchanged) values,
    where the reference `continue`s before `ema.update()`."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError("lr, eps and weight_decay must be >= 0")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"Invalid betas: {betas}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._dev = None           # mmr_opt_state, device resident
        self._ema_map = (None, None, -1)
        self._tables = {}          # group index -> (pointer key, ctypes table): rebuilt only when a pointer moves

    # ---- device-side scalars -------------------------------------------------------------------------
    def _state_tensor(self, device) -> Tensor:
        if self._dev is None:
            self._dev = torch.zeros(OPT_STATE_BYTES, dtype=torch.uint8, device=device)
        elif self._dev.device != device:
            raise ValueError("FusedAdamW: all parameters must live on one device")
        return self._dev

    @property
    def total_norm(self) -> Tensor:
        """Gradient norm of the last step (before clipping), 0-dim fp32 view on the device."""
        return self._dev[24:28].view(torch.float32)[0]

    @property
    def clip_coef(self) -> Tensor:
        return self._dev[28:32].view(torch.float32)[0]

    @property
    def step_count(self) -> Tensor:
        return self._dev[32:36].view(torch.int32)[0]

    @property
    def skipped(self) -> Tensor:
        """1 when the last step saw non-finite gradients and left every tensor untouched."""
        return self._dev[36:40].view(torch.int32)[0]

    # ---- the step ------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, *, max_norm: float = 0.0, ema: Optional[EMA] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        betas = self.param_groups[0]["betas"]
        shadow = None
        if ema is not None:
            if self._ema_map[0] is not ema or self._ema_map[2] != ema._rebinds:     # a rebound shadow tensor invalidates the map
                self._ema_map = (ema, ema.shadow_by_storage(), ema._rebinds)
            shadow = self._ema_map[1]
        tables: List = []
        done = set()
        dev = None
        keep = []
        for gi, group in enumerate(self.param_groups):
            if tuple(group["betas"]) != tuple(betas):
                raise ValueError("FusedAdamW: all parameter groups must share betas (one bias correction per step)")
            entries = []
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                _require_cuda(p, g)
                if g.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                if p.dtype != torch.float32 or g.dtype != torch.float32:
                    raise ValueError("FusedAdamW expects fp32 parameters and gradients")
                if not p.is_contiguous():
                    raise ValueError("FusedAdamW expects contiguous parameters")
                if not g.is_contiguous():
                    g = g.contiguous()
                    keep.append(g)
                dev = p.device
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                e = None
                if shadow is not None:
                    sh = shadow.get(p.data_ptr())
                    if sh is None:
                        raise ValueError("FusedAdamW.step(ema=...): a parameter is not tracked by the EMA object")
                    if sh.dtype != torch.float32 or sh.device != p.device or not sh.is_contiguous():
                        raise ValueError("EMA shadow tensors must be contiguous fp32 on the parameter's device")
                    e = sh.data_ptr()
                    done.add(p.data_ptr())
                entries.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), e,
                                p.numel()))
            if entries:
                key = tuple(entries)
                hit = self._tables.get(gi)
                if hit is None or hit[0] != key:
                    hit = (key, _table(entries))
                    self._tables[gi] = hit
                tables.append((group, hit[1], len(entries)))
        if not tables:
            return loss
        state = self._state_tensor(dev)
        sp, stream = state.data_ptr(), _stream()
        for _, tab, n in tables:
            _lib.check(lib.mmr_grad_sqnorm(tab, n, sp, stream), "mmr_grad_sqnorm")
        hp = OptHyper(0.0, betas[0], betas[1], 0.0, 0.0, float(max_norm or 0.0), 0.0, 0, 0)
        _lib.check(lib.mmr_opt_prepare(C.byref(hp), sp, stream), "mmr_opt_prepare")
        for group, tab, n in tables:
            hp = OptHyper(float(group["lr"]), betas[0], betas[1], float(group["eps"]), float(group["weight_decay"]),
                          float(max_norm or 0.0), ema.decay if ema is not None else 0.0, int(ema is not None), 0)
            _lib.check(lib.mmr_opt_apply(tab, n, C.byref(hp), sp, stream), "mmr_opt_apply")
        if ema is not None:
            ema.update(_exclude=done)      # entries the optimizer does not own (buffers, frozen tensors)
        # the kernels wrote through raw pointers: tell autograd (and every cache keyed on tensor versions, e.g. MULTModel's
        # packed weights) that the parameters changed, as an in-place torch op would
        torch.autograd.graph.increment_version([p for g in self.param_groups for p in g["params"] if p.grad is not None])
        return loss

    # ---- torch.optim.AdamW-compatible checkpoints ---------------------------------------------------------
    def state_dict(self):
        if self._dev is not None:
            n = float(int(self.step_count))            # host sync: checkpoint time only
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(n, dtype=torch.float32, device=st["exp_avg"].device)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [int(float(st["step"])) for st in self.state.values() if "step" in st]
        if steps:
            if len(set(steps)) != 1:
                raise ValueError("FusedAdamW: parameters with different step counts cannot share one bias correction")
            dev = next(iter(self.state.values()))["exp_avg"].device
            self._state_tensor(dev)[32:36].view(torch.int32).fill_(steps[0])


def grads_are_finite(optimizer: FusedAdamW) -> bool:
    """Host-side read of the guard of the last step (main.py:46-57); forces a device sync."""
    return int(optimizer.skipped) == 0
