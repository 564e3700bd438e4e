"""Evaluation-time consumers of the routing outputs.

`RoutingStatsAccumulator` replaces the per-batch host copies of the reference's `evaluate_epoch`
(MIMIC-IV/MortModel/Paired_Cross_Attention/main.py:1916-1933: `rc_raw.detach().float().cpu()`, `rc_report...cpu()`,
`prim_acts...cpu()` and three host-side sums per batch; averages at :2013-2016): the [routes, labels] sums of the raw and
reported routing coefficients, of the effective weights `rc_raw * prim_acts` and of the route activations accumulate on the
device (one launch per batch, no synchronisation) and come to the host once per split.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib
from .ops import _ptr, _require_cuda, _stream

N_ROUTES = 10


class RoutingStatsAccumulator:
    def __init__(self, num_labels: int, device="cuda"):
        if not 1 <= int(num_labels) <= 32:
            raise ValueError("num_labels must be in [1, 32]")
        self.K = int(num_labels)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RoutingStatsAccumulator runs only on a CUDA device (there is no CPU fallback)")
        self.sums = torch.zeros(3 * N_ROUTES * self.K + N_ROUTES, dtype=torch.float32, device=self.device)
        self.count = torch.zeros(1, dtype=torch.int64, device=self.device)

    def reset(self) -> None:
        self.sums.zero_()
        self.count.zero_()

    def update(self, rc_raw: Tensor, rc_report: Optional[Tensor], prim_acts: Tensor) -> None:
        """rc_raw [B,10,K] (fp32 or bf16, as the capsule head returns it), rc_report [B,10,K] or None, prim_acts [B,10]."""
        _require_cuda(rc_raw, prim_acts, rc_report)
        B = rc_raw.shape[0]
        if tuple(rc_raw.shape) != (B, N_ROUTES, self.K):
            raise ValueError(f"rc_raw must be [B,{N_ROUTES},{self.K}], got {tuple(rc_raw.shape)}")
        if prim_acts.dim() == 3 and prim_acts.shape[-1] == 1:
            prim_acts = prim_acts.squeeze(-1)
        if tuple(prim_acts.shape) != (B, N_ROUTES):
            raise ValueError(f"prim_acts must be [B,{N_ROUTES}], got {tuple(prim_acts.shape)}")
        if rc_report is not None and tuple(rc_report.shape) != tuple(rc_raw.shape):
            raise ValueError("rc_report must have rc_raw's shape")
        if B == 0:
            return
        raw = rc_raw.detach()
        if raw.dtype not in (torch.float32, torch.bfloat16):
            raw = raw.float()
        raw = raw.contiguous()
        rep = None if rc_report is None else rc_report.detach().float().contiguous()
        pa = prim_acts.detach().float().contiguous()
        rc = _lib.load().mmr_routing_stats_accumulate(_ptr(raw), 1 if raw.dtype == torch.bfloat16 else 0, _ptr(rep), _ptr(pa), B,
                                                      self.K, _ptr(self.sums), _ptr(self.count), _stream())
        _lib.check(rc, "mmr_routing_stats_accumulate")

    def result(self) -> Dict[str, Tensor]:
        """One device-to-host read: sums and per-sample averages (main.py:2013-2016 `rep_sum_mat / max(1, num_samples)`)."""
        s = self.sums.cpu()
        n = int(self.count.cpu())
        RK = N_ROUTES * self.K
        raw, rep, eff = (s[i * RK:(i + 1) * RK].view(N_ROUTES, self.K) for i in range(3))
        act = s[3 * RK:]
        d = max(1, n)
        return {"num_samples": n, "rc_raw_sum": raw, "rc_report_sum": rep, "eff_sum": eff, "prim_act_sum": act,
                "avg_rc_raw": raw / d, "avg_rc_report": rep / d, "avg_eff": eff / d, "avg_prim_act": act / d}
