"""Positional table of the reference, as a precomputed constant.

Reference: MIMIC-IV/PhenoModel/Paired_Cross_Attention/position_embedding.py:68-117 and its call
site MIMIC-IV/MortModel/Paired_Cross_Attention/transformer.py:65-72.  The caller passes a
LongTensor of ones, so the sinusoid table is cast to int64 (truncated toward zero) before it is
added: the positional term is a {-1,0,1} table.  The kernels take it as a caller-owned fp32 tensor.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
from torch._subclasses.fake_tensor import FakeTensor

_CACHE: Dict[Tuple[int, int, str], torch.Tensor] = {}


def truncated_sinusoid_table(T: int, dim: int, device) -> torch.Tensor:
    """fp32 [T, dim]: rows for positions 1..T (padding_idx = 0), int-truncated."""
    key = (int(T), int(dim), str(device))
    tab = _CACHE.get(key)
    if tab is None:
        half = dim // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float32) * (-(math.log(10000.0) / (half - 1))))
        ang = torch.arange(T + 1, dtype=torch.float32).unsqueeze(1) * freq.unsqueeze(0)
        full = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
        if dim % 2 == 1:
            full = torch.cat([full, torch.zeros(T + 1, 1)], dim=1)
        full[0, :] = 0.0
        tab = full.to(torch.int64)[1:T + 1].to(torch.float32).contiguous().to(device)
        if not isinstance(tab, FakeTensor):     # never let a tracing-time (fake) table leak into later real calls
            _CACHE[key] = tab
    return tab


class SinusoidalPositionalEmbedding(torch.nn.Module):
    """Parameter-free placeholder keeping the reference module tree (`embed_positions`)."""

    def __init__(self, embedding_dim: int, padding_idx: int = 0, left_pad: bool = False, init_size: int = 128):
        super().__init__()
        self.embedding_dim = int(embedding_dim)
        self.padding_idx = int(padding_idx)
        self.left_pad = bool(left_pad)

    def table(self, T: int, device) -> torch.Tensor:
        return truncated_sinusoid_table(T, self.embedding_dim, device)

    def max_positions(self) -> int:
        return int(1e5)
