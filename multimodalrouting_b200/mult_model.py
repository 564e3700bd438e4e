"""MULTModel: drop-in for MIMIC-IV/{Mort,Pheno}Model/Paired_Cross_Attention/mult_model.py.

Same constructor signature, module tree / state_dict keys and ``forward(x_l, x_n, x_i, mL, mN, mI)
-> {route: [B, d]}`` contract (mult_model.py:8-11,116-193).  forward() is ONE autograd node that
launches the hand-written sm_100a kernels: the six directional cross-modal encoders run as grouped
launches (all directions per launch), followed by masked-mean pooling, the pair projections and the
LNI trimodal composition.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops
from .position_embedding import truncated_sinusoid_table
from .transformer import TransformerEncoder

ROUTES = ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]
_CROSS = ["trans_l_with_n", "trans_l_with_i", "trans_n_with_l", "trans_n_with_i", "trans_i_with_l", "trans_i_with_n"]


class MULTModel(nn.Module):
    def __init__(self, orig_d_l, orig_d_n, orig_d_i, d_l, d_n, d_i, ionly, nonly, lonly, num_heads, layers,
                 self_layers, attn_dropout, attn_dropout_n, attn_dropout_i, relu_dropout, res_dropout, out_dropout,
                 embed_dropout, attn_mask):
        super().__init__()
        self.orig_d_l, self.orig_d_n, self.orig_d_i = orig_d_l, orig_d_n, orig_d_i
        self.d_l, self.d_n, self.d_i = d_l, d_n, d_i
        self.ionly, self.nonly, self.lonly = ionly, nonly, lonly
        self.num_heads = num_heads
        self.layers = layers
        self.self_layers = self_layers
        self.attn_dropout, self.attn_dropout_n, self.attn_dropout_i = attn_dropout, attn_dropout_n, attn_dropout_i
        self.relu_dropout, self.res_dropout = relu_dropout, res_dropout
        self.out_dropout, self.embed_dropout = out_dropout, embed_dropout
        self.attn_mask = attn_mask
        # The kernels are specialised for the reference's fixed hyper-parameters.
        if not (d_l == d_n == d_i == 256) or num_heads != 8:
            raise NotImplementedError("B200 route fusion is specialised for d_l=d_n=d_i=256, num_heads=8")
        if self_layers != 0:
            raise NotImplementedError("B200 route fusion implements self_layers=0 (unimodal routes are "
                                      "LayerNorm(embed) pooled), as in the reference configuration")
        if attn_mask:
            raise NotImplementedError("attn_mask=True (future mask) is not used by the reference drivers")
        if not 1 <= int(layers) <= 8:
            raise NotImplementedError("layers must be in [1, 8]")
        for od in (orig_d_l, orig_d_n, orig_d_i):
            if od % 16:
                raise NotImplementedError("input feature dims must be multiples of 16")

        self.proj_l = nn.Conv1d(self.orig_d_l, self.d_l, kernel_size=1, padding=0, bias=False)
        self.proj_n = nn.Conv1d(self.orig_d_n, self.d_n, kernel_size=1, padding=0, bias=False)
        self.proj_i = nn.Conv1d(self.orig_d_i, self.d_i, kernel_size=1, padding=0, bias=False)
        self.trans_l = self.get_network(self_type="l_only", layers=self.self_layers)
        self.trans_n = self.get_network(self_type="n_only", layers=self.self_layers)
        self.trans_i = self.get_network(self_type="i_only", layers=self.self_layers)
        self.trans_l_with_n = self.get_network(self_type="ln")
        self.trans_l_with_i = self.get_network(self_type="li")
        self.trans_n_with_l = self.get_network(self_type="nl")
        self.trans_n_with_i = self.get_network(self_type="ni")
        self.trans_i_with_l = self.get_network(self_type="il")
        self.trans_i_with_n = self.get_network(self_type="in")
        self.proj_n_to_l = nn.Identity()
        self.proj_i_to_l = nn.Identity()
        self.proj_pair_ln = nn.Linear(2 * self.d_l, self.d_l, bias=True)
        self.proj_pair_li = nn.Linear(2 * self.d_l, self.d_l, bias=True)
        self.proj_pair_ni = nn.Linear(2 * self.d_l, self.d_l, bias=True)
        self.final_lni = nn.Linear(3 * self.d_l, self.d_l, bias=True)
        #: "auto" follows torch.autocast like the reference (fp32 outside, bf16 kernels inside)
        self.compute_dtype = "auto"
        self._plist: Optional[List[nn.Parameter]] = None
        #: The kernels read compute-type copies of the GEMM weights ("packed" weights).  They are rebuilt only when a
        #: parameter changed (tensor version counters + addresses): evaluation and repeated forward passes between
        #: optimizer steps pack once.  Inside a CUDA-graph capture version counters cannot be consulted at replay time, so
        #: the packing is captured with the forward unless the caller declares the weights constant for the life of the
        #: graph (``static_weights = True``: inference graphs, or a forward/backward graph with the optimizer outside it
        #: that re-captures / calls ``refresh_packed_weights()`` after every update).
        self.static_weights = False
        self._packed = None          # (key, tensor)

    def get_network(self, self_type: str = "l", layers: int = -1):
        n_layers = self.layers if layers == -1 else layers
        q = self_type[0]
        if q == "l":
            embed_dim, attn_dropout = self.d_l, self.attn_dropout
        elif q == "n":
            embed_dim, attn_dropout = self.d_n, self.attn_dropout_n
        elif q == "i":
            embed_dim, attn_dropout = self.d_i, self.attn_dropout_i
        else:
            raise ValueError(f"Unknown network type: {self_type}")
        return TransformerEncoder(embed_dim=embed_dim, num_heads=self.num_heads, layers=n_layers,
                                  attn_dropout=attn_dropout, relu_dropout=self.relu_dropout,
                                  res_dropout=self.res_dropout, embed_dropout=self.embed_dropout,
                                  attn_mask=self.attn_mask)

    # parameter tensors in the order of include/mmr_b200.h (== registration / state_dict order)
    def _param_list(self) -> List[nn.Parameter]:
        ps = [self.proj_l.weight, self.proj_n.weight, self.proj_i.weight]
        for enc in (self.trans_l, self.trans_n, self.trans_i):
            ps += [enc.layer_norm.weight, enc.layer_norm.bias]
        for name in _CROSS:
            enc = getattr(self, name)
            for layer in enc.layers:
                ps += [layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias,
                       layer.self_attn.out_proj.weight, layer.self_attn.out_proj.bias,
                       layer.fc1.weight, layer.fc1.bias, layer.fc2.weight, layer.fc2.bias,
                       layer.layer_norms[0].weight, layer.layer_norms[0].bias,
                       layer.layer_norms[1].weight, layer.layer_norms[1].bias]
            ps += [enc.layer_norm.weight, enc.layer_norm.bias]
        for lin in (self.proj_pair_ln, self.proj_pair_li, self.proj_pair_ni, self.final_lni):
            ps += [lin.weight, lin.bias]
        return ps

    def _packed_weights(self, params, dtype, engine):
        """Cached packed weights for the current parameter versions, or None (= pack inside the op)."""
        p0 = params[0]
        if not p0.is_cuda or isinstance(p0, torch._subclasses.fake_tensor.FakeTensor):
            return None
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing and not self.static_weights:
            return None
        key = (dtype, engine, p0.device, tuple((p.data_ptr(), p._version) for p in params))
        hit = self._packed
        if hit is not None and (hit[0] == key or (capturing and hit[0][:3] == key[:3])):
            return hit[1]
        if capturing:
            return None          # nothing cached yet: pack inside the capture
        with torch.no_grad():
            packed = ops.route_fusion_pack([p.detach() for p in params], int(self.layers), dtype, engine)
        self._packed = (key, packed)
        return packed

    def refresh_packed_weights(self):
        """Re-packs IN PLACE into the cached buffer (whose address a captured graph may hold) after the parameters
        were updated outside that graph."""
        hit = self._packed
        if hit is None:
            return
        dtype, engine = hit[0][0], hit[0][1]
        params = self._param_list()
        with torch.no_grad():
            hit[1].copy_(ops.route_fusion_pack([p.detach() for p in params], int(self.layers), dtype, engine))
        self._packed = ((dtype, engine, params[0].device, tuple((p.data_ptr(), p._version) for p in params)), hit[1])

    def _ensure_float_mask(self, m, B, T, device):
        if m is None:
            return None
        if m.dim() == 1:
            m = m.unsqueeze(0).expand(B, -1)
        return m.to(device=device).float().contiguous()

    def _check_dropout(self):
        if self.training and any(p > 0 for p in (self.attn_dropout, self.attn_dropout_n, self.attn_dropout_i,
                                                 self.relu_dropout, self.res_dropout, self.embed_dropout)):
            raise NotImplementedError("dropout > 0 in training mode is not implemented by the fused kernels "
                                      "(every reference config uses dropout 0.0, SURVEY.md 0.4)")

    def forward(self, x_l, x_n, x_i, mL=None, mN=None, mI=None) -> Dict[str, torch.Tensor]:
        assert x_l.dim() == 3 and x_n.dim() == 3 and x_i.dim() == 3
        B, TL, _ = x_l.shape
        BN, TN, _ = x_n.shape
        BI, TI, _ = x_i.shape
        assert B == BN == BI
        if 0 in (B, TL, TN, TI):
            # the reference fails in MultiheadAttention's head reshape (multihead_attention.py:93-97) with this error
            raise RuntimeError(f"cannot reshape tensor of 0 elements: empty batch / modality (B={B}, TL={TL}, TN={TN}, TI={TI}) "
                               "is not a valid input of the cross-modal encoders")
        if x_l.shape[2] != self.orig_d_l or x_n.shape[2] != self.orig_d_n or x_i.shape[2] != self.orig_d_i:
            raise ValueError("input feature dims do not match orig_d_l / orig_d_n / orig_d_i")
        self._check_dropout()
        device = x_l.device
        mL = self._ensure_float_mask(mL, B, TL, device)
        mN = self._ensure_float_mask(mN, B, TN, device)
        mI = self._ensure_float_mask(mI, B, TI, device)
        for m, T in ((mL, TL), (mN, TN), (mI, TI)):
            # the reference fails in the broadcast `x * mask` (transformer.py:77-79) with a RuntimeError; the kernels
            # index the mask as [B, T], so a wrong shape must never reach them
            if m is not None and tuple(m.shape) != (B, T):
                raise RuntimeError(f"The size of the mask {tuple(m.shape)} must match the size of the token axis "
                                   f"[B, T] = [{B}, {T}]")
        pos = truncated_sinusoid_table(max(TL, TN, TI), self.d_l, device)
        params = self._param_list()
        # parameters that never take part (proj_* when orig_d == d) get no gradient, as in the reference
        plist = []
        for p, od in zip(params[:3], (self.orig_d_l, self.orig_d_n, self.orig_d_i)):
            plist.append(p.detach() if od == 256 else p)
        plist += params[3:]
        dtype = ops.resolve_dtype(self.compute_dtype)
        engine = ops.resolve_engine()
        outs = ops.route_fusion(x_l, x_n, x_i, mL, mN, mI, pos, plist, int(self.layers), dtype, engine,
                                self._packed_weights(params, dtype, engine))
        target_dtype = self.final_lni.weight.dtype
        return {r: (o if o.dtype == target_dtype else o.to(target_dtype)) for r, o in zip(ROUTES, outs)}
