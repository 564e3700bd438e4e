"""Parameter container mirroring the reference MultiheadAttention
(MIMIC-IV/PhenoModel/Paired_Cross_Attention/multihead_attention.py:6-46).

Same attribute names, shapes and initialisation so state_dicts are interchangeable; the arithmetic
(multihead_attention.py:48-148) runs inside the fused route-fusion kernels driven by MULTModel.
"""
import torch
from torch import nn
from torch.nn import Parameter


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, attn_dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False):
        super().__init__()
        if add_bias_kv or add_zero_attn or not bias:
            raise NotImplementedError("the B200 path implements the configuration the reference uses: "
                                      "bias=True, add_bias_kv=False, add_zero_attn=False")
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.attn_dropout = attn_dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.scaling = self.head_dim ** -0.5
        self.in_proj_weight = Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = False
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, *args, **kwargs):
        raise RuntimeError("MultiheadAttention is fused into MULTModel.forward on the B200 path; "
                           "call the enclosing MULTModel")
