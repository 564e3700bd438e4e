"""Parameter containers mirroring the reference TransformerEncoder / TransformerEncoderLayer
(MIMIC-IV/MortModel/Paired_Cross_Attention/transformer.py:11-54,118-147,243-248).

The forward arithmetic (transformer.py:56-115,149-216) is executed by the fused kernels that
MULTModel.forward launches for all six directions at once; these classes only own the parameters
under the reference's names so checkpoints / EMA / optimizers see an identical module tree.
"""
import math

from torch import nn

from .multihead_attention import MultiheadAttention
from .position_embedding import SinusoidalPositionalEmbedding


def Linear(in_features, out_features, bias=True):
    m = nn.Linear(in_features, out_features, bias)
    nn.init.xavier_uniform_(m.weight)
    if bias:
        nn.init.constant_(m.bias, 0.0)
    return m


class TransformerEncoderLayer(nn.Module):
    def __init__(self, embed_dim, num_heads=4, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1, attn_mask=False):
        super().__init__()
        self.embed_dim = int(embed_dim)
        self.num_heads = int(num_heads)
        self.attn_mask = bool(attn_mask)
        self.self_attn = MultiheadAttention(embed_dim=self.embed_dim, num_heads=self.num_heads,
                                            attn_dropout=float(attn_dropout))
        self.relu_dropout = float(relu_dropout)
        self.res_dropout = float(res_dropout)
        self.normalize_before = True
        self.fc1 = Linear(self.embed_dim, 4 * self.embed_dim)
        self.fc2 = Linear(4 * self.embed_dim, self.embed_dim)
        self.layer_norms = nn.ModuleList([nn.LayerNorm(self.embed_dim) for _ in range(2)])

    def forward(self, *args, **kwargs):
        raise RuntimeError("TransformerEncoderLayer is fused into MULTModel.forward on the B200 path")


class TransformerEncoder(nn.Module):
    def __init__(self, embed_dim, num_heads, layers, attn_dropout=0.0, relu_dropout=0.0, res_dropout=0.0,
                 embed_dropout=0.0, attn_mask=False, use_positional=True, padding_idx=0, left_pad=False):
        super().__init__()
        if not use_positional:
            raise NotImplementedError("the B200 path always adds the (int-truncated) positional table")
        self.dropout = float(embed_dropout)
        self.attn_dropout = float(attn_dropout)
        self.embed_dim = int(embed_dim)
        self.embed_scale = math.sqrt(self.embed_dim)
        self.attn_mask = bool(attn_mask)
        self.embed_positions = SinusoidalPositionalEmbedding(self.embed_dim, padding_idx=padding_idx, left_pad=left_pad)
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(embed_dim=self.embed_dim, num_heads=num_heads, attn_dropout=attn_dropout,
                                    relu_dropout=relu_dropout, res_dropout=res_dropout, attn_mask=attn_mask)
            for _ in range(int(layers))])
        self.normalize = True
        self.layer_norm = nn.LayerNorm(self.embed_dim)

    def forward(self, *args, **kwargs):
        raise RuntimeError("TransformerEncoder is fused into MULTModel.forward on the B200 path")
