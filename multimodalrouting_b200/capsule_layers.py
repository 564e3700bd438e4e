"""CapsuleFC parameter container mirroring the reference
(MIMIC-IV/MortModel/Paired_Cross_Attention/capsule_layers.py:7-52).

forward() (capsule_layers.py:75-117) runs inside the persistent routing kernel (csrc/routing.cuh)
launched by CapsuleMortalityHead / forward_capsule_from_route_dict.
"""
import numpy as np
import torch
from torch import nn


class CapsuleFC(nn.Module):
    def __init__(self, in_n_capsules, in_d_capsules, out_n_capsules, out_d_capsules, n_rank, dp, dim_pose_to_vote,
                 uniform_routing_coefficient=False, act_type="EM", small_std=False):
        super().__init__()
        self.in_n_capsules = in_n_capsules
        self.in_d_capsules = in_d_capsules
        self.out_n_capsules = out_n_capsules
        self.out_d_capsules = out_d_capsules
        self.n_rank = n_rank
        self.weight_init_const = np.sqrt(out_n_capsules / (in_d_capsules * in_n_capsules))
        self.w = nn.Parameter(self.weight_init_const *
                              torch.randn(in_n_capsules, in_d_capsules, out_n_capsules, out_d_capsules))
        self.dropout_rate = float(dp)
        if not small_std:
            raise NotImplementedError("layer norm will destroy interpretability, thus not available")
        self.nonlinear_act = nn.Sequential()
        self.drop = nn.Dropout(self.dropout_rate)
        self.scale = 1.0 / (out_d_capsules ** 0.5)
        self.act_type = act_type
        if act_type == "EM":
            self.beta_u = nn.Parameter(torch.randn(out_n_capsules))
            self.beta_a = nn.Parameter(torch.randn(out_n_capsules))
        elif act_type == "Hubert":
            raise NotImplementedError("act_type='Hubert' is not used by the reference drivers")
        self.uniform_routing_coefficient = bool(uniform_routing_coefficient)

    def forward(self, *args, **kwargs):
        raise RuntimeError("CapsuleFC is fused into the capsule-routing kernel; call CapsuleMortalityHead")
