"""routing_and_heads for the Mort variant (reference: MIMIC-IV/MortModel/Paired_Cross_Attention/routing_and_heads.py)."""
from .._routing_impl import (RouteDimAdapter, RoutePrimaryProjector, _CapsuleHeadBase,  # noqa: F401
                             forward_capsule_from_multmodel, forward_capsule_from_route_dict,
                             make_route_inputs_mult, route_given_pheno)
from ..env_config import CFG, DEVICE, ROUTES  # noqa: F401
from ..mult_model import MULTModel  # noqa: F401


class CapsuleMortalityHead(_CapsuleHeadBase):
    VARIANT = "mort"


__all__ = ["RoutePrimaryProjector", "RouteDimAdapter", "make_route_inputs_mult", "CapsuleMortalityHead",
           "forward_capsule_from_route_dict", "forward_capsule_from_multmodel", "route_given_pheno"]
