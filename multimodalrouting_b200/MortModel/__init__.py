"""Drop-in modules for MIMIC-IV/MortModel/Paired_Cross_Attention (variant 'mort')."""
from .. import capsule_layers, env_config, mult_model, multihead_attention, position_embedding, transformer  # noqa: F401
from . import routing_and_heads  # noqa: F401
