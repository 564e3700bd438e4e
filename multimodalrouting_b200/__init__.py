"""multimodalrouting_b200: B200-native (sm_100a) route fusion + capsule routing.

Drop-in for the hot path of AI-for-Health-Data/MultimodalRouting:

    from multimodalrouting_b200.MortModel import mult_model, routing_and_heads     # K=2 mortality
    from multimodalrouting_b200.PhenoModel import mult_model, routing_and_heads    # 25 phenotypes

Same nn.Module names, constructor/forward signatures and state_dict keys as the reference's
MIMIC-IV/{Mort,Pheno}Model/Paired_Cross_Attention modules; the arithmetic runs in hand-written CUDA
kernels (csrc/) behind the C ABI of include/mmr_b200.h.  No CPU / PyTorch fallback exists.
"""
from .env_config import ROUTES  # noqa: F401
from .mult_model import MULTModel  # noqa: F401

__version__ = "0.1.0"
