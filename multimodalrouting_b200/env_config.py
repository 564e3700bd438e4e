"""The few constants of the reference's env_config that the hot path reads.

Reference: MIMIC-IV/MortModel/Paired_Cross_Attention/env_config.py:53 (ROUTES), :157-158
(route_prior_floor / route_prior_ceiling; the defaults always win).  The
reference's config plumbing (YAML / env / CLI overrides, seeding, checkpoint dirs) is out of scope.
"""
from types import SimpleNamespace

ROUTES = ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]
ROUTE_NAMES = list(ROUTES)
BLOCKS = {"uni": ["L", "N", "I"], "bi": ["LN", "NL", "LI", "IL", "NI", "IN"], "tri": ["LNI"]}

CFG = SimpleNamespace(
    d=256, dropout=0.0, verbose=False,
    capsule_pc_dim=32, capsule_mc_caps_dim=64, capsule_num_routing=3, capsule_act_type="EM",
    route_prior_floor=0.02, route_prior_ceiling=0.98, routing_coef_mode="none", routing_coef_eps=1e-6,
)
DEVICE = "cuda"
