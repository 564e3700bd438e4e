"""Synthetic weights / inputs for the oracle-side tests and golden generators (TEST INFRASTRUCTURE).

The generator itself lives in the package (multimodalrouting_b200/synth.py) because bench.py's product arm needs the same
workload and must not import anything from oracle/; this module only re-exports it, so fixtures stay bit-identical."""
from multimodalrouting_b200.synth import *  # noqa: F401,F403
from multimodalrouting_b200.synth import _clamp_norm, _fill  # noqa: F401
