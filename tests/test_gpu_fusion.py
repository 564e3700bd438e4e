"""-m gpu: end-to-end parity of the drop-in modules (route fusion + routing, fwd + bwd) against the
golden vectors produced by the unmodified reference (fp32) and against the oracle under autocast
(bf16).  Tolerances are the ones BASELINE.json's north_star states: fp32 1e-4 relative, bf16 2e-2."""
import os

import pytest
import torch

from gpu_common import run_case
from helpers import GOLDEN_CASES, check_grad_checksums, load_golden, max_rel, r_grad_probe, rebuild_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp32_matches_reference_golden(name):
    gold = load_golden(name)
    c = gold["case"]
    sdm, sdp, sdh, inp = rebuild_case(c)
    out = run_case(c, sdm, sdp, sdh, inp, r_probe=r_grad_probe(c, gold["R"].shape))
    assert max_rel(out["routes"], gold["routes"]) < 1e-4, "route embeddings"
    assert max_rel(out["logits"], gold["logits"]) < 1e-4, "logits"
    assert max_rel(out["alpha"], gold["alpha"]) < 1e-4, "alpha"
    assert max_rel(out["R"], gold["R"]) < 1e-4, "R"
    if c["variant"] == "mort":
        assert torch.equal((out["logits"][:, 1] > out["logits"][:, 0]).cpu(), gold["logits"][:, 1] > gold["logits"][:, 0])
    else:
        assert torch.equal((out["logits"] > 0).cpu(), gold["logits"] > 0)
    assert abs(out["loss"] - gold["loss"]) < 1e-5
    none = sorted(k for k, g in out["grads"].items() if g is None)
    assert none == gold["grad_none"], (none, gold["grad_none"])
    for k, g in gold["grad_full"].items():
        assert max_rel(out["grads"][k], g) < 5e-4, f"grad {k}"
    check_grad_checksums(out["grads"], gold["grad_checksum"], 5e-4, name)


def _bf16_case(name, engine):
    os.environ["MMR_B200_GEMM"] = engine
    try:
        gold = load_golden(name)
        c = gold["case"]
        sdm, sdp, sdh, inp = rebuild_case(c)
        out = run_case(c, sdm, sdp, sdh, inp, autocast=True, r_probe=r_grad_probe(c, gold["R"].shape))
    finally:
        os.environ.pop("MMR_B200_GEMM", None)
    # bf16 budget (north_star): 2e-2 on logits and route weights alpha / R, identical argmax
    assert max_rel(out["routes"], gold["routes"]) < 2e-2, "route embeddings"
    assert max_rel(out["logits"], gold["logits"]) < 2e-2, "logits"
    assert max_rel(out["alpha"], gold["alpha"]) < 2e-2, "alpha"
    assert max_rel(out["R"], gold["R"]) < 2e-2, "R"
    for k, g in gold["grad_full"].items():
        assert max_rel(out["grads"][k], g) < 6e-2, f"grad {k}"
    return out, gold


@pytest.mark.parametrize("name", ["mort_cfg1", "pheno_sharp4", "mort_missing", "pheno_odd"])
def test_bf16_simt_engine(name):
    _bf16_case(name, "simt")
