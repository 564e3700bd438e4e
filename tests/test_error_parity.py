"""CPU: the drop-in surface raises what the reference raises ("Error convention").

tests/golden/error_cases.json was produced by oracle/gen_golden_errors.py, which runs the malformed calls of
tests/error_cases.py against the unmodified reference modules.  Here the same calls go to multimodalrouting_b200 with CPU
tensors: every check must fire in the host-side validation, i.e. before the op that would refuse CPU tensors."""
import json
import os

import pytest
import torch

import error_cases as ec
from helpers import ROOT

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "error_cases.json")))
EXC = {"RuntimeError": RuntimeError, "ValueError": ValueError, "TypeError": TypeError, "AssertionError": AssertionError,
       "KeyError": KeyError}


def _modules(variant):
    import multimodalrouting_b200 as mmr
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    mult = mmr.MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    return rh, mult, rh.RoutePrimaryProjector(256, 32), rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM",
                                                                               num_classes=2 if variant == "mort" else 25)


@pytest.mark.parametrize("variant", ["mort", "pheno"])
@pytest.mark.parametrize("case", sorted(ec.CASES))
def test_same_exception_type_as_the_reference(variant, case):
    gold = GOLD[variant][case]
    rh, mult, proj, head = _modules(variant)
    ns = dict(mult=mult, proj=proj, head=head, rh=rh, torch=torch, ok=ec.good_inputs())
    if gold["raises"] == "ok":       # the reference accepts the call: here validation passes and the CPU tensors are refused
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ec.CASES[case](ns)
        return
    with pytest.raises(EXC[gold["raises"]]) as ei:
        ec.CASES[case](ns)
    assert "no CPU fallback" not in str(ei.value), "the malformed call reached the device op instead of the validation"
    if gold["raises"] != "AssertionError" and not any(t in case for t in ("mask", "empty", "wrong", "proj_")):
        # messages of the explicit checks are the reference's own text (prefix)
        assert str(ei.value)[:40] == gold["msg"][:40]


def test_every_case_is_pinned_for_both_variants():
    for v in ("mort", "pheno"):
        assert set(GOLD[v]) == set(ec.CASES)
        assert all(r["raises"] in EXC for r in GOLD[v].values())     # the reference rejects every one of these calls
