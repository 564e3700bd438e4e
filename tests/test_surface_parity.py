"""CPU: the drop-in modules expose the names the reference drivers import, with the reference's signatures
("Python surface (must match)").  tests/golden/surface.json comes from the unmodified reference
(oracle/gen_golden_surface.py: ast over main.py's imports, inspect.signature over the imported modules)."""
import inspect
import json
import os

import pytest

from helpers import ROOT

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "surface.json")))


def _mods(variant):
    import multimodalrouting_b200 as mmr
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    return mmr.mult_model, rh


def _sig(fn):
    return [[p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
            for p in inspect.signature(fn).parameters.values()]


@pytest.mark.parametrize("variant", ["mort", "pheno"])
def test_driver_imports_resolve(variant):
    mult_model, rh = _mods(variant)
    imp = GOLD[variant]["driver_imports"]
    assert "MULTModel" in imp["mult_model"] and len(imp["routing_and_heads"]) >= 6
    for name in imp["mult_model"]:
        assert hasattr(mult_model, name), name
    for name in imp["routing_and_heads"]:
        assert hasattr(rh, name), name


@pytest.mark.parametrize("variant", ["mort", "pheno"])
@pytest.mark.parametrize("name", sorted(GOLD["mort"]["signatures"]))
def test_signature_matches_reference(variant, name):
    mult_model, rh = _mods(variant)
    obj = mult_model if name.startswith("MULTModel") else rh
    for part in name.split("."):
        obj = getattr(obj, part)
    ours, ref = _sig(obj), GOLD[variant]["signatures"][name]
    assert [p[0] for p in ours] == [p[0] for p in ref], f"{name}: parameter names / order"
    assert [p[1] for p in ours] == [p[1] for p in ref], f"{name}: parameter kinds"
    assert [p[2] for p in ours] == [p[2] for p in ref], f"{name}: defaults"
