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


@pytest.mark.parametrize("variant", ["mort", "pheno"])
def test_default_init_matches_reference(variant):
    """SURVEY.md section 8b "Init semantics": constructed under the same seed, the drop-in modules consume the RNG in the
    reference's order and hold the same tensors (fixture: oracle/gen_golden_init.py run on the unmodified reference)."""
    import torch
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "init_checksums.json")))
    c = g["configs"][variant]
    import multimodalrouting_b200 as mmr
    _, rh = _mods(variant)
    torch.manual_seed(g["seed"])
    mult = mmr.MULTModel(256, c["orig_d_n"], 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=c["K"])
    seen = set()
    for tag, m in (("mult", mult), ("proj", proj), ("head", head)):
        for k, t in m.state_dict().items():
            ref = g[variant][f"{tag}.{k}"]
            d = t.detach().double()
            assert list(t.shape) == ref[0], k
            for got, want in ((float(d.sum()), ref[1]), (float(d.norm()), ref[2]), (float(d.flatten()[0]), ref[3])):
                assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), f"{tag}.{k}"
            seen.add(f"{tag}.{k}")
    assert seen == set(g[variant]) and len(seen) > 300
    # the quirks of the reference's init the tests elsewhere have to perturb: zero decision embedding / biases
    assert float(head.embedding.detach().abs().max()) == 0.0 and float(head.bias.detach().abs().max()) == 0.0


@pytest.mark.parametrize("variant", ["mort", "pheno"])
def test_small_helpers_behave_like_the_reference(variant):
    """route_given_pheno (plain tensor code in the drop-in too) and make_route_inputs_mult's call convention, against outputs
    recorded from the unmodified reference (oracle/gen_golden_surface.py)."""
    import torch
    _, rh = _mods(variant)
    g = GOLD[variant]["route_given_pheno"]
    q, m1, m2 = torch.tensor(g["q"]), torch.tensor(g["m1"]), torch.tensor(g["m2"])
    assert torch.equal(rh.route_given_pheno(q), torch.tensor(g["none"]))
    assert torch.equal(rh.route_given_pheno(q, m1), torch.tensor(g["mask1d"]))
    assert torch.equal(rh.route_given_pheno(q, route_mask=m2), torch.tensor(g["mask2d"]))

    class Stub(torch.nn.Module):
        def forward(self, *a, **kw):
            self.seen = {"n_positional": len(a), "kwargs": sorted(kw), "mask_is_none": sorted(k for k, v in kw.items() if v is None)}
            return {r: torch.zeros(1, 4) for r in rh.ROUTES}
    stub = Stub()
    z = {"L": {"seq": torch.zeros(1, 2, 4), "mask": torch.ones(1, 2)}, "N": {"seq": torch.zeros(1, 3, 4)},
         "I": {"seq": torch.zeros(1, 2, 4), "mask": None}}
    out = rh.make_route_inputs_mult(z, stub)
    ref = GOLD[variant]["make_route_inputs_mult"]
    assert stub.seen == ref["call"] and sorted(out) == ref["returns_keys"]
