"""Malformed calls of the drop-in surface ("Error convention"), shared by
oracle/gen_golden_errors.py (runs them against the UNMODIFIED reference modules and records what is raised) and
tests/test_error_parity.py (runs them against multimodalrouting_b200 on CPU tensors: validation must fire before any
device work, with the same exception type).

Every case is `name -> callable(ns)`; `ns` provides mult, proj, head, rh (the routing_and_heads module), torch and a dict
of well-formed inputs `ok` (B=2)."""
import torch

ROUTES = ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]
B = 2


def good_inputs():
    g = torch.Generator().manual_seed(31)
    return dict(x_l=torch.randn(B, 6, 256, generator=g), x_n=torch.randn(B, 4, 256, generator=g),
                x_i=torch.randn(B, 5, 256, generator=g), mL=torch.ones(B, 6), mN=torch.ones(B, 4), mI=torch.ones(B, 5),
                routes={r: torch.randn(B, 256, generator=g) for r in ROUTES}, pose=torch.randn(B, 10, 32, generator=g),
                act=torch.rand(B, 10, generator=g))


def _dict_call(ns, routes, **kw):
    return ns["rh"].forward_capsule_from_route_dict(routes, ns["proj"], ns["head"], **kw)


def _without(d, k):
    d = dict(d)
    d.pop(k)
    return d


class _WrongKeys(torch.nn.Module):
    def forward(self, x_l, x_n, x_i, mL=None, mN=None, mI=None):
        return {"L": x_l.mean(1), "N": x_n.mean(1)}


CASES = {
    # forward_capsule_from_route_dict (routing_and_heads.py:282-303, 316-327)
    "dict_missing_key": lambda ns: _dict_call(ns, _without(ns["ok"]["routes"], "LNI")),
    "dict_extra_key": lambda ns: _dict_call(ns, {**ns["ok"]["routes"], "XYZ": torch.zeros(B, 256)}),
    "dict_non_tensor": lambda ns: _dict_call(ns, {**ns["ok"]["routes"], "NL": [1.0, 2.0]}),
    "dict_rank1_route": lambda ns: _dict_call(ns, {**ns["ok"]["routes"], "IL": torch.zeros(256)}),
    "dict_rank3_route": lambda ns: _dict_call(ns, {**ns["ok"]["routes"], "IL": torch.zeros(B, 2, 256)}),
    "dict_route_mask_rank3": lambda ns: _dict_call(ns, ns["ok"]["routes"], route_mask=torch.ones(B, 10, 1)),
    # shapes that would otherwise reach the kernels as raw pointers: the reference fails in F.linear / the broadcasts
    "dict_route_wrong_width": lambda ns: _dict_call(ns, {**ns["ok"]["routes"], "IL": torch.zeros(B, 128)}),
    "dict_route_wrong_batch": lambda ns: _dict_call(ns, {**ns["ok"]["routes"], "IL": torch.zeros(B + 1, 256)}),
    "dict_route_mask_wrong_routes": lambda ns: _dict_call(ns, ns["ok"]["routes"], route_mask=torch.ones(B, 9)),
    "dict_route_mask_wrong_batch": lambda ns: _dict_call(ns, ns["ok"]["routes"], route_mask=torch.ones(B + 1, 10)),
    "dict_route_mask_1d_wrong": lambda ns: _dict_call(ns, ns["ok"]["routes"], route_mask=torch.ones(7)),
    "dict_override_wrong_routes": lambda ns: _dict_call(ns, ns["ok"]["routes"], acts_override=torch.rand(B, 9, 1)),
    "head_pose_wrong_width": lambda ns: ns["head"](torch.zeros(B, 10, 16), ns["ok"]["act"]),
    "head_pose_wrong_routes": lambda ns: ns["head"](torch.zeros(B, 9, 32), ns["ok"]["act"]),
    "head_act_wrong_routes": lambda ns: ns["head"](ns["ok"]["pose"], torch.rand(B, 9)),
    "head_act_wrong_batch": lambda ns: ns["head"](ns["ok"]["pose"], torch.rand(B + 1, 10)),
    "head_route_mask_wrong_routes": lambda ns: ns["head"](ns["ok"]["pose"], ns["ok"]["act"], route_mask=torch.ones(B, 9)),
    "proj_route_wrong_width": lambda ns: ns["proj"]({**ns["ok"]["routes"], "IL": torch.zeros(B, 128)}),
    "proj_route_wrong_batch": lambda ns: ns["proj"]({**ns["ok"]["routes"], "IL": torch.zeros(B + 1, 256)}),
    "proj_missing_key": lambda ns: ns["proj"](_without(ns["ok"]["routes"], "LNI")),
    # CapsuleMortalityHead.forward (routing_and_heads.py:200-222)
    "head_act_rank1": lambda ns: ns["head"](ns["ok"]["pose"], torch.rand(10)),
    "head_act_rank3_wide": lambda ns: ns["head"](ns["ok"]["pose"], torch.rand(B, 10, 2)),
    "head_route_mask_rank3": lambda ns: ns["head"](ns["ok"]["pose"], ns["ok"]["act"], route_mask=torch.ones(B, 10, 1)),
    # MULTModel.forward (mult_model.py:117-121)
    "mult_rank2_input": lambda ns: ns["mult"](ns["ok"]["x_l"][:, 0], ns["ok"]["x_n"], ns["ok"]["x_i"]),
    "mult_batch_mismatch": lambda ns: ns["mult"](ns["ok"]["x_l"], ns["ok"]["x_n"][:1], ns["ok"]["x_i"]),
    # empty batch / empty modality: the reference fails in MultiheadAttention's head reshape (RuntimeError)
    "mult_empty_batch": lambda ns: ns["mult"](torch.zeros(0, 6, 256), torch.zeros(0, 4, 256), torch.zeros(0, 5, 256)),
    "mult_empty_modality": lambda ns: ns["mult"](ns["ok"]["x_l"], torch.zeros(B, 0, 256), ns["ok"]["x_i"]),
    # masks whose shape does not match the token axis (multihead_attention.py:121-124 for key masks)
    "mult_key_mask_too_long": lambda ns: ns["mult"](ns["ok"]["x_l"], ns["ok"]["x_n"], ns["ok"]["x_i"],
                                                    mL=ns["ok"]["mL"], mN=torch.ones(B, 7), mI=ns["ok"]["mI"]),
    "mult_mask_wrong_batch": lambda ns: ns["mult"](ns["ok"]["x_l"], ns["ok"]["x_n"], ns["ok"]["x_i"],
                                                   mL=torch.ones(B + 1, 6), mN=ns["ok"]["mN"], mI=ns["ok"]["mI"]),
    # make_route_inputs_mult / forward_capsule_from_multmodel with a producer that returns the wrong key set
    "multmodel_wrong_keys": lambda ns: ns["rh"].forward_capsule_from_multmodel(
        _WrongKeys(), ns["ok"]["x_l"], ns["ok"]["x_n"], ns["ok"]["x_i"], ns["proj"], ns["head"]),
}
