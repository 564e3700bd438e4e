"""Generate tests/golden/surface.json: the Python surface of the reference's hot-path modules (TEST INFRASTRUCTURE; build
container only, needs /root/reference).

    python oracle/gen_golden_surface.py

Per variant (mort / pheno): (a) the names the driver imports from `mult_model` and `routing_and_heads` (main.py:30-45,
parsed with ast -- main.py itself cannot be imported), (b) the signature (parameter names, kinds, defaults) of every
constructor / forward / function of SURVEY.md section 8b, taken with inspect from the imported reference modules.
tests/test_surface_parity.py holds the drop-in modules to both."""
import ast
import contextlib
import inspect
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/MIMIC-IV"
DIRS = {"mort": f"{REF}/MortModel/Paired_Cross_Attention", "pheno": f"{REF}/PhenoModel/Paired_Cross_Attention"}
CALLABLES = ["MULTModel.__init__", "MULTModel.forward", "RoutePrimaryProjector.__init__", "RoutePrimaryProjector.forward",
             "RouteDimAdapter.__init__", "RouteDimAdapter.forward", "CapsuleMortalityHead.__init__",
             "CapsuleMortalityHead.forward", "forward_capsule_from_route_dict", "forward_capsule_from_multmodel",
             "make_route_inputs_mult", "route_given_pheno"]


def sig_of(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        out.append([p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


def resolve(ns, dotted):
    obj = ns
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


def run_variant(variant):
    sys.path.insert(0, DIRS["pheno"])
    if variant == "mort":
        sys.path.insert(0, DIRS["mort"])
    with contextlib.redirect_stdout(io.StringIO()):
        import mult_model
        import routing_and_heads as rh
    imports = {}
    for node in ast.parse(open(os.path.join(DIRS[variant], "main.py")).read()).body:
        if isinstance(node, ast.ImportFrom) and node.module in ("mult_model", "routing_and_heads"):
            imports.setdefault(node.module, []).extend(a.name for a in node.names)
    sigs = {}
    for name in CALLABLES:
        mod = mult_model if name.startswith("MULTModel") else rh
        sigs[name] = sig_of(resolve(mod, name))
    # small helpers that stay plain tensor / dict code in the drop-in: pin their behaviour too
    import torch
    g = torch.Generator().manual_seed(5)
    q = torch.rand(2, 10, 3, generator=g)
    m2 = (torch.rand(2, 10, generator=g) < 0.7).float()
    m1 = torch.tensor([1., 0, 1, 1, 0, 1, 1, 1, 0, 1])
    rgp = {"q": q.tolist(), "m1": m1.tolist(), "m2": m2.tolist(),
           "none": rh.route_given_pheno(q).tolist(), "mask1d": rh.route_given_pheno(q, m1).tolist(),
           "mask2d": rh.route_given_pheno(q, route_mask=m2).tolist()}

    class Stub(torch.nn.Module):
        def forward(self, *a, **kw):
            self.seen = {"n_positional": len(a), "kwargs": sorted(kw), "mask_is_none": sorted(k for k, v in kw.items() if v is None)}
            return {r: torch.zeros(1, 4) for r in rh.ROUTES}
    stub = Stub()
    z = {"L": {"seq": torch.zeros(1, 2, 4), "mask": torch.ones(1, 2)}, "N": {"seq": torch.zeros(1, 3, 4)},
         "I": {"seq": torch.zeros(1, 2, 4), "mask": None}}
    out = rh.make_route_inputs_mult(z, stub)
    mri = {"call": stub.seen, "returns_keys": sorted(out)}
    print(json.dumps({"driver_imports": imports, "signatures": sigs, "route_given_pheno": rgp, "make_route_inputs_mult": mri}))


def main():
    res = {}
    for v in ("mort", "pheno"):
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", v], capture_output=True, text=True, check=True)
        res[v] = json.loads(p.stdout.strip().splitlines()[-1])
    json.dump(res, open(os.path.join(ROOT, "tests", "golden", "surface.json"), "w"), indent=1, sort_keys=True)
    print({v: {k: len(x) for k, x in r.items()} for v, r in res.items()})
    print(res["mort"]["driver_imports"])


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--variant":
        run_variant(sys.argv[2])
    else:
        main()
