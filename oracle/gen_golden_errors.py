"""Generate tests/golden/error_cases.json by running the malformed calls of tests/error_cases.py against the UNMODIFIED
reference modules (TEST INFRASTRUCTURE; build container only, needs /root/reference):

    python oracle/gen_golden_errors.py

Records, per variant (mort / pheno) and case, the exception type the reference raises (or "ok" when it accepts the
call) and the first 120 characters of the message.  tests/test_error_parity.py holds the drop-in modules to the types."""
import contextlib
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/MIMIC-IV"
M_DIR = f"{REF}/MortModel/Paired_Cross_Attention"
P_DIR = f"{REF}/PhenoModel/Paired_Cross_Attention"


def run_variant(variant):
    sys.path.insert(0, P_DIR)
    if variant == "mort":
        sys.path.insert(0, M_DIR)
    with contextlib.redirect_stdout(io.StringIO()):
        import mult_model
        import routing_and_heads as rh
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import error_cases as ec
    with contextlib.redirect_stdout(io.StringIO()):
        mult = mult_model.MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
        proj = rh.RoutePrimaryProjector(256, 32)
        head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=2 if variant == "mort" else 25)
    out = {}
    for name, fn in ec.CASES.items():
        ns = dict(mult=mult, proj=proj, head=head, rh=rh, torch=torch, ok=ec.good_inputs())
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                fn(ns)
            out[name] = {"raises": "ok", "msg": ""}
        except Exception as e:                       # noqa: BLE001 -- the point is to record whatever is raised
            out[name] = {"raises": type(e).__name__, "msg": str(e)[:120]}
    print(json.dumps(out))


def main():
    res = {}
    for v in ("mort", "pheno"):
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", v], capture_output=True, text=True, check=True)
        res[v] = json.loads(p.stdout.strip().splitlines()[-1])
    path = os.path.join(ROOT, "tests", "golden", "error_cases.json")
    json.dump(res, open(path, "w"), indent=1, sort_keys=True)
    for v, cases in res.items():
        for k, r in cases.items():
            print(f"{v:5s} {k:28s} {r['raises']:16s} {r['msg'][:70]}")


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--variant":
        run_variant(sys.argv[2])
    else:
        main()
