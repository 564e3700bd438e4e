"""Generate tests/golden/init_checksums.json: per-parameter checksums of the reference modules constructed under a fixed
seed (TEST INFRASTRUCTURE; build container only, needs /root/reference).

    python oracle/gen_golden_init.py

SURVEY.md section 8b "Init semantics": xavier-uniform attention / FFN weights, zero biases, default Linear / Conv1d /
LayerNorm init elsewhere, capsule.w = sqrt(K / (32 * 10)) * randn, zero decision embedding.  Constructing the drop-in
modules under the same seed must consume the RNG in the same order and give the same tensors
(tests/test_surface_parity.py::test_default_init_matches_reference)."""
import contextlib
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/MIMIC-IV"
SEED = 1234
CONFIGS = {"mort": dict(orig_d_n=768, K=2), "pheno": dict(orig_d_n=256, K=25)}


def checksums(mods):
    out = {}
    for tag, m in mods:
        for k, t in m.state_dict().items():
            d = t.detach().double()
            out[f"{tag}.{k}"] = [list(t.shape), float(d.sum()), float(d.norm()), float(d.flatten()[0])]
    return out


def run_variant(variant):
    sys.path.insert(0, f"{REF}/PhenoModel/Paired_Cross_Attention")
    if variant == "mort":
        sys.path.insert(0, f"{REF}/MortModel/Paired_Cross_Attention")
    with contextlib.redirect_stdout(io.StringIO()):
        import mult_model
        import routing_and_heads as rh
    import torch
    c = CONFIGS[variant]
    torch.manual_seed(SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        mult = mult_model.MULTModel(256, c["orig_d_n"], 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
        proj = rh.RoutePrimaryProjector(256, 32)
        head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=c["K"])
    print(json.dumps(checksums((("mult", mult), ("proj", proj), ("head", head)))))


def main():
    res = {"seed": SEED, "configs": CONFIGS}
    for v in CONFIGS:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", v], capture_output=True, text=True, check=True)
        res[v] = json.loads(p.stdout.strip().splitlines()[-1])
    json.dump(res, open(os.path.join(ROOT, "tests", "golden", "init_checksums.json"), "w"), indent=0, sort_keys=True)
    print({v: len(res[v]) for v in CONFIGS})


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--variant":
        run_variant(sys.argv[2])
    else:
        main()
