"""Generate tests/golden/*.pt by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/<case>.pt

Each case rebuilds seeded weights/inputs with ``oracle/synth.py``, loads them into
the reference nn.Modules (strict ``load_state_dict``), runs
``forward_capsule_from_multmodel`` + loss + backward on CPU fp32, and stores the
outputs, selected full gradients, and a (norm, projection) checksum for every
parameter/input gradient.  Mort and Pheno variants are imported in separate
subprocesses because both use the bare module name ``routing_and_heads``
(SURVEY.md section 8c).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import subprocess
import sys
import zlib

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/MIMIC-IV"
M_DIR = f"{REF}/MortModel/Paired_Cross_Attention"
P_DIR = f"{REF}/PhenoModel/Paired_Cross_Attention"

# name -> kwargs.  Kept small: the whole golden set is a few MB.
CASES = {
    # BASELINE config 1 shape (Mort, orig_d_n=768, B small)
    "mort_cfg1": dict(variant="mort", K=2, orig_d_n=768, B=4, seed=101, sharp=1.0, temp=1.0,
                      detach=False, missing=False, mask_mode="full"),
    # Pheno 25-label, sharpened routing regime
    "pheno_sharp4": dict(variant="pheno", K=25, orig_d_n=256, B=4, seed=202, sharp=4.0, temp=1.0,
                         detach=False, missing=False, mask_mode="full"),
    # Pheno warm-up regime: temperature 2, detached priors (P/main.py:2742-2743)
    "pheno_warm": dict(variant="pheno", K=25, orig_d_n=256, B=3, seed=303, sharp=2.0, temp=2.0,
                       detach=True, missing=False, mask_mode="full"),
    # missing-modality (config 4), both variants
    "mort_missing": dict(variant="mort", K=2, orig_d_n=256, B=8, seed=404, sharp=4.0, temp=1.2,
                         detach=False, missing=True, mask_mode="full"),
    "pheno_missing": dict(variant="pheno", K=25, orig_d_n=256, B=8, seed=505, sharp=4.0, temp=2.0,
                          detach=False, missing=True, mask_mode="full"),
    # no masks at all (mL=mN=mI=None, route_mask=None): clamp-all branch
    "mort_nomask": dict(variant="mort", K=2, orig_d_n=256, B=2, seed=606, sharp=2.0, temp=1.5,
                        detach=False, missing=False, mask_mode="none"),
    # whole-batch route dropout with a 1-D route mask (M/main.py:3027-3033)
    "pheno_rm1d": dict(variant="pheno", K=25, orig_d_n=256, B=3, seed=707, sharp=4.0, temp=1.0,
                       detach=False, missing=False, mask_mode="rm1d"),
    # short odd shapes + 3 labels (INSPECT-like head), exercises ragged tiles
    "pheno_odd": dict(variant="pheno", K=3, orig_d_n=256, B=5, seed=808, sharp=4.0, temp=1.0,
                      detach=False, missing=True, mask_mode="full", TL=21, TN=5, TI=9),
    # acts_override: externally supplied route priors replace the projector's sigmoid activations (routing_and_heads.py:314)
    "pheno_override": dict(variant="pheno", K=25, orig_d_n=256, B=4, seed=1111, sharp=3.0, temp=1.2,
                           detach=False, missing=True, mask_mode="full", override=True, long=True),
    # two routing iterations instead of the drivers' three (CapsuleMortalityHead(num_routing=2), routing_and_heads.py:233-247)
    "mort_iter2": dict(variant="mort", K=2, orig_d_n=256, B=4, seed=1212, sharp=4.0, temp=1.0,
                       detach=False, missing=True, mask_mode="full", iters=2, long=True),
    # two cross-modal layers instead of four (MULTModel(layers=2), mult_model.py:59-81)
    "pheno_layers2": dict(variant="pheno", K=25, orig_d_n=256, B=3, seed=1313, sharp=3.0, temp=1.0,
                          detach=False, missing=True, mask_mode="full", layers=2, long=True),
    # Conv1d(k=1) input projections on all three modalities (orig_d_l=64, orig_d_n=768, orig_d_i=32; mult_model.py:30-32,134-136)
    "mort_proj_all": dict(variant="mort", K=2, orig_d_n=768, B=3, seed=1414, sharp=3.0, temp=1.0,
                          detach=False, missing=True, mask_mode="full", orig_d_l=64, orig_d_i=32, long=True),
    # unsharpened missing-modality pair (sharp=1: routing is well conditioned for every patient, so the bf16 bar of the GPU
    # tests applies to all of them) and the gradient through an externally supplied acts_override (Mort variant)
    "pheno_missing1": dict(variant="pheno", K=25, orig_d_n=256, B=8, seed=1515, sharp=1.0, temp=1.0,
                           detach=False, missing=True, mask_mode="full", long=True),
    "mort_missing1": dict(variant="mort", K=2, orig_d_n=256, B=8, seed=1616, sharp=1.0, temp=1.0,
                          detach=False, missing=True, mask_mode="full", long=True),
    "pheno_override_grad": dict(variant="pheno", K=25, orig_d_n=256, B=4, seed=1818, sharp=2.0, temp=1.5,
                                detach=False, missing=True, mask_mode="full", override=True, override_grad=True, long=True),
    "mort_override": dict(variant="mort", K=2, orig_d_n=256, B=4, seed=1717, sharp=3.0, temp=1.5,
                          detach=False, missing=True, mask_mode="full", override=True, override_grad=True, long=True),
    # long sequences ("long": only the CPU oracle test iterates them; the GPU tests reach these token counts through the
    # oracle): PhenoModel's own default structured_seq_len=256 (P/env_config.py:96) and the INSPECT token counts of
    # BASELINE configs[4] with a 3-label head
    "pheno_tl256": dict(variant="pheno", K=25, orig_d_n=256, B=2, seed=909, sharp=2.0, temp=1.0,
                        detach=False, missing=True, mask_mode="full", TL=256, TN=16, TI=49, long=True),
    "pheno_inspect": dict(variant="pheno", K=3, orig_d_n=256, B=2, seed=1010, sharp=8.0, temp=1.0,
                          detach=False, missing=False, mask_mode="full", TL=512, TN=128, TI=196, long=True),
}

FULL_GRADS = ["embedding", "bias", "pose_to_mc.weight", "x_n",
              "trans_l_with_n.layers.0.self_attn.in_proj_bias",
              "trans_l_with_n.layers.0.layer_norms.0.weight",
              "trans_l_with_n.layers.0.layer_norms.0.bias",
              "trans_n_with_i.layers.3.layer_norms.1.weight",
              "trans_i_with_l.layers.2.fc2.bias",
              "trans_i_with_n.layer_norm.weight", "trans_l.layer_norm.bias",
              "proj_pair_li.bias", "final_lni.bias", "proj.LNI.bias", "proj.NL.weight"]


def proj_vec(name: str, n: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randn(n, generator=g)


def checksum(name: str, t: torch.Tensor):
    f = t.detach().double().flatten()
    return [float(f.norm()), float(f @ proj_vec(name, f.numel()).double()), float(f.sum())]


def build_case_inputs(c):
    sys.path.insert(0, ROOT)
    from oracle import synth
    sdm, sdp, sdh = synth.make_state(K=c["K"], orig_d_n=c["orig_d_n"], seed=c["seed"], sharp=c["sharp"],
                                     layers=c.get("layers", 4), orig_d_l=c.get("orig_d_l", 0), orig_d_i=c.get("orig_d_i", 0))
    inp = synth.make_inputs(B=c["B"], d_n=c["orig_d_n"], K=c["K"], seed=c["seed"] + 1, missing=c["missing"],
                            TL=c.get("TL", 48), TN=c.get("TN", 16), TI=c.get("TI", 49),
                            d_l=c.get("orig_d_l", 0) or 256, d_i=c.get("orig_d_i", 0) or 256)
    if c["mask_mode"] == "none":
        inp["mL"] = inp["mN"] = inp["mI"] = None
        inp["route_mask"] = None
    elif c["mask_mode"] == "rm1d":
        rm = torch.ones(10)
        rm[[1, 4, 8]] = 0.0
        inp["route_mask"] = rm
    if c.get("override"):
        inp["acts_override"] = torch.rand(c["B"], 10, 1, generator=torch.Generator().manual_seed(c["seed"] + 3))
    return sdm, sdp, sdh, inp


def run_variant(variant: str):
    """Executed in a subprocess: import the reference for one variant and run its cases."""
    sys.path.insert(0, P_DIR)
    if variant == "mort":
        sys.path.insert(0, M_DIR)
    with contextlib.redirect_stdout(io.StringIO()):
        import mult_model            # noqa
        import routing_and_heads as rh   # noqa
    sys.path.insert(0, ROOT)
    from oracle import synth
    torch.manual_seed(0)
    torch.set_num_threads(8)
    only = [n for n in os.environ.get("MMR_GOLDEN_ONLY", "").split(",") if n]   # regenerate a subset, leave the rest untouched
    for name, c in CASES.items():
        if c["variant"] != variant or (only and name not in only):
            continue
        sdm, sdp, sdh, inp = build_case_inputs(c)
        with contextlib.redirect_stdout(io.StringIO()):
            mult = mult_model.MULTModel(c.get("orig_d_l", 0) or 256, c["orig_d_n"], c.get("orig_d_i", 0) or 256, 256, 256, 256, True, True, True,
                                        8, c.get("layers", 4), 0, 0., 0., 0., 0., 0., 0., 0., False)
            proj = rh.RoutePrimaryProjector(256, 32)
            head = rh.CapsuleMortalityHead(32, 64, c.get("iters", 3), 0.0, "EM", num_classes=c["K"])
        mult.load_state_dict(sdm, strict=True)
        proj.load_state_dict(sdp, strict=True)
        head.load_state_dict(sdh, strict=True)
        xs = {k: inp[k].clone().requires_grad_(True) for k in ("x_l", "x_n", "x_i")}
        if c.get("override_grad"):
            inp["acts_override"] = inp["acts_override"].clone().requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            logits, alpha, routes, R = rh.forward_capsule_from_multmodel(
                mult, xs["x_l"], xs["x_n"], xs["x_i"], proj, head,
                mL=inp["mL"], mN=inp["mN"], mI=inp["mI"],
                route_adapter=rh.RouteDimAdapter(256, 256, 256, 256),
                route_mask=inp["route_mask"], act_temperature=c["temp"], detach_priors=c["detach"],
                acts_override=inp.get("acts_override"))
        loss = synth.loss_fn(logits, inp["y"], variant)
        # also push a gradient through R so the routing-coefficient path is pinned
        gR = torch.randn(R.shape, generator=torch.Generator().manual_seed(c["seed"] + 7))
        total = loss + 0.05 * (R * gR).sum()
        total.backward()
        grads = {}
        for mod in (mult, proj, head):
            for n, p in mod.named_parameters():
                grads[n] = p.grad
        for k, v in xs.items():
            grads[k] = v.grad
        if c.get("override_grad"):
            grads["acts_override"] = inp["acts_override"].grad
        out = {
            "case": dict(c), "logits": logits.detach(), "alpha": alpha.detach(), "R": R.detach(),
            "routes": torch.stack([routes[r].detach() for r in synth.ROUTES], dim=1),
            "loss": float(loss), "total": float(total),
            "grad_none": sorted(n for n, g in grads.items() if g is None),
            "grad_checksum": {n: checksum(n, g) for n, g in grads.items() if g is not None},
            "grad_full": {n: grads[n].detach().clone() for n in FULL_GRADS + ["acts_override"] if grads.get(n) is not None},
        }
        torch.save(out, os.path.join(GOLD, name + ".pt"))
        print(f"[golden] {name}: loss={float(loss):.6f} |logits|max={float(logits.abs().max()):.4f} "
              f"R range=[{float(R.min()):.4f},{float(R.max()):.4f}]", file=sys.stderr)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1:
        run_variant(sys.argv[1])
        return
    for v in ("mort", "pheno"):
        subprocess.run([sys.executable, os.path.abspath(__file__), v], check=True)
    meta = {"generator": "oracle/gen_golden.py", "reference": "AI-for-Health-Data/MultimodalRouting",
            "torch": torch.__version__, "cases": sorted(CASES)}
    with open(os.path.join(GOLD, "META.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
