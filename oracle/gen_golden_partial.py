"""Generate tests/golden/partial_fusion.pt by running the UNMODIFIED reference modules (TEST INFRASTRUCTURE).

    python oracle/gen_golden_partial.py      (build container only: needs /root/reference)

CrossAttentionFusion (mean and first pooling) and TriTokenAttentionFusion of
/root/reference/MIMIC-IV/PhenoModel/Partial/Cross_Attention/routing_and_heads.py, fp32 on the CPU, seeded weights / inputs from
multimodalrouting_b200/synth.py (make_fusion_state / make_fusion_inputs); stores outputs, input gradients and every parameter
gradient for the loss sum(out * probe).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference/MIMIC-IV/PhenoModel/Partial/Cross_Attention"

CASES = {
    "cross_mean": dict(kind="cross", pool="mean", B=5, TA="L", TB="N", TL=13, TN=7, TI=9, seed=11),
    "cross_first": dict(kind="cross", pool="first", B=4, TA="I", TB="L", TL=20, TN=5, TI=11, seed=12),
    "cross_long": dict(kind="cross", pool="mean", B=3, TA="L", TB="I", TL=70, TN=4, TI=90, seed=13),
    "tri": dict(kind="tri", B=5, TL=13, TN=7, TI=9, seed=14),
    "tri_long": dict(kind="tri", B=3, TL=48, TN=16, TI=49, seed=15),
}


def main():
    from multimodalrouting_b200 import synth
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import routing_and_heads as rh          # the reference module, unmodified
    gold = {}
    for name, c in CASES.items():
        sd = synth.make_fusion_state(c["kind"], c["seed"])
        inp = synth.make_fusion_inputs(c["B"], c["TL"], c["TN"], c["TI"], c["seed"] + 1000)
        g = torch.Generator().manual_seed(c["seed"] + 2000)
        probe = torch.randn(c["B"], 256, generator=g)
        if c["kind"] == "cross":
            mod = rh.CrossAttentionFusion(256, 8, 0.0, c["pool"])
            mod.load_state_dict(sd, strict=True)
            A = inp[c["TA"]].clone().requires_grad_(True)
            Bs = inp[c["TB"]].clone().requires_grad_(True)
            out = mod(A, inp["m" + c["TA"]], Bs, inp["m" + c["TB"]])
            ins = {"A": A, "B": Bs}
        else:
            mod = rh.TriTokenAttentionFusion(256, 8, 0.0)
            mod.load_state_dict(sd, strict=True)
            xs = {k: inp[k].clone().requires_grad_(True) for k in ("L", "N", "I")}
            out = mod(xs["L"], inp["mL"], xs["N"], inp["mN"], xs["I"], inp["mI"])
            ins = xs
        (out * probe).sum().backward()
        gold[name] = {"case": c, "out": out.detach().clone(),
                      "d_in": {k: v.grad.detach().clone() for k, v in ins.items()},
                      "d_param": {k: p.grad.detach().clone() for k, p in mod.named_parameters()}}
        print(name, tuple(out.shape), float(out.abs().max()))
    # large gradient tensors keep a (norm, projection on a seeded random direction) checksum instead of the values
    for name in gold:
        gq = torch.Generator().manual_seed(99)
        for grp in ("d_in", "d_param"):
            for k, v in list(gold[name][grp].items()):
                if v.numel() <= 4096:
                    continue
                r = torch.randn(v.shape, generator=gq)
                gold[name][grp][k] = {"norm": float(v.norm()), "proj": float((v * r).sum()), "shape": tuple(v.shape)}
    path = os.path.join(ROOT, "tests", "golden", "partial_fusion.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
