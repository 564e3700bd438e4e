"""The seven attention-fusion route constructors of the Partial/ variant (make_route_inputs) at the MIMIC-IV token counts:
forward + backward per call, CUDA events -- multimodalrouting_b200/partial_fusion.py against the reference's eager path on the
same GPU (the oracle restatement of the reference modules on CUDA under torch.autocast(bfloat16): test infrastructure used as a
reported baseline, like bench.py's reference_gpu_eager).
    python tools/bench_partial_fusion.py [--B 512 --iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalrouting_b200 import partial_fusion as pf, synth  # noqa: E402
from oracle import partial_oracle as po  # noqa: E402

PAIRS = {"LN": ("L", "N"), "NL": ("N", "L"), "LI": ("L", "I"), "IL": ("I", "L"), "NI": ("N", "I"), "IN": ("I", "N")}


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    inp = synth.make_fusion_inputs(a.B, 48, 16, 49, 7)
    x = {k: inp[k].cuda().requires_grad_(True) for k in ("L", "N", "I")}
    m = {k: inp["m" + k].cuda() for k in ("L", "N", "I")}
    sds = {k: synth.make_fusion_state("cross", 100 + i) for i, k in enumerate(PAIRS)}
    sds["LNI"] = synth.make_fusion_state("tri", 200)
    fus = pf.build_fusions(256, device="cuda")
    for k in fus:
        fus[k].load_state_dict(sds[k])
    z = {k: {"seq": x[k], "mask": m[k], "pool": pf.masked_mean(x[k].detach(), m[k])} for k in ("L", "N", "I")}
    sdc = {k: {n: t.cuda().requires_grad_(True) for n, t in sd.items()} for k, sd in sds.items()}

    def ours():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            r = pf.make_route_inputs(z, fus)
        sum(v.float().sum() for k, v in r.items() if k in fus).backward()

    def ref():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = [po.cross_attention_fusion(sdc[k], x[qa], m[qa], x[kb], m[kb]) for k, (qa, kb) in PAIRS.items()]
            outs.append(po.tri_token_fusion(sdc["LNI"], x["L"], m["L"], x["N"], m["N"], x["I"], m["I"]))
        sum(o.float().sum() for o in outs).backward()

    t_ours, t_ref = timed(ours, a.iters), timed(ref, a.iters)
    # the same two steps captured into CUDA graphs: device time without the host's launch / dispatch cost
    from multimodalrouting_b200.graphs import GraphedStep

    def nograd(fn):
        def run():
            for k in x:
                x[k].grad = None
            for f in fus.values():
                for p in f.parameters():
                    p.grad = None
            for sd in sdc.values():
                for t in sd.values():
                    t.grad = None
            fn()
            return x["L"].grad
        return run
    g_ours, g_ref = GraphedStep(nograd(ours)), GraphedStep(nograd(ref))
    tg_ours, tg_ref = timed(g_ours, a.iters), timed(g_ref, a.iters)
    print(json.dumps({"ms_ours_graph": tg_ours, "ms_reference_graph": tg_ref, "speedup_graph": tg_ref / tg_ours,"what": "Partial-variant route constructors (6 CrossAttentionFusion + TriTokenAttentionFusion), fwd+bwd, bf16 autocast, "
                              "eagerly issued", "B": a.B, "tokens": [48, 16, 49], "ms_ours": t_ours, "ms_reference_eager_gpu": t_ref,
                      "speedup": t_ref / t_ours, "patients_per_s_ours": a.B / t_ours * 1e3}))


if __name__ == "__main__":
    main()
