"""Attention-forward engines on one GPU: mma.sync tiles (attention_mma.cuh, default) vs tcgen05 / TMEM / TMA
(attention_tc.cuh, MMR_ATTN=tc).  Times the attention-forward kernel class with CUDA events on the launching stream
(mmr_prof_*; one stream, eager launches) plus the whole fwd+bwd step, for the INSPECT token counts (BASELINE configs[4])
and the MIMIC shapes (configs[1]).  Prints one JSON line per (shape, engine).

    python tools/bench_attn_engines.py [--quick]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodalrouting_b200 import synth  # noqa: E402
from multimodalrouting_b200 import MULTModel, _lib  # noqa: E402
from multimodalrouting_b200.PhenoModel import routing_and_heads as rh  # noqa: E402

NAMES = ["gemm_tc", "wgrad_tc", "attn_fwd", "attn_bwd", "gemm_simt", "routing", "fusion_fwd_call", "fusion_bwd_call"]


def flops_qk_pv(B, TL, TN, TI, layers=4):
    """QK^T + PV FLOPs of one forward: 4 Tq Tk d per (direction, layer)  (SURVEY.md section 8d)."""
    T = {"L": TL, "N": TN, "I": TI}
    dirs = [("L", "N"), ("L", "I"), ("N", "L"), ("N", "I"), ("I", "L"), ("I", "N")]
    return B * layers * sum(4 * T[q] * T[k] * 256 for q, k in dirs)


def run(K, B, TL, TN, TI, iters):
    sdm, sdp, sdh = synth.make_state(K=K, seed=3)
    mult = MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
    mult.load_state_dict(sdm); proj.load_state_dict(sdp); head.load_state_dict(sdh)
    mult, proj, head = mult.cuda(), proj.cuda(), head.cuda()
    inp = synth.make_inputs(B=B, K=K, seed=4, TL=TL, TN=TN, TI=TI, missing=True)
    d = {k: v.cuda() for k, v in inp.items()}
    lib = _lib.load()

    def step():
        for m in (mult, proj, head):
            m.zero_grad(set_to_none=True)
        xs = [d[k].detach().requires_grad_(True) for k in ("x_l", "x_n", "x_i")]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, alpha, routes, R = rh.forward_capsule_from_multmodel(
                mult, xs[0], xs[1], xs[2], proj, head, mL=d["mL"], mN=d["mN"], mI=d["mI"],
                route_adapter=rh.RouteDimAdapter(256, 256, 256, 256), route_mask=d["route_mask"])
        synth.loss_fn(logits, d["y"], "pheno").backward()
        return logits

    os.environ["MMR_WGRAD_STREAM"] = "0"
    ref = None
    engines = ["mma", "tc2", "tc1"]        # tcN: tcgen05 engine with N heads per CTA (MMR_ATTN_TC_HEADS)
    if "--variants" in sys.argv:           # opt-in variants: p = two K / V stages (prefetch), m = per-patient 3-D tensor maps
        engines += ["tc2p", "tc1p", "tc2m"]
    for eng in engines:
        os.environ["MMR_ATTN"] = "tc" if eng.startswith("tc") else eng
        os.environ["MMR_ATTN_TC_HEADS"] = eng[2] if eng.startswith("tc") else "2"
        os.environ["MMR_ATTN_TC_PREFETCH"] = "1" if eng.startswith("tc") and "p" in eng[3:] else "0"
        os.environ["MMR_ATTN_TC_MAP3D"] = "1" if eng.startswith("tc") and "m" in eng[3:] else "0"
        for _ in range(3):
            logits = step()
        torch.cuda.synchronize()
        if ref is None:
            ref = logits.detach().float().clone()
        err = float((logits.detach().float() - ref).abs().max() / (ref.abs().max() + 1e-9))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / iters
        lib.mmr_prof_enable(1)
        for _ in range(iters):
            step()
        torch.cuda.synchronize()
        msc = (C.c_double * 8)(); nc = (C.c_longlong * 8)()
        lib.mmr_prof_collect(msc, nc)
        lib.mmr_prof_enable(0)
        attn_ms = msc[2] / iters
        fl = flops_qk_pv(B, TL, TN, TI)
        print(json.dumps({"shape": [TL, TN, TI], "B": B, "engine": eng, "attn_fwd_ms_per_step": round(attn_ms, 4),
                          "attn_fwd_launches_per_step": nc[2] / iters, "attn_fwd_tflops": round(fl / (attn_ms * 1e-3) / 1e12, 2),
                          "attn_bwd_ms_per_step": round(msc[3] / iters, 4), "step_ms_eager": round(step_ms, 3),
                          "logits_max_rel_vs_mma": err}), flush=True)
    os.environ.pop("MMR_ATTN", None)
    os.environ.pop("MMR_ATTN_TC_HEADS", None)
    os.environ.pop("MMR_ATTN_TC_PREFETCH", None)
    os.environ.pop("MMR_ATTN_TC_MAP3D", None)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    run(3, 64 if quick else 256, 512, 128, 196, 3 if quick else 5)
    run(25, 512, 48, 16, 49, 3 if quick else 5)
