"""Shared helpers for the parity tests (golden loading, checksum comparison)."""
import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import synth  # noqa: E402

GOLDEN_CASES = ["mort_cfg1", "pheno_sharp4", "pheno_warm", "mort_missing", "pheno_missing",
                "mort_nomask", "pheno_rm1d", "pheno_odd"]
# long sequences (PhenoModel's structured_seq_len=256; INSPECT token counts of BASELINE configs[4]): pin the ORACLE to the
# reference at these token counts; the GPU tests reach them through the oracle (test_bf16_mma_attention_..., tools/stress_shapes.py)
GOLDEN_LONG = ["pheno_tl256", "pheno_inspect", "pheno_override", "mort_iter2", "pheno_layers2", "mort_proj_all",   # + acts_override, num_routing=2, layers=2, Conv1d on L/N/I
               "pheno_missing1", "mort_missing1", "mort_override", "pheno_override_grad"]   # sharp=1 missing-modality pair; gradient through acts_override


def load_golden(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def rebuild_case(c):
    """Same construction as oracle/gen_golden.py:build_case_inputs (no reference needed)."""
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
        if c.get("override_grad"):
            inp["acts_override"].requires_grad_(True)
    return sdm, sdp, sdh, inp


def r_grad_probe(c, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(c["seed"] + 7))


def proj_vec(name, n):
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randn(n, generator=g)


def checksum(name, t):
    f = t.detach().double().flatten().cpu()
    return [float(f.norm()), float(f @ proj_vec(name, f.numel()).double()), float(f.sum())]


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b, floor=1e-6):
    """max |a-b| / (max|b| + floor): elementwise error relative to the tensor's scale."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + floor))


def check_grad_checksums(got: dict, gold: dict, tol: float, what=""):
    """got: name -> tensor; gold: name -> [norm, proj, sum].  Error model: an error vector e with
    |e| <= tol*|g| moves the random projection by ~tol*|g|."""
    bad = []
    for n, (gn, gp, _gs) in gold.items():
        if n not in got or got[n] is None:
            bad.append((n, "missing"))
            continue
        cn, cp, _ = checksum(n, got[n])
        scale = max(gn, 1e-12)
        if abs(cn - gn) > tol * scale + 1e-9 or abs(cp - gp) > 6 * tol * scale + 1e-9:
            bad.append((n, f"norm {cn:.6e} vs {gn:.6e}; proj {cp:.6e} vs {gp:.6e}"))
    assert not bad, f"{what}: {len(bad)} gradient checksum mismatches, first: {bad[:5]}"


def fp64_truth(c, sdm, sdp, sdh, inp):
    """The oracle evaluated in float64: the exact-arithmetic answer.  Routing with sharpened votes is
    ill-conditioned for some patients (a 1e-6 input perturbation can move logits by 3e-4), so the
    reference's own fp32 output carries that much rounding noise; tests therefore accept an
    implementation that is as close to the fp64 truth as the reference's fp32 result is."""
    from oracle import route_fusion_oracle as orc
    d = lambda sd: {k: v.double() for k, v in sd.items()}
    f = lambda t: None if t is None else t.detach().double()
    logits, alpha, routes, R = orc.full_forward(
        d(sdm), d(sdp), d(sdh), inp["x_l"].double(), inp["x_n"].double(), inp["x_i"].double(),
        f(inp["mL"]), f(inp["mN"]), f(inp["mI"]), variant=c["variant"], route_mask=f(inp["route_mask"]),
        act_temperature=c["temp"], detach_priors=c["detach"], acts_override=f(inp.get("acts_override")),
        num_routing=c.get("iters", 3), layers=c.get("layers", 4))
    return {"logits": logits, "alpha": alpha, "R": R,
            "routes": torch.stack([routes[r] for r in synth.ROUTES], dim=1)}


def routing_amplification(c, sdp, sdh, inp, routes_bt, eps=1e-4):
    """Per-patient condition estimate of the routing stage: relative change of (logits, R) caused by a
    relative perturbation eps of the route embeddings, divided by eps (fp64)."""
    from oracle import route_fusion_oracle as orc
    d = lambda sd: {k: v.double() for k, v in sd.items()}
    rm = None if inp["route_mask"] is None else inp["route_mask"].double()
    base = {r: routes_bt[:, i].double() for i, r in enumerate(synth.ROUTES)}
    g = torch.Generator().manual_seed(1234)
    pert = {r: v * (1 + eps * torch.randn(v.shape, generator=g, dtype=torch.float64)) for r, v in base.items()}
    ao = None if inp.get("acts_override") is None else inp["acts_override"].detach().double()
    l0, _, R0 = orc.routing_forward(d(sdp), d(sdh), base, variant=c["variant"], route_mask=rm, act_temperature=c["temp"],
                                    acts_override=ao, num_routing=c.get("iters", 3))
    l1, _, R1 = orc.routing_forward(d(sdp), d(sdh), pert, variant=c["variant"], route_mask=rm, act_temperature=c["temp"],
                                    acts_override=ao, num_routing=c.get("iters", 3))
    dl = (l1 - l0).abs().amax(dim=1) / l0.abs().max().clamp_min(1e-12)
    dR = (R1 - R0).abs().amax(dim=(1, 2)) / R0.abs().max().clamp_min(1e-12)
    return torch.maximum(dl, dR) / eps


def oracle_run(c, sdm, sdp, sdh, inp, r_probe, dtype, device="cpu", autocast=False):
    """Forward + loss (+ R probe) + backward of the oracle in `dtype` on `device`.  float64 on the CPU gives the
    exact-arithmetic answer used by the conditioning-aware criteria; ``device="cuda", autocast=True`` is the
    reference's own mixed-precision path (same modules under torch.autocast("cuda", bfloat16), TF32 as the drivers set it,
    M/main.py:2613-2621,2641) -- the yardstick for the bf16 kernels.  Returns (outputs, gradients)."""
    from oracle import route_fusion_oracle as orc
    cv = lambda sd: {k: v.detach().to(device=device, dtype=dtype).clone().requires_grad_(True) for k, v in sd.items()}
    f = lambda t: None if t is None else t.detach().to(device=device, dtype=dtype)
    a, b, h = cv(sdm), cv(sdp), cv(sdh)
    xs = {k: inp[k].detach().to(device=device, dtype=dtype).clone().requires_grad_(True) for k in ("x_l", "x_n", "x_i")}
    ao = f(inp.get("acts_override"))
    if ao is not None and c.get("override_grad"):
        ao.requires_grad_(True)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
    with ctx:
        logits, alpha, routes, R = orc.full_forward(a, b, h, xs["x_l"], xs["x_n"], xs["x_i"], f(inp["mL"]), f(inp["mN"]),
                                                    f(inp["mI"]), variant=c["variant"], route_mask=f(inp["route_mask"]),
                                                    act_temperature=c["temp"], detach_priors=c["detach"],
                                                    acts_override=ao, num_routing=c.get("iters", 3), layers=c.get("layers", 4))
    total = synth.loss_fn(logits.float() if autocast else logits, inp["y"].to(device=device, dtype=logits.dtype if not autocast else torch.float32), c["variant"])
    if r_probe is not None:
        total = total + 0.05 * (R.to(total.dtype) * r_probe.to(device=device, dtype=total.dtype)).sum()
    total.backward()
    g = {}
    for sd in (a, b, h):
        for k, v in sd.items():
            g[k] = v.grad
    for k, v in xs.items():
        g[k] = v.grad
    if ao is not None and ao.requires_grad:
        g["acts_override"] = ao.grad
    out = {"logits": logits.detach(), "alpha": alpha.detach(), "R": R.detach(),
           "routes": torch.stack([routes[r].detach() for r in synth.ROUTES], dim=1)}
    return out, g


def oracle_grads(c, sdm, sdp, sdh, inp, r_probe, dtype):
    """All parameter / input gradients of the oracle (same loss + R probe as gpu_common.run_case) in
    `dtype`; float64 gives the exact-arithmetic gradients used by the conditioning-aware criterion."""
    return oracle_run(c, sdm, sdp, sdh, inp, r_probe, dtype)[1]


def per_patient_err(a, b, floor=1e-6):
    """[B] max |a-b| over everything but the batch dimension, relative to the whole tensor's scale."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = (a - b).abs().reshape(a.shape[0], -1).amax(dim=1)
    return d / (b.abs().max() + floor)
