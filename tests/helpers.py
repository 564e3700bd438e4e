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


def load_golden(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def rebuild_case(c):
    """Same construction as oracle/gen_golden.py:build_case_inputs (no reference needed)."""
    sdm, sdp, sdh = synth.make_state(K=c["K"], orig_d_n=c["orig_d_n"], seed=c["seed"], sharp=c["sharp"])
    inp = synth.make_inputs(B=c["B"], d_n=c["orig_d_n"], K=c["K"], seed=c["seed"] + 1, missing=c["missing"],
                            TL=c.get("TL", 48), TN=c.get("TN", 16), TI=c.get("TI", 49))
    if c["mask_mode"] == "none":
        inp["mL"] = inp["mN"] = inp["mI"] = None
        inp["route_mask"] = None
    elif c["mask_mode"] == "rm1d":
        rm = torch.ones(10)
        rm[[1, 4, 8]] = 0.0
        inp["route_mask"] = rm
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
