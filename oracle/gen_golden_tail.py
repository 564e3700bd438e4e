"""Generate tests/golden/tail_*.pt by executing the UNMODIFIED reference definitions (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):   python oracle/gen_golden_tail.py

MortModel/PhenoModel main.py cannot be imported (module-level matplotlib / transformers downloads), so the
definitions on the two adjacent steps -- `_clamp_norm`, `_safe_tensor`, `_sanitize_encoder_out`, `grads_are_finite`,
`class EMA` -- are extracted from the source files with `ast` and executed as they are.  The optimizer is the
reference's own choice, torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ (main.py:2886-2890, 3143-3165).
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import sys
from typing import Dict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
M_MAIN = "/root/reference/MIMIC-IV/MortModel/Paired_Cross_Attention/main.py"
P_MAIN = "/root/reference/MIMIC-IV/PhenoModel/Paired_Cross_Attention/main.py"


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "Dict": Dict}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def make_rows(gen, shape, big_every=3, scale=1.0):
    x = torch.randn(*shape, generator=gen) * scale
    flat = x.view(-1, shape[-1])
    flat[::big_every] *= 6.0          # rows whose norm exceeds 20 (sqrt(256) * 6 = 96)
    flat[1::5] *= 0.05                # tiny rows
    return x


def sanitize_cases():
    m = extract(M_MAIN, ["_clamp_norm", "_safe_tensor", "_sanitize_encoder_out"])
    p = extract(P_MAIN, ["_sanitize_encoder_out"])
    gen = torch.Generator().manual_seed(9001)
    out = {}
    for name, shape, dtype in (("seq256", (3, 7, 256), torch.float32), ("seq768_bf16", (2, 5, 768), torch.bfloat16),
                               ("pool512_f16", (6, 512), torch.float16), ("seq12", (2, 3, 12), torch.float32)):
        x = make_rows(gen, shape, scale=0.2 if dtype == torch.float16 else 1.0).to(dtype)
        # finite input: forward + gradient of a random upstream gradient
        xr = x.clone().requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            y = m["_sanitize_encoder_out"]({"seq": xr, "mask": torch.ones(shape[:-1], dtype=torch.int64)}, "t")
        dy = torch.randn(*shape, generator=gen)
        (dx,) = torch.autograd.grad(y["seq"], xr, dy)
        # non-finite entries: forward only
        xb = x.clone()
        fb = xb.view(-1, shape[-1])
        fb[0, 1] = float("nan"); fb[1, 2] = float("inf"); fb[2, 3] = float("-inf")
        with contextlib.redirect_stdout(io.StringIO()):
            yb = m["_sanitize_encoder_out"]({"seq": xb}, "t")["seq"]
            yp = p["_sanitize_encoder_out"]({"seq": xb}, "t")["seq"]
        out[name] = dict(x=x, y=y["seq"].detach(), mask_dtype=str(y["mask"].dtype), dy=dy, dx=dx.detach(),
                         x_bad=xb, y_bad_mort=yb.detach(), y_bad_pheno=yp.detach())
    return out


def tail_case():
    m = extract(M_MAIN, ["grads_are_finite", "EMA"])
    gen = torch.Generator().manual_seed(9002)
    shapes = [(33, 64), (7,), (8195,), (16, 100), (1,), (256,)]
    lr, wd, max_norm, decay = 2e-4, 1e-4, 0.3, 0.999          # env_config lr; main.py:813 weight_decay; grad_clip 0.3
    model = torch.nn.Module()
    for i, sh in enumerate(shapes):
        model.register_parameter(f"p{i}", torch.nn.Parameter(torch.randn(*sh, generator=gen) * 0.5))
    params = list(model.parameters())
    init = [p.detach().clone() for p in params]
    opt = torch.optim.AdamW([{"params": params[:3], "lr": lr, "weight_decay": wd, "name": "enc"},
                             {"params": params[3:], "lr": lr, "weight_decay": wd, "name": "head"}])
    ema = m["EMA"]([model], decay=decay)
    steps, norms, skipped = [], [], []
    for s in range(5):
        scale = [1e-3, 0.5, 2e-4, 1.0, 3e-3][s]       # some steps clip (norm > 0.3), some do not
        grads = [torch.randn(*sh, generator=gen) * scale for sh in shapes]
        if s == 3:
            grads[2][17] = float("nan")                        # non-finite gradient: the whole step is skipped
        steps.append([g.clone() for g in grads])
        for p, g in zip(params, grads):
            p.grad = g
        total = torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm)
        norms.append(float(total))
        if not m["grads_are_finite"](params):
            skipped.append(True)
            opt.zero_grad(set_to_none=True)
            continue
        skipped.append(False)
        opt.step()
        ema.update()
        opt.zero_grad(set_to_none=True)
    return dict(shapes=shapes, lr=lr, weight_decay=wd, max_norm=max_norm, ema_decay=decay, betas=(0.9, 0.999),
                eps=1e-8, init=init, grads=steps, norms=norms, skipped=skipped,
                params=[p.detach().clone() for p in params],
                exp_avg=[opt.state[p]["exp_avg"].clone() for p in params],
                exp_avg_sq=[opt.state[p]["exp_avg_sq"].clone() for p in params],
                ema=[ema.shadow[0][f"p{i}"].clone() for i in range(len(shapes))],
                step=int(opt.state[params[0]]["step"]))


def route_mask_case():
    """Both of the reference's builders, executed as they are (they must agree with each other)."""
    px = "/root/reference/MIMIC-IV/PhenoModel/Partial/Cross_Attention"
    routes = ["L", "N", "I", "LN", "NL", "LI", "IL", "NI", "IN", "LNI"]      # env_config.ROUTES (M/env_config.py:53)
    src = open(f"{px}/routing_and_heads.py").read()
    tree = ast.parse(src)
    from typing import Optional
    ns = {"torch": torch, "Optional": Optional, "ROUTES": routes}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "build_route_mask_from_presence":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "rh", "exec"), ns)
    m = extract(f"{px}/main.py", [])
    ns2 = {"torch": torch, "N_ROUTES": 10}
    for node in ast.parse(open(f"{px}/main.py").read()).body:
        if isinstance(node, ast.FunctionDef) and node.name == "build_route_mask_from_modalities":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "main", "exec"), ns2)
    gen = torch.Generator().manual_seed(9003)
    has = [(torch.rand(37, generator=gen) < p).float() for p in (0.9, 0.7, 0.7)]
    a = ns["build_route_mask_from_presence"](*has)
    b = ns2["build_route_mask_from_modalities"](*has)
    assert torch.equal(a, b)
    return dict(hasL=has[0], hasN=has[1], hasI=has[2], mask=a)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.save(route_mask_case(), os.path.join(GOLD, "tail_route_mask.pt"))
    torch.save(sanitize_cases(), os.path.join(GOLD, "tail_sanitize.pt"))
    torch.save(tail_case(), os.path.join(GOLD, "tail_adamw_ema.pt"))
    for f in ("tail_sanitize.pt", "tail_adamw_ema.pt"):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    sys.exit(main())
