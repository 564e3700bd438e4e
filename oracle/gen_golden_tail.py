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


def _inline(path, first, last, ns):
    """Execute the reference's own statements `first`..`last` (1-based, inclusive; they sit inside the training loop of
    main(), so they are dedented, not imported) in namespace `ns`."""
    import textwrap
    lines = open(path).read().splitlines()[first - 1:last]
    exec(compile(textwrap.dedent("\n".join(lines)), f"{path}:{first}-{last}", "exec"), ns)
    return ns


def _anchor(path, text):
    for i, ln in enumerate(open(path).read().splitlines(), 1):
        if ln.strip().startswith(text):
            return i
    raise AssertionError(f"{text!r} not found in {path}")


def loss_cases():
    """Mort: main.py `y_f = y.float().view(-1, 1)` .. `loss = base_loss - ent_bonus + uniform_pen` (3092-3126) executed
    as they are; Pheno: coerce_rc_to_report / assert_routing_over_routes / _safe_tensor definitions executed as they
    are, then `loss = bce(logits, y.float())` .. the uniform term (2793-2812) executed as they are."""
    from typing import Optional
    gen = torch.Generator().manual_seed(9004)
    out = {"mort": [], "pheno": []}
    # ---- Mort
    m = extract(M_MAIN, ["death_logit_from_logits2", "_safe_tensor"])
    a0, a1 = _anchor(M_MAIN, "y_f = y.float().view(-1, 1)"), _anchor(M_MAIN, "loss = base_loss - ent_bonus + uniform_pen")
    assert (a0, a1) == (3092, 3126), (a0, a1)
    B = 37
    for name, ls, le, we, lu, wu, cur, bad in (("plain", 0.02, 0.0, 0, 0.0, 0, 1, False), ("regs", 0.0, 0.01, 0, 0.1, 0, 1, False),
                                               ("gated", 0.05, 0.02, 3, 0.1, 2, 2, False), ("nonfinite", 0.02, 0.01, 0, 0.1, 0, 1, True)):
        logits = (torch.randn(B, 2, generator=gen) * 2.0)
        if bad:
            logits[3, 1] = float("nan"); logits[5, 0] = float("inf"); logits[7, 1] = float("-inf")
        y = (torch.rand(B, generator=gen) < 0.15).long()
        pa = torch.sigmoid(torch.randn(B, 10, generator=gen)) * (torch.rand(B, 10, generator=gen) < 0.8).float()
        lg = logits.clone().requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            ns = dict(torch=torch, y=y, death_logit_from_logits2=m["death_logit_from_logits2"], epoch=cur - 1, start_epoch=-5,
                      step=1, label_smoothing=ls, loss_fn_train=torch.nn.BCEWithLogitsLoss(), rc_raw=None,
                      logits=m["_safe_tensor"](lg.float(), "logits(fp32)"), prim_acts=m["_safe_tensor"](pa.float(), "prim_acts(fp32)"),
                      route_entropy_lambda=le, route_entropy_warmup_epochs=we, route_uniform_lambda=lu, route_uniform_warmup_epochs=wu)
            _inline(M_MAIN, a0, a1, ns)
        (dl,) = torch.autograd.grad(ns["loss"], lg)
        out["mort"].append(dict(name=name, logits=logits, y=y, prim_acts=pa, label_smoothing=ls, lam_ent=le, warm_ent=we,
                                lam_uni=lu, warm_uni=wu, cur_epoch=cur, loss=ns["loss"].detach(), base=ns["base_loss"].detach(),
                                ent=ns["ent_bonus"].detach(), uni=ns["uniform_pen"].detach(), dlogits=dl))
    # ---- Pheno
    rh = {}
    for node in ast.parse(open("/root/reference/MIMIC-IV/PhenoModel/Paired_Cross_Attention/routing_and_heads.py").read()).body:
        if isinstance(node, ast.FunctionDef) and node.name == "route_given_pheno":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "rh", "exec"), rh)
    p = extract(P_MAIN, ["_safe_tensor", "assert_routing_over_routes"])
    pc = {"torch": torch, "Optional": Optional, "route_given_pheno": rh["route_given_pheno"]}
    for node in ast.parse(open(P_MAIN).read()).body:
        if isinstance(node, ast.FunctionDef) and node.name == "coerce_rc_to_report":
            exec(compile(ast.Module(body=[node], type_ignores=[]), P_MAIN, "exec"), pc)
    b0, b1 = _anchor(P_MAIN, "loss = bce(logits, y.float())"), _anchor(P_MAIN, "loss = loss + route_uniform_lambda * uniform_loss")
    assert (b0, b1) == (2793, 2812), (b0, b1)

    def rc_routes(B, K, mask):
        q = torch.rand(B, 10, K, generator=gen) ** 2 + 1e-3
        if mask is not None:
            q = q * mask.unsqueeze(-1)
        return q / q.sum(dim=1, keepdim=True).clamp_min(1e-10)

    for name, B, K, kind, use_mask, le, we, lu, wu, cur in (
            ("routes_mask", 37, 25, "routes", True, 0.01, 0, 0.1, 0, 1.0), ("routes_nomask", 16, 25, "routes", False, 0.01, 0, 0.1, 0, 1.0),
            ("routes_bf16", 37, 25, "routes_bf16", True, 0.01, 0, 0.1, 0, 1.0), ("forced", 21, 3, "raw", True, 0.01, 0, 0.0, 0, 1.0),
            ("forced_nomask", 9, 25, "raw", False, 0.0, 0, 0.1, 0, 1.0), ("gated", 12, 25, "routes", True, 0.01, 1, 0.1, 2, 1.0),
            ("over_labels", 8, 25, "labels", True, 0.01, 0, 0.1, 0, 1.0), ("all_masked", 8, 25, "routes_dead", True, 0.01, 0, 0.1, 0, 1.0)):
        mask = None
        if use_mask:
            mask = (torch.rand(B, 10, generator=gen) < 0.7).float()
            mask[:, 0] = 1.0
            if kind == "routes_dead":
                mask[2] = 0.0
        if kind in ("routes", "routes_dead"):
            rc = rc_routes(B, K, mask)
        elif kind == "routes_bf16":
            rc = rc_routes(B, K, mask).to(torch.bfloat16)
        elif kind == "labels":
            rc = torch.softmax(torch.randn(B, 10, K, generator=gen), dim=2)
        else:
            rc = torch.randn(B, 10, K, generator=gen) * 0.3 + 0.2
            rc[1, 2, 0] = float("nan"); rc[2, 3, 1] = float("inf")
            if mask is not None:
                rc[4] = -1.0                                   # every route clamps to 0: denominators < 1e-8 -> uniform over kept routes
        logits = torch.randn(B, K, generator=gen) * 2.0
        if name == "forced":
            logits[0, 1] = float("nan"); logits[1, 2] = float("inf")
        y = (torch.rand(B, K, generator=gen) < 0.2).float()
        pw = torch.clamp(torch.rand(K, generator=gen) * 6.0, 0.1, 5.0)
        pa = torch.sigmoid(torch.randn(B, 10, generator=gen))
        rec = dict(name=name, logits=logits, y=y, pos_weight=pw, rc_raw=rc, prim_acts=pa, route_mask=mask, lam_ent=le, warm_ent=we,
                   lam_uni=lu, warm_uni=wu, cur_epoch=cur, raises=None)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                rc_report, info = pc["coerce_rc_to_report"](rc, pa, mask, split_name="TRAIN")
        except TypeError as e:
            rec["raises"] = ("TypeError", str(e))
            out["pheno"].append(rec)
            continue
        rec.update(rc_report=rc_report, info=info)
        try:
            p["assert_routing_over_routes"](rc_report, routes_dim=1, atol=1e-3, name="TRAIN.rc_report")
        except AssertionError as e:
            rec["raises"] = ("AssertionError", str(e)[:80])
        lg = logits.clone().requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            ns = dict(torch=torch, y=y, epoch=0, bce=torch.nn.BCEWithLogitsLoss(pos_weight=pw, reduction="mean"),
                      logits=p["_safe_tensor"](lg.float(), "logits(fp32)"), routing_coef=rc_report, cur_epoch=cur,
                      route_entropy_lambda=le, route_entropy_warmup_epochs=we, route_uniform_lambda=lu, route_uniform_warmup_epochs=wu)
            _inline(P_MAIN, b0, b1, ns)
        (dl,) = torch.autograd.grad(ns["loss"], lg)
        rec.update(loss=ns["loss"].detach(), dlogits=dl)
        out["pheno"].append(rec)
    return out


def proj_cases():
    """Route-input projections: the reference's own constructor statements are executed as they are
    (encoders.py `self.proj = nn.Sequential(nn.LayerNorm(hidden), nn.Linear(hidden, d, bias=False))` and
    `self.token_proj = nn.Linear(self.token_in_dim, d, bias=False)`), then the resulting torch modules are applied as the
    encoders do (`chunk_emb = self.proj(chunk_emb)`, `I_seq = self.token_proj(tokens)`).  To keep the fixture small the two
    chunk cases share one module and the Linear weight gradient is stored as its first rows + (norm, random projection)."""
    import types
    ENC = "/root/reference/MIMIC-IV/MortModel/Paired_Cross_Attention/encoders.py"
    out = {"state": {}}
    gen = torch.Generator().manual_seed(9100)
    mods = {}
    for key, hidden in (("chunk", 768), ("token", 512)):
        self = types.SimpleNamespace(token_in_dim=hidden)
        ns = {"nn": torch.nn, "self": self, "hidden": hidden, "d": 256}
        torch.manual_seed(9100 + hidden)
        if key == "chunk":
            a = _anchor(ENC, "self.proj = nn.Sequential(")
            _inline(ENC, a, a + 3, ns)
            mods[key] = self.proj
            with torch.no_grad():        # move the LayerNorm affine away from (1, 0) so that its gradient paths are exercised
                self.proj[0].weight.add_(0.1 * torch.randn(hidden, generator=gen))
                self.proj[0].bias.add_(0.1 * torch.randn(hidden, generator=gen))
        else:
            a = _anchor(ENC, "self.token_proj = nn.Linear(")
            _inline(ENC, a, a, ns)
            mods[key] = self.token_proj
        out["state"][key] = {k: v.detach().clone() for k, v in mods[key].state_dict().items()}
    for name, key, rows_shape, hidden in (("chunk768", "chunk", (2, 16), 768), ("chunk768_ragged", "chunk", (37,), 768),
                                          ("token512", "token", (2, 49), 512)):
        mod = mods[key]
        mod.zero_grad(set_to_none=True)
        x = (torch.randn(*rows_shape, hidden, generator=gen) * 1.5 + 0.3).requires_grad_(True)
        y = mod(x)
        dy = torch.randn(y.shape, generator=gen)
        y.backward(dy)
        grads = {}
        for k, v in mod.named_parameters():
            g = v.grad.detach()
            if g.dim() == 2:
                pv = torch.randn(g.numel(), generator=torch.Generator().manual_seed(77)).double()
                grads[k] = dict(head=g[:8].clone(), norm=float(g.double().norm()), proj=float(g.double().flatten() @ pv))
            else:
                grads[k] = g.clone()
        out[name] = dict(module=key, x=x.detach(), y=y.detach(), dy=dy, dx=x.grad.detach(), grads=grads)
    return out


def route_stats_case():
    """evaluate_epoch's per-batch accumulation, main.py `rc_raw_cpu = rc_raw.detach().float().cpu()` .. `eff_sum_mat += eff_sum`
    (1918-1933), executed as it is over three batches (one with bf16 routing coefficients, as the autocast path returns them)."""
    gen = torch.Generator().manual_seed(9200)
    K = 25
    a0 = _anchor(M_MAIN, "rc_raw_cpu    = rc_raw.detach().float().cpu()")
    a1 = _anchor(M_MAIN, "eff_sum_mat    += eff_sum")
    ns = {"torch": torch, "rc_raw_sum_mat": None, "rep_sum_mat": None, "eff_sum_mat": None}
    batches, act_sum, n = [], torch.zeros(10), 0
    for i, B in enumerate((7, 64, 33)):
        raw = torch.softmax(torch.randn(B, 10, K, generator=gen), dim=1)
        if i == 1:
            raw = raw.bfloat16()
        rep = torch.softmax(torch.randn(B, 10, K, generator=gen), dim=1)
        pa = torch.rand(B, 10, generator=gen)
        ns.update(rc_raw=raw, rc_report=rep, prim_acts=pa)
        _inline(M_MAIN, a0, a1, ns)
        act_sum += pa.sum(0)             # main.py: act_sum += prim_acts.sum(dim=0) in the same loop
        n += B
        batches.append(dict(rc_raw=raw, rc_report=rep, prim_acts=pa))
    return dict(K=K, batches=batches, num_samples=n, rc_raw_sum=ns["rc_raw_sum_mat"], rc_report_sum=ns["rep_sum_mat"],
                eff_sum=ns["eff_sum_mat"], prim_act_sum=act_sum,
                avg_rc_report=ns["rep_sum_mat"] / max(1, n))    # main.py:2013-2014


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.save(route_mask_case(), os.path.join(GOLD, "tail_route_mask.pt"))
    torch.save(sanitize_cases(), os.path.join(GOLD, "tail_sanitize.pt"))
    torch.save(tail_case(), os.path.join(GOLD, "tail_adamw_ema.pt"))
    torch.save(loss_cases(), os.path.join(GOLD, "tail_loss.pt"))
    torch.save(proj_cases(), os.path.join(GOLD, "tail_proj.pt"))
    torch.save(route_stats_case(), os.path.join(GOLD, "tail_route_stats.pt"))
    for f in ("tail_sanitize.pt", "tail_adamw_ema.pt", "tail_loss.pt"):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    sys.exit(main())
