"""Diagnostic: run-to-run differences of the gradients (eager vs eager, eager vs CUDA-graph replay).

    MMR_WGRAD_STREAM=0|1 python tools/diag_determinism.py

Prints the five largest max-relative differences per comparison.  Data-path tensors must be bit-identical between runs
(the only non-deterministic accumulations are fp32 atomics into parameter gradients, ~1e-6); a larger difference with the
side stream enabled would be an ordering bug between the two streams of the backward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    from gpu_common import build_modules, to_dev
    from multimodalrouting_b200.graphs import GraphedStep
    from multimodalrouting_b200 import synth
    B = int(os.environ.get("B", "16"))
    c = dict(variant="pheno", K=25, orig_d_n=256, temp=1.0, detach=False)
    sdm, sdp, sdh = synth.make_state(K=25, seed=31, sharp=2.0)
    rh, mult, proj, head = build_modules(c, sdm, sdp, sdh)
    modules = (mult, proj, head)
    static = to_dev(synth.make_inputs(B=B, K=25, seed=33, missing=True))
    adapter = rh.RouteDimAdapter(256, 256, 256, 256)
    keep = {}

    def fwd_bwd():
        for m in modules:
            m.zero_grad(set_to_none=True)
        xs = [static[k].detach().requires_grad_(True) for k in ("x_l", "x_n", "x_i")]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, alpha, _, R = rh.forward_capsule_from_multmodel(
                mult, xs[0], xs[1], xs[2], proj, head, mL=static["mL"], mN=static["mN"], mI=static["mI"],
                route_adapter=adapter, route_mask=static["route_mask"])
        loss = synth.loss_fn(logits.float(), static["y"], "pheno")
        loss.backward()
        keep.update(dx_l=xs[0].grad, dx_n=xs[1].grad, dx_i=xs[2].grad, logits=logits)
        return loss

    def snapshot():
        torch.cuda.synchronize()
        out = {k: v.detach().clone() for k, v in keep.items()}
        for i, m in enumerate(modules):
            for n, p in m.named_parameters():
                if p.grad is not None:
                    out[f"{i}.{n}"] = p.grad.detach().clone()
        return out

    def poison():
        """Fills ~3 GB of the caching allocator's free blocks with NaN: an op that reads scratch it never wrote
        shows up as NaN (or as a large difference) in the next step."""
        xs = [torch.full((1 << 28,), float("nan"), device="cuda") for _ in range(3)]
        torch.cuda.synchronize()
        del xs

    def report(tag, a, b):
        bad = [k for k in a if not bool(torch.isfinite(a[k].float()).all()) or not bool(torch.isfinite(b[k].float()).all())]
        if bad:
            print(f"[{tag}] NON-FINITE in {len(bad)} tensors, e.g. {bad[:4]}", flush=True)
        data = {k: float((a[k].float() - b[k].float()).abs().max()) / (float(a[k].abs().max()) + 1e-30)
                for k in ("dx_l", "dx_n", "dx_i", "logits")}
        n_diff = sum(1 for k in a if not torch.equal(a[k], b[k]))
        print(f"[{tag}] data path: " + ", ".join(f"{k} {v:.1e}" for k, v in data.items()) +
              f"; tensors that differ at all: {n_diff}/{len(a)}", flush=True)
        errs = []
        for k in a:
            scale = float(a[k].abs().max()) + 1e-30
            errs.append((float((a[k].float() - b[k].float()).abs().max()) / scale, k))
        errs.sort(reverse=True)
        print(f"[{tag}] WGRAD_STREAM={os.environ.get('MMR_WGRAD_STREAM', '1')} B={B}: " +
              "; ".join(f"{k} {e:.2e}" for e, k in errs[:5]), flush=True)

    step = GraphedStep(fwd_bwd, warmup=2)
    step(); g1 = snapshot()
    step(); g2 = snapshot()
    fwd_bwd(); e1 = snapshot()
    fwd_bwd(); e2 = snapshot()
    poison()
    fwd_bwd(); e3 = snapshot()
    report("eager vs eager after NaN-poisoning the free blocks", e1, e3)
    report("graph vs graph", g1, g2)
    report("eager vs eager", e1, e2)
    report("graph vs eager", g1, e1)


if __name__ == "__main__":
    main()
