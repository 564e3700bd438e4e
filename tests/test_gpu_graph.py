"""-m gpu: the CUDA-graph step (multimodalrouting_b200.graphs.GraphedStep, the form bench.py times) reproduces the
eagerly issued step: same loss, same outputs, same gradients, also after the static inputs are overwritten."""
import pytest
import torch

from gpu_common import build_modules, to_dev
from oracle import synth

pytestmark = pytest.mark.gpu


def test_graph_replay_matches_eager_step():
    from multimodalrouting_b200.graphs import GraphedStep
    c = dict(variant="pheno", K=25, orig_d_n=256, temp=1.0, detach=False)
    sdm, sdp, sdh = synth.make_state(K=25, seed=31, sharp=2.0)
    rh, mult, proj, head = build_modules(c, sdm, sdp, sdh)
    modules = (mult, proj, head)
    batches = [to_dev(synth.make_inputs(B=16, K=25, seed=s, missing=(s == 33))) for s in (32, 33)]
    static = {k: v.clone() for k, v in batches[0].items()}
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
        keep.update(logits=logits, alpha=alpha, R=R, dx=xs[0].grad)
        return loss

    def snapshot(loss):
        torch.cuda.synchronize()
        out = {"loss": loss.detach().clone(), "logits": keep["logits"].detach().clone(),
               "R": keep["R"].detach().clone(), "dx": keep["dx"].clone()}
        for i, m in enumerate(modules):
            for n, p in m.named_parameters():
                if p.grad is not None:
                    out[f"{i}.{n}"] = p.grad.detach().clone()
        return out

    # graph first: AccumulateGrad nodes created by an eager backward on the default stream would tie the capture to
    # the legacy stream (cudaErrorStreamCaptureImplicit); bench.py captures before any eager step for the same reason
    step = GraphedStep(fwd_bwd, warmup=2)
    replayed = []
    for b in batches:
        for k in static:
            static[k].copy_(b[k])
        replayed.append(snapshot(step()))
    for b, got in zip(batches, replayed):
        for k in static:
            static[k].copy_(b[k])
        ref = snapshot(fwd_bwd())
        assert got.keys() == ref.keys()
        for k in ref:
            scale = float(ref[k].abs().max()) + 1e-20
            err = float((got[k].float() - ref[k].float()).abs().max()) / scale
            # forward results are deterministic.  Gradients are not, replayed or not: fp32 atomics (split-K weight
            # gradients, cross-direction sums) feed bf16 roundings, and tools/diag_determinism.py measures up to 7e-3
            # between two runs of the SAME path at this batch size (1e-7 with the fp32 kernels) -- so the bar here is
            # the bf16 budget of BASELINE.json, not bit equality.
            tol = 1e-5 if k in ("loss", "logits", "R") else 2e-2
            assert err <= tol, f"{k}: {err:.3e}"
