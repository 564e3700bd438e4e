"""-m gpu: several consecutive TRAINING steps of the drop-in path -- producer epilogue -> route fusion -> capsule
routing -> loss -> backward -> clip + AdamW + EMA -- against the same sequence run by the CPU oracle (oracle forward /
backward + torch semantics of the tail).  fp32 kernels; the loss trajectory must match to 1e-4 relative (the fp32 bar of
BASELINE.json), parameters after the last step are compared on the scale of what the optimizer moved."""
import pytest
import torch

from gpu_common import build_modules, to_dev
from oracle import route_fusion_oracle as orc
from oracle import synth
from oracle import tail_oracle as to

pytestmark = pytest.mark.gpu

STEPS, LR, WD, CLIP, DECAY = 4, 2e-4, 1e-4, 0.3, 0.99


@pytest.mark.parametrize("variant,K,d_n", [("mort", 2, 768), ("pheno", 25, 256)])
def test_training_steps_match_oracle(variant, K, d_n):
    from multimodalrouting_b200 import optim, producers
    c = dict(variant=variant, K=K, orig_d_n=d_n, temp=1.2, detach=False)
    sdm, sdp, sdh = synth.make_state(K=K, orig_d_n=d_n, seed=51, sharp=2.0)
    batches = [synth.make_inputs(B=6, K=K, d_n=d_n, seed=60 + s, missing=True) for s in range(STEPS)]
    for b in batches:                       # un-clamped encoder outputs: the producer epilogue has work to do
        for k in ("x_l", "x_n", "x_i"):
            b[k] = b[k] * 1.7

    # ---- drop-in path on the GPU
    rh, mult, proj, head = build_modules(c, sdm, sdp, sdh)
    modules = (mult, proj, head)
    params = [p for m in modules for p in m.parameters()]
    names = [f"{i}.{n}" for i, m in enumerate(modules) for n, _ in m.named_parameters()]
    opt = optim.FusedAdamW(params, lr=LR, weight_decay=WD)
    ema = optim.EMA(modules, decay=DECAY)
    adapter = rh.RouteDimAdapter(256, 256, 256, 256)
    losses, norms = [], []
    for b in batches:
        d = to_dev(b)
        z = {m: producers._sanitize_encoder_out({"seq": d[k], "mask": d[mk]}, m)
             for m, k, mk in (("L", "x_l", "mL"), ("N", "x_n", "mN"), ("I", "x_i", "mI"))}
        logits, _, _, _ = rh.forward_capsule_from_multmodel(
            mult, z["L"]["seq"], z["N"]["seq"], z["I"]["seq"], proj, head, mL=z["L"]["mask"], mN=z["N"]["mask"],
            mI=z["I"]["mask"], route_adapter=adapter, route_mask=d["route_mask"], act_temperature=1.2)
        loss = synth.loss_fn(logits, d["y"], variant)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step(max_norm=CLIP, ema=ema)
        losses.append(float(loss.detach()))
        norms.append(float(opt.total_norm))
    assert int(opt.step_count) == STEPS

    # ---- the same sequence on the CPU oracle
    sds = [{k: v.clone().requires_grad_(True) for k, v in sd.items()} for sd in (sdm, sdp, sdh)]
    flat = [(f"{i}.{k}", v) for i, sd in enumerate(sds) for k, v in sd.items()]
    assert [n for n, _ in flat] == names
    m_ = {n: torch.zeros_like(v) for n, v in flat}
    v_ = {n: torch.zeros_like(v) for n, v in flat}
    e_ = {n: v.detach().clone() for n, v in flat}
    ref_losses, ref_norms, step = [], [], 0
    for b in batches:
        for _, v in flat:
            v.grad = None
        xs = [to.sanitize_mort(b[k]) for k in ("x_l", "x_n", "x_i")]
        logits, _, _, _ = orc.full_forward(sds[0], sds[1], sds[2], xs[0], xs[1], xs[2], b["mL"], b["mN"], b["mI"],
                                           variant=variant, route_mask=b["route_mask"], act_temperature=1.2)
        loss = synth.loss_fn(logits, b["y"], variant)
        loss.backward()
        used = [(n, v) for n, v in flat if v.grad is not None]
        grads = [v.grad.clone() for _, v in used]
        ref_norms.append(to.clip_grad_norm(grads, CLIP))
        assert to.grads_are_finite(grads)
        step += 1
        with torch.no_grad():
            ps = [v for _, v in used]
            to.adamw_step(ps, grads, [m_[n] for n, _ in used], [v_[n] for n, _ in used], step, LR, 0.9, 0.999, 1e-8, WD)
            to.ema_update([e_[n] for n, _ in used], ps, DECAY)
        ref_losses.append(float(loss.detach()))

    for a, r in zip(losses, ref_losses):
        assert abs(a - r) <= 1e-4 * abs(r), (losses, ref_losses)
    for a, r in zip(norms, ref_norms):
        assert abs(a - r) <= 2e-3 * r, (norms, ref_norms)
    # parameters: Adam divides by sqrt(v), so an entry whose gradient is at the fp32 noise floor can move by a full
    # +-lr per step in either implementation; judge the update on the scale of what the optimizer moved
    moved_budget = STEPS * LR
    sh = ema.shadow_by_storage()
    worst_frac = 0.0
    for (n, ref), p in zip(flat, params):
        diff = (p.detach().cpu() - ref.detach()).abs()
        assert float(diff.max()) <= 2.5 * moved_budget, n
        worst_frac = max(worst_frac, float((diff > 0.02 * moved_budget).float().mean()))
        ediff = (sh[p.data_ptr()].cpu() - e_[n]).abs()
        assert float(ediff.max()) <= 2.5 * moved_budget * (1 - DECAY) * STEPS, n
    assert worst_frac < 0.02, worst_frac      # >98 % of every tensor's entries agree to 2 % of the applied update
