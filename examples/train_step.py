"""A complete PhenoModel training step on the B200 path, as the reference's train loop would run it after switching
imports -- needs a B200; shown here as the usage example of the pieces either side of the
route-fusion hot path:

    encoder outputs --_sanitize_encoder_out--> MULTModel + capsule routing --pheno_train_loss--> backward
                    --FusedAdamW.step(max_norm, ema)--> next step

Reference lines: PhenoModel/Paired_Cross_Attention/main.py:2700-2860 (train loop), :1452-1460 (sanitize),
:1412-1443 (capsule_forward_from_encoded), :2764-2812 (loss), :2814-2860 (clip / finite guard / step / EMA).
Every stage is device-side and never synchronises the host, so forward + loss + backward is captured into one CUDA
graph (the optimizer tail is graph-capturable too; it is issued eagerly here so that LR schedulers keep working on the
Python side).  Synthetic tensors stand in for the BEHRT / BioClinicalBERT / CNN encoder outputs.

    python examples/train_step.py [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodalrouting_b200 import MULTModel, losses, optim, producers, synth  # noqa: E402
from multimodalrouting_b200.graphs import GraphedStep  # noqa: E402
from multimodalrouting_b200.PhenoModel import routing_and_heads as rh  # noqa: E402


def main(steps: int = 10, B: int = 512, K: int = 25):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    mult = MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False).to(dev)
    projector = rh.RoutePrimaryProjector(256, 32).to(dev)
    cap_head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K).to(dev)
    route_adapter = rh.RouteDimAdapter(256, 256, 256, 256)
    modules = (mult, projector, cap_head)
    params = [p for m in modules for p in m.parameters()]
    optimizer = optim.FusedAdamW(params, lr=2e-4, weight_decay=1e-4)
    ema = optim.EMA(modules, decay=0.999)
    pos_weight = torch.full((K,), 2.0, device=dev)

    # static input tensors of the captured step; a data loader copies each batch into them
    batch = {k: v.to(dev) for k, v in synth.make_inputs(B=B, K=K, seed=1, missing=True).items()}
    loss_state = losses.LossState(dev)

    def fwd_bwd():
        for m in modules:
            m.zero_grad(set_to_none=True)
        z = {m: producers._sanitize_encoder_out({"seq": batch[k], "mask": batch[mk]}, m, variant="pheno")
             for m, k, mk in (("L", "x_l", "mL"), ("N", "x_n", "mN"), ("I", "x_i", "mI"))}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, prim_acts, _, rc_raw = rh.forward_capsule_from_multmodel(
                mult, z["L"]["seq"], z["N"]["seq"], z["I"]["seq"], projector, cap_head,
                mL=z["L"]["mask"], mN=z["N"]["mask"], mI=z["I"]["mask"], route_adapter=route_adapter,
                route_mask=batch["route_mask"], act_temperature=1.0, detach_priors=False)
        loss = losses.pheno_train_loss(logits, batch["y"], rc_raw, prim_acts, batch["route_mask"], pos_weight=pos_weight,
                                       route_entropy_lambda=0.01, route_uniform_lambda=0.1, cur_epoch=3.0,
                                       state=loss_state).loss
        loss.backward()
        return loss

    step = GraphedStep(fwd_bwd, warmup=2)
    for s in range(steps):
        fresh = synth.make_inputs(B=B, K=K, seed=2 + s, missing=True)
        for k, v in fresh.items():
            batch[k].copy_(v, non_blocking=True)
        loss = step()                                       # one graph launch: sanitize + fusion + routing + loss + backward
        optimizer.step(max_norm=0.3, ema=ema)               # clip + finite guard + AdamW + EMA, no host sync
        if s % 5 == 0:                                      # the only host reads, for logging
            print(f"step {s}: loss {float(loss):.4f}  base {float(loss_state.base):.4f}  grad norm {float(optimizer.total_norm):.3f}  "
                  f"skipped {int(optimizer.skipped)}  rc info: {loss_state.check()}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
