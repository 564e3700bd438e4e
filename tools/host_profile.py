"""cProfile of one eager training step's HOST side (B=64 so the GPU is never the bottleneck)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from multimodalrouting_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
rh, mult, proj, head, _ = bench.build_models(dev)
modules = (mult, proj, head)
inp = synth.make_inputs(B=64, K=bench.K_LABELS, seed=1)
d = {k: v.to(dev) for k, v in inp.items()}
adapter = rh.RouteDimAdapter(256, 256, 256, 256)
lossf = torch.nn.BCEWithLogitsLoss()


def step():
    for m in modules:
        m.zero_grad(set_to_none=True)
    xs = [d[k].detach().requires_grad_(True) for k in ("x_l", "x_n", "x_i")]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, alpha, routes, R = rh.forward_capsule_from_multmodel(
            mult, xs[0], xs[1], xs[2], proj, head, mL=d["mL"], mN=d["mN"], mI=d["mI"], route_adapter=adapter,
            route_mask=d["route_mask"])
    loss = lossf(logits.float(), d["y"])
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
