"""Diagnostic (GPU box): where does the bf16 gradient noise of the pheno_inspect golden come from?  Prints, per engine
combination, the worst gradient errors against the fp64 oracle next to the reference's own autocast path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from gpu_common import run_case  # noqa: E402
from helpers import load_golden, max_rel, oracle_run, r_grad_probe, rebuild_case, rel_err  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pheno_inspect"
gold = load_golden(name)
c = gold["case"]
sdm, sdp, sdh, inp = rebuild_case(c)
probe = r_grad_probe(c, gold["R"].shape)
t64, g64 = oracle_run(c, sdm, sdp, sdh, inp, probe, torch.float64)
r16, g16 = oracle_run(c, sdm, sdp, sdh, inp, probe, torch.float32, device="cuda", autocast=True)
keys = [k for k, t in g64.items() if t is not None]
eref = {k: max_rel(g16[k], g64[k]) for k in keys}
print("reference bf16: routes %.2e logits %.2e R %.2e ; worst grad %.2e" % (
    max_rel(r16["routes"], t64["routes"]), max_rel(r16["logits"], t64["logits"]), max_rel(r16["R"], t64["R"]), max(eref.values())))
l2ref = {k: rel_err(g16[k], g64[k]) for k in keys}
print("reference bf16 worst L2-relative grad error %.3e" % max(l2ref.values()))
for gemm, attn in (("tc", "mma"), ("tc", "simt"), ("simt", "mma"), ("simt", "simt")):
    os.environ["MMR_B200_GEMM"] = gemm
    os.environ["MMR_ATTN"] = attn
    out = run_case(c, sdm, sdp, sdh, inp, autocast=True, r_probe=probe)
    e = {k: max_rel(out["grads"][k], g64[k]) for k in keys}
    l2 = {k: rel_err(out["grads"][k], g64[k]) for k in keys}
    wl2 = sorted(l2.items(), key=lambda kv: -kv[1] / max(l2ref[kv[0]], 1e-9))[:3]
    print(f"   L2: worst {max(l2.values()):.3e}; n > max(8e-2, 3x ref): {sum(1 for k in keys if l2[k] > max(8e-2, 3 * l2ref[k]))}; "
          + ", ".join(f"{k.split('.')[-3:]}: {v:.3f}/{l2ref[k]:.3f}" for k, v in wl2))
    worst = sorted(e.items(), key=lambda kv: -kv[1] / max(eref[kv[0]], 1e-9))[:4]
    print(f"gemm={gemm} attn={attn}: routes {max_rel(out['routes'], t64['routes']):.2e} logits {max_rel(out['logits'], t64['logits']):.2e} "
          f"R {max_rel(out['R'], t64['R']):.2e}; worst grad {max(e.values()):.2e}; n>3x ref: "
          f"{sum(1 for k in keys if e[k] > max(8e-2, 3 * eref[k]))}; " + ", ".join(f"{k.split('.')[-3:]}: {v:.3f}/{eref[k]:.3f}" for k, v in worst))
