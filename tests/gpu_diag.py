"""GPU diagnostics (not a test): prints per-stage error magnitudes, each section in its own process
so that a CUDA fault in one section cannot poison the others.  Usage: python tests/gpu_diag.py [section]"""
import os
import subprocess
import sys
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

SECTIONS = ["simt_gemm", "routing", "fusion_fp32", "fusion_bf16_simt", "tc_gemm", "tc_wgrad", "fusion_bf16_tc"]


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def sec_simt_gemm():
    import torch
    from multimodalrouting_b200 import ops
    g = torch.Generator().manual_seed(1)
    for dt, name in ((ops.DTYPE_F32, "f32"), (ops.DTYPE_BF16, "bf16")):
        for (M, N, K) in [(128, 256, 256), (300, 256, 1024), (77, 64, 48)]:
            A = torch.randn(M, K, generator=g).cuda(); B = torch.randn(N, K, generator=g).cuda()
            if dt == ops.DTYPE_BF16:
                A, B = A.bfloat16(), B.bfloat16()
            bias = torch.randn(N, generator=g).cuda()
            C = ops.debug_gemm(ops.GEMM_SIMT, dt, False, A, B, bias)
            print(f"simt {name} TN {M}x{N}x{K}: {rel(C, A.double() @ B.double().t() + bias.double()):.2e}")
            Y = torch.randn(M, 96, generator=g).cuda(); X = torch.randn(M, 72, generator=g).cuda()
            if dt == ops.DTYPE_BF16:
                Y, X = Y.bfloat16(), X.bfloat16()
            W = ops.debug_gemm(ops.GEMM_SIMT, dt, True, Y, X, None)
            print(f"simt {name} wgrad rows={M}: {rel(W, Y.double().t() @ X.double()):.2e}")


def sec_tc_gemm():
    import torch
    from multimodalrouting_b200 import ops
    g = torch.Generator().manual_seed(2)
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (300, 512, 1024), (4096, 256, 2048)]:
        A = torch.randn(M, K, generator=g).cuda().bfloat16(); B = torch.randn(N, K, generator=g).cuda().bfloat16()
        bias = torch.randn(N, generator=g).cuda()
        C = ops.debug_gemm(ops.GEMM_TC, ops.DTYPE_BF16, False, A, B, bias)
        torch.cuda.synchronize()
        print(f"tc TN {M}x{N}x{K}: {rel(C, A.double() @ B.double().t() + bias.double()):.2e}", flush=True)


def sec_tc_wgrad():
    import torch
    from multimodalrouting_b200 import ops
    g = torch.Generator().manual_seed(3)
    for (Kr, M, N) in [(64, 128, 256), (128, 128, 256), (1000, 256, 1024), (777, 2048, 256)]:
        Y = torch.randn(Kr, M, generator=g).cuda().bfloat16(); X = torch.randn(Kr, N, generator=g).cuda().bfloat16()
        W = ops.debug_gemm(ops.GEMM_TC, ops.DTYPE_BF16, True, Y, X, None)
        torch.cuda.synchronize()
        print(f"tc wgrad rows={Kr} {M}x{N}: {rel(W, Y.double().t() @ X.double()):.2e}", flush=True)


def sec_routing():
    import torch
    from oracle import route_fusion_oracle as orc, synth
    for variant, K, temp, masked in (("mort", 2, 1.0, True), ("pheno", 25, 2.0, True), ("pheno", 25, 1.0, False)):
        if variant == "mort":
            from multimodalrouting_b200.MortModel import routing_and_heads as rh
        else:
            from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
        B = 37
        _, sdp, sdh = synth.make_state(K=K, seed=5 + K, sharp=4.0)
        g = torch.Generator().manual_seed(11)
        embs = {r: torch.randn(B, 256, generator=g) for r in synth.ROUTES}
        rm = (torch.rand(B, 10, generator=g) < 0.75).float() if masked else None
        gl = torch.randn(B, K, generator=g); gR = torch.randn(B, 10, K, generator=g)
        po = {k: v.clone().requires_grad_(True) for k, v in sdp.items()}
        ho = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
        eo = {r: v.clone().requires_grad_(True) for r, v in embs.items()}
        lo, ao, Ro = orc.routing_forward(po, ho, eo, variant=variant, route_mask=rm, act_temperature=temp)
        ((lo * gl).sum() + (Ro * gR).sum()).backward()
        proj = rh.RoutePrimaryProjector(256, 32); head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
        proj.load_state_dict(sdp); head.load_state_dict(sdh); proj, head = proj.cuda(), head.cuda()
        ed = {r: v.clone().cuda().requires_grad_(True) for r, v in embs.items()}
        l, a, _, R = rh.forward_capsule_from_route_dict(ed, proj, head, route_mask=None if rm is None else rm.cuda(),
                                                        act_temperature=temp)
        print(f"routing {variant} K={K} T={temp} fwd: logits {rel(l, lo):.2e} alpha {rel(a, ao):.2e} R {rel(R, Ro):.2e}")
        ((l * gl.cuda()).sum() + (R * gR.cuda()).sum()).backward()
        torch.cuda.synchronize()
        e_emb = max(rel(ed[r].grad, eo[r].grad) for r in synth.ROUTES)
        e_pw = max(rel(proj.proj[r].weight.grad, po[f"proj.{r}.weight"].grad) for r in synth.ROUTES)
        e_pb = max(rel(proj.proj[r].bias.grad, po[f"proj.{r}.bias"].grad) for r in synth.ROUTES)
        print(f"   bwd: d_emb {e_emb:.2e} d_projw {e_pw:.2e} d_projb {e_pb:.2e} d_w {rel(head.capsule.w.grad, ho['capsule.w'].grad):.2e} "
              f"d_mc {rel(head.pose_to_mc.weight.grad, ho['pose_to_mc.weight'].grad):.2e} d_emb {rel(head.embedding.grad, ho['embedding'].grad):.2e} "
              f"d_bias {rel(head.bias.grad, ho['bias'].grad):.2e}", flush=True)


def _fusion(autocast, engine, names):
    import torch
    from gpu_common import run_case
    from helpers import load_golden, r_grad_probe, rebuild_case, checksum
    if engine:
        os.environ["MMR_B200_GEMM"] = engine
    for name in names:
        gold = load_golden(name); c = gold["case"]
        sdm, sdp, sdh, inp = rebuild_case(c)
        t0 = time.time()
        out = run_case(c, sdm, sdp, sdh, inp, autocast=autocast, r_probe=r_grad_probe(c, gold["R"].shape))
        torch.cuda.synchronize()
        per_route = [rel(out["routes"][:, i], gold["routes"][:, i]) for i in range(10)]
        print(f"{name}: routes " + " ".join(f"{e:.1e}" for e in per_route))
        print(f"   logits {rel(out['logits'], gold['logits']):.2e} alpha {rel(out['alpha'], gold['alpha']):.2e} "
              f"R {rel(out['R'], gold['R']):.2e} loss {out['loss']:.6f} vs {gold['loss']:.6f}  ({time.time() - t0:.1f}s)")
        worst = []
        for n, (gn, gp, gs) in gold["grad_checksum"].items():
            g = out["grads"].get(n)
            if g is None:
                worst.append((float("inf"), n, "MISSING")); continue
            cn, cp, _ = checksum(n, g)
            worst.append((max(abs(cn - gn), abs(cp - gp) / 6) / max(gn, 1e-12), n, f"norm {cn:.3e}/{gn:.3e}"))
        worst.sort(reverse=True)
        print("   worst grad checksums: " + "; ".join(f"{n} {e:.1e} ({s})" for e, n, s in worst[:6]))
        fulls = sorted(((rel(out["grads"][k], g), k) for k, g in gold["grad_full"].items() if out["grads"].get(k) is not None), reverse=True)
        print("   worst full grads: " + "; ".join(f"{k} {e:.1e}" for e, k in fulls[:6]), flush=True)


def sec_fusion_fp32():
    _fusion(False, None, ["mort_cfg1", "pheno_sharp4", "mort_missing", "pheno_odd", "mort_nomask"])


def sec_fusion_bf16_simt():
    _fusion(True, "simt", ["mort_cfg1", "pheno_sharp4", "pheno_odd"])


def sec_fusion_bf16_tc():
    _fusion(True, "tc", ["mort_cfg1", "pheno_sharp4", "mort_missing", "pheno_odd"])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        try:
            globals()["sec_" + sys.argv[1]]()
        except Exception:
            traceback.print_exc()
            sys.exit(1)
        sys.exit(0)
    for s in SECTIONS:
        print(f"===== {s} =====", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=420, capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"[{s}] FAILED rc={r.returncode}\n" + r.stderr[-3000:])
        except subprocess.TimeoutExpired:
            print(f"[{s}] TIMEOUT")
