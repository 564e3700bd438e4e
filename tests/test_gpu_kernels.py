"""-m gpu: unit parity of the GEMM engines and the routing kernels through the C ABI."""
import pytest
import torch

from helpers import max_rel
from oracle import route_fusion_oracle as orc
from oracle import synth

pytestmark = pytest.mark.gpu


def _ops():
    from multimodalrouting_b200 import ops
    return ops


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (300, 256, 1024), (77, 64, 48), (1024, 1024, 256)])
def test_simt_gemm_fp32(M, N, K):
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    C = ops.debug_gemm(ops.GEMM_SIMT, ops.DTYPE_F32, False, A, B, bias)
    ref = A.double() @ B.double().t() + bias.double()
    assert max_rel(C, ref) < 1e-5
    Y = torch.randn(M, 96, generator=g).cuda()
    X = torch.randn(M, 72, generator=g).cuda()
    W = ops.debug_gemm(ops.GEMM_SIMT, ops.DTYPE_F32, True, Y, X, None)
    assert max_rel(W, Y.double().t() @ X.double()) < 1e-5


def _accurate(mine, ref32, truth64, tol, what):
    """Within `tol` of the fp32 oracle, or -- where the computation is ill-conditioned -- at least as
    close to the exact (fp64) answer as 3x the fp32 oracle's own rounding error."""
    e = max_rel(mine, ref32)
    if e < tol:
        return
    e_ref, e_mine = max_rel(ref32, truth64), max_rel(mine, truth64)
    assert e_mine <= max(tol, 3.0 * e_ref), f"{what}: vs fp32 oracle {e:.2e}; vs fp64 mine {e_mine:.2e} oracle {e_ref:.2e}"


def _oracle_routing(dt, sdp, sdh, embs, rm, gl, gR, variant, temp, detach):
    po = {k: v.clone().to(dt).requires_grad_(True) for k, v in sdp.items()}
    ho = {k: v.clone().to(dt).requires_grad_(True) for k, v in sdh.items()}
    eo = {r: v.clone().to(dt).requires_grad_(True) for r, v in embs.items()}
    lo, ao, Ro = orc.routing_forward(po, ho, eo, variant=variant, route_mask=None if rm is None else rm.to(dt),
                                     act_temperature=temp, detach_priors=detach)
    ((lo * gl.to(dt)).sum() + (Ro * gR.to(dt)).sum()).backward()
    return lo, ao, Ro, po, ho, eo


@pytest.mark.parametrize("variant,K,temp,detach,masked", [
    ("mort", 2, 1.0, False, True), ("pheno", 25, 1.0, False, True), ("pheno", 25, 2.0, True, True),
    ("mort", 2, 1.3, False, False), ("pheno", 3, 1.7, False, True), ("pheno", 32, 1.0, False, True)])
def test_routing_fwd_bwd_vs_oracle(variant, K, temp, detach, masked):
    """Routing kernels alone (route embeddings given) against the oracle, fp32, incl. all gradients."""
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    B = 37
    _, sdp, sdh = synth.make_state(K=K, seed=5 + K, sharp=2.0)
    g = torch.Generator().manual_seed(11)
    embs = {r: 0.5 * torch.randn(B, 256, generator=g) for r in synth.ROUTES}
    rm = None
    if masked:
        rm = (torch.rand(B, 10, generator=g) < 0.75).float()
        rm[0] = 0.0          # a patient with every route masked
        rm[1] = 1.0
    gl = torch.randn(B, K, generator=g)
    gR = torch.randn(B, 10, K, generator=g)
    lo, ao, Ro, po, ho, eo = _oracle_routing(torch.float32, sdp, sdh, embs, rm, gl, gR, variant, temp, detach)
    lt, at, Rt, pt, ht, et = _oracle_routing(torch.float64, sdp, sdh, embs, rm, gl, gR, variant, temp, detach)
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
    proj.load_state_dict(sdp)
    head.load_state_dict(sdh)
    proj, head = proj.cuda(), head.cuda()
    ed = {r: v.clone().cuda().requires_grad_(True) for r, v in embs.items()}
    l, a, _, R = rh.forward_capsule_from_route_dict(ed, proj, head, route_mask=None if rm is None else rm.cuda(),
                                                    act_temperature=temp, detach_priors=detach)
    _accurate(l, lo, lt, 1e-4, "logits"); _accurate(a, ao, at, 1e-4, "alpha"); _accurate(R, Ro, Rt, 1e-4, "R")
    ((l * gl.cuda()).sum() + (R * gR.cuda()).sum()).backward()
    for r in synth.ROUTES:
        _accurate(ed[r].grad, eo[r].grad, et[r].grad, 2e-4, f"d emb {r}")
        _accurate(proj.proj[r].weight.grad, po[f"proj.{r}.weight"].grad, pt[f"proj.{r}.weight"].grad, 2e-4, f"d proj_w {r}")
        _accurate(proj.proj[r].bias.grad, po[f"proj.{r}.bias"].grad, pt[f"proj.{r}.bias"].grad, 2e-4, f"d proj_b {r}")
    for k, p in (("capsule.w", head.capsule.w), ("pose_to_mc.weight", head.pose_to_mc.weight),
                 ("embedding", head.embedding), ("bias", head.bias)):
        _accurate(p.grad, ho[k].grad, ht[k].grad, 2e-4, f"d {k}")
    assert head.capsule.beta_u.grad is None and head.capsule.beta_a.grad is None


def test_head_forward_from_poses():
    """CapsuleMortalityHead.forward(prim_pose, prim_act, route_mask) standalone (both variants)."""
    for variant in ("mort", "pheno"):
        if variant == "mort":
            from multimodalrouting_b200.MortModel import routing_and_heads as rh
        else:
            from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
        K, B = 25, 19
        _, _, sdh = synth.make_state(K=K, seed=9, sharp=2.0)
        g = torch.Generator().manual_seed(3)
        pose = 0.5 * torch.randn(B, 10, 32, generator=g)
        act = torch.rand(B, 10, generator=g)
        rm = (torch.rand(B, 10, generator=g) < 0.8).float()
        gl = torch.randn(B, K, generator=g)
        res = {}
        for dt in (torch.float32, torch.float64):
            ho = {k: v.clone().to(dt).requires_grad_(True) for k, v in sdh.items()}
            p0, a0 = pose.clone().to(dt).requires_grad_(True), act.clone().to(dt).requires_grad_(True)
            lo, alo, Ro = orc.capsule_head_forward(ho, p0, a0, rm.to(dt), variant=variant)
            (lo * gl.to(dt)).sum().backward()
            res[dt] = (lo, alo, Ro, p0, a0, ho)
        o32, o64 = res[torch.float32], res[torch.float64]
        head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
        head.load_state_dict(sdh)
        head = head.cuda()
        p1, a1 = pose.clone().cuda().requires_grad_(True), act.clone().cuda().requires_grad_(True)
        l, al, R = head(p1, a1, route_mask=rm.cuda())
        _accurate(l, o32[0], o64[0], 1e-4, "logits"); _accurate(al, o32[1], o64[1], 1e-4, "alpha")
        _accurate(R, o32[2], o64[2], 1e-4, "R")
        (l * gl.cuda()).sum().backward()
        _accurate(p1.grad, o32[3].grad, o64[3].grad, 2e-4, "d pose")
        if variant == "pheno":
            _accurate(a1.grad, o32[4].grad, o64[4].grad, 2e-4, "d act")
        _accurate(head.capsule.w.grad, o32[5]["capsule.w"].grad, o64[5]["capsule.w"].grad, 2e-4, "d w")


@pytest.mark.parametrize("variant,K,masked", [("pheno", 25, True), ("mort", 2, True), ("pheno", 32, False), ("pheno", 3, True)])
def test_routing_reduced_precision_tensor_core_vs_cuda_core(variant, K, masked, monkeypatch):
    """Reduced-precision routing (what runs under autocast): the tensor-core projector / vote contraction
    (mma.sync, fp16 weight copies from mmr_routing_pack_weights) against the CUDA-core path of the same mode and the fp32
    oracle.  Bars: logits / alpha / R within 2e-2 of the oracle (BASELINE.json bf16 tolerance); gradients at least
    as close to the oracle as the CUDA-core reduced-precision path (x2) or 5e-2."""
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    B = 45          # 11 full tiles of 4 patients + a ragged one
    _, sdp, sdh = synth.make_state(K=K, seed=15 + K, sharp=2.0)
    g = torch.Generator().manual_seed(21)
    embs = {r: 0.5 * torch.randn(B, 256, generator=g) for r in synth.ROUTES}
    rm = None
    if masked:
        rm = (torch.rand(B, 10, generator=g) < 0.75).float()
        rm[0] = 0.0
        rm[1] = 1.0
    gl = torch.randn(B, K, generator=g)
    gR = torch.randn(B, 10, K, generator=g)
    lo, ao, Ro, po, ho, eo = _oracle_routing(torch.float32, sdp, sdh, embs, rm, gl, gR, variant, 1.2, False)
    res = {}
    for eng in ("0", "1"):
        monkeypatch.setenv("MMR_RT_TC", eng)
        proj = rh.RoutePrimaryProjector(256, 32)
        head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
        proj.load_state_dict(sdp); head.load_state_dict(sdh)
        proj, head = proj.cuda(), head.cuda()
        ed = {r: v.clone().cuda().requires_grad_(True) for r, v in embs.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            l, a, _, R = rh.forward_capsule_from_route_dict(ed, proj, head, route_mask=None if rm is None else rm.cuda(),
                                                            act_temperature=1.2)
        ((l.float() * gl.cuda()).sum() + (R.float() * gR.cuda()).sum()).backward()
        grads = {f"emb {r}": ed[r].grad for r in synth.ROUTES}
        grads.update({f"proj_w {r}": proj.proj[r].weight.grad for r in synth.ROUTES})
        grads.update({"capsule.w": head.capsule.w.grad, "pose_to_mc": head.pose_to_mc.weight.grad,
                      "embedding": head.embedding.grad, "bias": head.bias.grad})
        res[eng] = (l.float(), a.float(), R.float(), grads)
    ref_g = {f"emb {r}": eo[r].grad for r in synth.ROUTES}
    ref_g.update({f"proj_w {r}": po[f"proj.{r}.weight"].grad for r in synth.ROUTES})
    ref_g.update({"capsule.w": ho["capsule.w"].grad, "pose_to_mc": ho["pose_to_mc.weight"].grad,
                  "embedding": ho["embedding"].grad, "bias": ho["bias"].grad})
    for eng in ("0", "1"):
        l, a, R, _ = res[eng]
        assert max_rel(l, lo) < 2e-2 and max_rel(a, ao) < 2e-2 and max_rel(R, Ro) < 2e-2, eng
        if rm is not None:          # masked routes: exactly zero alpha and R, as in the fp32 path
            assert float((a.cpu() * (1 - rm)).abs().max()) == 0.0
            assert float((R.cpu() * (1 - rm).unsqueeze(-1)).abs().max()) == 0.0
    for k, ref in ref_g.items():
        e_tc, e_cc = max_rel(res["1"][3][k], ref), max_rel(res["0"][3][k], ref)
        assert bool(torch.isfinite(res["1"][3][k]).all()), k
        assert e_tc <= max(5e-2, 2.0 * e_cc), f"grad {k}: tensor-core {e_tc:.2e} cuda-core {e_cc:.2e}"


@pytest.mark.parametrize("variant,K,nit,masked,B", [
    ("pheno", 25, 3, True, 45), ("pheno", 25, 1, True, 17), ("pheno", 16, 3, True, 33), ("pheno", 12, 2, False, 40),
    ("pheno", 8, 3, True, 21), ("mort", 5, 2, True, 50), ("pheno", 3, 3, True, 70), ("mort", 2, 3, True, 130), ("mort", 1, 3, False, 9)])
def test_routing_split_path_vs_tile_path(variant, K, nit, masked, B, monkeypatch):
    """Reduced-precision routing, round-2 split path (csrc/routing_split.cuh: projector / vote GEMMs on 16-patient tiles, one
    patient per 1-4 warps for the agreement iterations; every lane-layout instantiation KP = 2, 4, 8, 16, 32 and num_routing
    1..3) against the tile-of-4 kernels of the same mode (MMR_RT_SPLIT=0) and the fp32 oracle, ragged last tiles included."""
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    _, sdp, sdh = synth.make_state(K=K, seed=31 + K, sharp=2.0)
    g = torch.Generator().manual_seed(100 + K + nit)
    embs = {r: 0.5 * torch.randn(B, 256, generator=g) for r in synth.ROUTES}
    rm = None
    if masked:
        rm = (torch.rand(B, 10, generator=g) < 0.75).float()
        rm[0] = 0.0
        rm[1] = 1.0
    gl = torch.randn(B, K, generator=g)
    gR = torch.randn(B, 10, K, generator=g)
    po = {k: v.clone().requires_grad_(True) for k, v in sdp.items()}
    ho = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
    eo = {r: v.clone().requires_grad_(True) for r, v in embs.items()}
    lo, ao, Ro = orc.routing_forward(po, ho, eo, variant=variant, route_mask=rm, act_temperature=1.3, num_routing=nit)
    ((lo * gl).sum() + (Ro * gR).sum()).backward()
    res = {}
    for split in ("0", "1"):
        monkeypatch.setenv("MMR_RT_SPLIT", split)
        proj = rh.RoutePrimaryProjector(256, 32)
        head = rh.CapsuleMortalityHead(32, 64, nit, 0.0, "EM", num_classes=K)
        proj.load_state_dict(sdp); head.load_state_dict(sdh)
        proj, head = proj.cuda(), head.cuda()
        ed = {r: v.clone().cuda().requires_grad_(True) for r, v in embs.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            l, a, _, R = rh.forward_capsule_from_route_dict(ed, proj, head, route_mask=None if rm is None else rm.cuda(),
                                                            act_temperature=1.3)
        ((l.float() * gl.cuda()).sum() + (R.float() * gR.cuda()).sum()).backward()
        grads = {f"emb {r}": ed[r].grad for r in synth.ROUTES}
        grads.update({f"proj_w {r}": proj.proj[r].weight.grad for r in synth.ROUTES})
        grads.update({f"proj_b {r}": proj.proj[r].bias.grad for r in synth.ROUTES})
        grads.update({"capsule.w": head.capsule.w.grad, "pose_to_mc": head.pose_to_mc.weight.grad,
                      "embedding": head.embedding.grad, "bias": head.bias.grad})
        res[split] = (l.float().detach(), a.float().detach(), R.float().detach(), grads)
    ref_g = {f"emb {r}": eo[r].grad for r in synth.ROUTES}
    ref_g.update({f"proj_w {r}": po[f"proj.{r}.weight"].grad for r in synth.ROUTES})
    ref_g.update({f"proj_b {r}": po[f"proj.{r}.bias"].grad for r in synth.ROUTES})
    ref_g.update({"capsule.w": ho["capsule.w"].grad, "pose_to_mc": ho["pose_to_mc.weight"].grad,
                  "embedding": ho["embedding"].grad, "bias": ho["bias"].grad})
    for split in ("0", "1"):
        l, a, R, _ = res[split]
        assert max_rel(l, lo) < 2e-2 and max_rel(a, ao) < 2e-2 and max_rel(R, Ro) < 2e-2, split
        if rm is not None:
            assert float((a.cpu() * (1 - rm)).abs().max()) == 0.0
            assert float((R.cpu() * (1 - rm).unsqueeze(-1)).abs().max()) == 0.0
    # same precision mode, same fp16 weight copies: the two paths differ only in summation order
    assert max_rel(res["1"][0], res["0"][0].cpu()) < 5e-3 and max_rel(res["1"][2], res["0"][2].cpu()) < 5e-3
    for k, ref in ref_g.items():
        if ref is None:
            continue
        e_s, e_t = max_rel(res["1"][3][k], ref), max_rel(res["0"][3][k], ref)
        assert bool(torch.isfinite(res["1"][3][k]).all()), k
        assert e_s <= max(5e-2, 2.0 * e_t), f"grad {k}: split {e_s:.2e} tile {e_t:.2e}"


@pytest.mark.parametrize("variant", ["mort", "pheno"])
def test_routing_split_path_head_from_poses(variant, monkeypatch):
    """CapsuleMortalityHead.forward(prim_pose, prim_act, route_mask) under autocast: the from_poses form of the split path
    (no projector launch; d pose / d act returned) against the tile-of-4 kernels and the fp32 oracle."""
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    K, B = 25, 37
    _, _, sdh = synth.make_state(K=K, seed=9, sharp=2.0)
    g = torch.Generator().manual_seed(3)
    pose = 0.5 * torch.randn(B, 10, 32, generator=g)
    act = torch.rand(B, 10, generator=g)
    rm = (torch.rand(B, 10, generator=g) < 0.8).float()
    gl = torch.randn(B, K, generator=g)
    ho = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
    p0, a0 = pose.clone().requires_grad_(True), act.clone().requires_grad_(True)
    lo, alo, Ro = orc.capsule_head_forward(ho, p0, a0, rm, variant=variant)
    (lo * gl).sum().backward()
    res = {}
    for split in ("0", "1"):
        monkeypatch.setenv("MMR_RT_SPLIT", split)
        head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
        head.load_state_dict(sdh)
        head = head.cuda()
        p1, a1 = pose.clone().cuda().requires_grad_(True), act.clone().cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            l, al, R = head(p1, a1, route_mask=rm.cuda())
        (l.float() * gl.cuda()).sum().backward()
        res[split] = (l.float().detach(), R.float().detach(), p1.grad, a1.grad, head.capsule.w.grad)
    for split in ("0", "1"):
        assert max_rel(res[split][0], lo) < 2e-2 and max_rel(res[split][1], Ro) < 2e-2, split
    assert max_rel(res["1"][0], res["0"][0].cpu()) < 5e-3
    for i, ref in ((2, p0.grad), (4, ho["capsule.w"].grad)) + (((3, a0.grad),) if variant == "pheno" else ()):
        e_s, e_t = max_rel(res["1"][i], ref), max_rel(res["0"][i], ref)
        assert e_s <= max(5e-2, 2.0 * e_t), f"grad {i}: split {e_s:.2e} tile {e_t:.2e}"


@pytest.mark.parametrize("B", [1, 19, 64])
def test_projector_standalone_forward_backward_vs_oracle(B):
    """RoutePrimaryProjector.forward by itself (routing_and_heads.py:111-121; csrc/projector.cuh): poses / acts and every
    gradient (route embeddings, the 10 weight / bias pairs) against the oracle's projector_forward, fp32 bar 1e-4."""
    from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    _, sdp, _ = synth.make_state(K=3, seed=77, sharp=2.0)
    g = torch.Generator().manual_seed(5)
    embs = {r: torch.randn(B, 256, generator=g) for r in synth.ROUTES}
    gp, ga = torch.randn(B, 10, 32, generator=g), torch.randn(B, 10, 1, generator=g)
    po = {k: v.clone().requires_grad_(True) for k, v in sdp.items()}
    eo = {r: v.clone().requires_grad_(True) for r, v in embs.items()}
    poses_o, acts_o = orc.projector_forward(po, eo)
    ((poses_o * gp).sum() + (acts_o * ga).sum()).backward()
    proj = rh.RoutePrimaryProjector(256, 32)
    proj.load_state_dict(sdp)
    proj = proj.cuda()
    ed = {r: v.clone().cuda().requires_grad_(True) for r, v in embs.items()}
    poses, acts = proj(ed)
    assert tuple(poses.shape) == (B, 10, 32) and tuple(acts.shape) == (B, 10, 1)
    assert max_rel(poses, poses_o) < 1e-4 and max_rel(acts, acts_o) < 1e-4
    ((poses * gp.cuda()).sum() + (acts * ga.cuda()).sum()).backward()
    for r in synth.ROUTES:
        assert max_rel(ed[r].grad, eo[r].grad) < 1e-4, f"d emb {r}"
        assert max_rel(proj.proj[r].weight.grad, po[f"proj.{r}.weight"].grad) < 1e-4, f"d W {r}"
        assert max_rel(proj.proj[r].bias.grad, po[f"proj.{r}.bias"].grad) < 1e-4, f"d b {r}"
    # only the poses are used downstream (d_acts = None inside autograd)
    for p in proj.parameters():
        p.grad = None
    e2 = {r: v.clone().cuda().requires_grad_(True) for r, v in embs.items()}
    (proj(e2)[0] * gp.cuda()).sum().backward()
    e3 = {r: v.clone().requires_grad_(True) for r, v in embs.items()}
    (orc.projector_forward(sdp, e3)[0] * gp).sum().backward()
    assert max_rel(e2["LNI"].grad, e3["LNI"].grad) < 1e-4
