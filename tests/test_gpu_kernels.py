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


@pytest.mark.parametrize("variant,K,temp,detach,masked", [
    ("mort", 2, 1.0, False, True), ("pheno", 25, 1.0, False, True), ("pheno", 25, 2.0, True, True),
    ("mort", 2, 1.3, False, False), ("pheno", 3, 1.7, False, True), ("pheno", 32, 1.0, False, True)])
def test_routing_fwd_bwd_vs_oracle(variant, K, temp, detach, masked):
    """Routing kernels alone (route embeddings given) against the oracle, fp32, incl. all gradients."""
    from multimodalrouting_b200 import ops  # noqa: F401
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    B = 37
    _, sdp, sdh = synth.make_state(K=K, seed=5 + K, sharp=4.0)
    g = torch.Generator().manual_seed(11)
    embs = {r: torch.randn(B, 256, generator=g) for r in synth.ROUTES}
    rm = None
    if masked:
        rm = (torch.rand(B, 10, generator=g) < 0.75).float()
        rm[0] = 0.0          # a patient with every route masked
        rm[1] = 1.0
    gl = torch.randn(B, K, generator=g)
    gR = torch.randn(B, 10, K, generator=g)
    # oracle (CPU fp32)
    po = {k: v.clone().requires_grad_(True) for k, v in sdp.items()}
    ho = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
    eo = {r: v.clone().requires_grad_(True) for r, v in embs.items()}
    lo, ao, Ro = orc.routing_forward(po, ho, eo, variant=variant, route_mask=rm, act_temperature=temp,
                                     detach_priors=detach)
    ((lo * gl).sum() + (Ro * gR).sum()).backward()
    # device path
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
    proj.load_state_dict(sdp)
    head.load_state_dict(sdh)
    proj, head = proj.cuda(), head.cuda()
    ed = {r: v.clone().cuda().requires_grad_(True) for r, v in embs.items()}
    l, a, _, R = rh.forward_capsule_from_route_dict(ed, proj, head, route_mask=None if rm is None else rm.cuda(),
                                                    act_temperature=temp, detach_priors=detach)
    assert max_rel(l, lo) < 1e-4 and max_rel(a, ao) < 1e-4 and max_rel(R, Ro) < 1e-4
    ((l * gl.cuda()).sum() + (R * gR.cuda()).sum()).backward()
    for r in synth.ROUTES:
        assert max_rel(ed[r].grad, eo[r].grad) < 5e-4, f"d emb {r}"
        assert max_rel(proj.proj[r].weight.grad, po[f"proj.{r}.weight"].grad) < 5e-4, f"d proj_w {r}"
        assert max_rel(proj.proj[r].bias.grad, po[f"proj.{r}.bias"].grad) < 5e-4, f"d proj_b {r}"
    assert max_rel(head.capsule.w.grad, ho["capsule.w"].grad) < 5e-4
    assert max_rel(head.pose_to_mc.weight.grad, ho["pose_to_mc.weight"].grad) < 5e-4
    assert max_rel(head.embedding.grad, ho["embedding"].grad) < 5e-4
    assert max_rel(head.bias.grad, ho["bias"].grad) < 5e-4
    assert head.capsule.beta_u.grad is None and head.capsule.beta_a.grad is None


def test_head_forward_from_poses():
    """CapsuleMortalityHead.forward(prim_pose, prim_act, route_mask) standalone (both variants)."""
    for variant in ("mort", "pheno"):
        if variant == "mort":
            from multimodalrouting_b200.MortModel import routing_and_heads as rh
        else:
            from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
        K, B = 25, 19
        _, _, sdh = synth.make_state(K=K, seed=9, sharp=3.0)
        g = torch.Generator().manual_seed(3)
        pose = torch.randn(B, 10, 32, generator=g)
        act = torch.rand(B, 10, generator=g)
        rm = (torch.rand(B, 10, generator=g) < 0.8).float()
        ho = {k: v.clone().requires_grad_(True) for k, v in sdh.items()}
        p0, a0 = pose.clone().requires_grad_(True), act.clone().requires_grad_(True)
        lo, alo, Ro = orc.capsule_head_forward(ho, p0, a0, rm, variant=variant)
        gl = torch.randn(B, K, generator=g)
        (lo * gl).sum().backward()
        head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
        head.load_state_dict(sdh)
        head = head.cuda()
        p1, a1 = pose.clone().cuda().requires_grad_(True), act.clone().cuda().requires_grad_(True)
        l, al, R = head(p1, a1, route_mask=rm.cuda())
        assert max_rel(l, lo) < 1e-4 and max_rel(al, alo) < 1e-4 and max_rel(R, Ro) < 1e-4
        (l * gl.cuda()).sum().backward()
        assert max_rel(p1.grad, p0.grad) < 5e-4
        if variant == "pheno":
            assert max_rel(a1.grad, a0.grad) < 5e-4
        assert max_rel(head.capsule.w.grad, ho["capsule.w"].grad) < 5e-4
