"""CPU: the fake (meta) implementations registered for torch.compile return the metadata of the real ops (SURVEY.md section
8b: "Must provide fake/meta kernels").  Fake CUDA tensors need no GPU; the expected byte counts come from the host-only
planner of the C ABI (mmr_fusion_sizes) and from the real ops' own allocation code."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode


def _params(mult):
    return [p.detach() for p in mult._param_list()]


@pytest.mark.parametrize("B,TL,TN,TI,d_n,dtype", [(4, 48, 16, 49, 256, 1), (3, 21, 5, 9, 768, 0)])
def test_route_fusion_fakes(B, TL, TN, TI, d_n, dtype):
    import multimodalrouting_b200 as mmr
    from multimodalrouting_b200 import ops
    mult = mmr.MULTModel(256, d_n, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    shapes = [tuple(p.shape) for p in _params(mult)]
    with FakeTensorMode():
        xs = [torch.empty(B, T, d, device="cuda") for T, d in ((TL, 256), (TN, d_n), (TI, 256))]
        ms = [torch.empty(B, T, device="cuda") for T in (TL, TN, TI)]
        pos = torch.empty(max(TL, TN, TI), 256, device="cuda")
        params = [torch.empty(s, device="cuda") for s in shapes]
        packed = ops.route_fusion_pack(params, 4, dtype, 0)
        routes, saved = ops.route_fusion_fwd(xs[0], xs[1], xs[2], ms[0], ms[1], ms[2], pos, params, packed, 4, dtype, 0)
        dims = ops._fusion_dims(xs[0], xs[1], xs[2], 4, dtype, 0)
        packed_b, saved_b, _, _ = ops.fusion_sizes(dims)
        assert tuple(routes.shape) == (10, B, 256) and routes.dtype == torch.float32 and routes.device.type == "cuda"
        assert packed.dtype == torch.uint8 and packed.numel() == packed_b > 0
        assert saved.dtype == torch.uint8 and saved.numel() == saved_b > 0
        d_routes = torch.empty(10, B, 256, device="cuda")
        dxl, dxn, dxi, flat = ops.route_fusion_bwd(xs[0], xs[1], xs[2], ms[0], ms[1], ms[2], params, packed, saved, d_routes,
                                                   [True] * len(params), 4, dtype, 0, [])
        assert dxl.shape == xs[0].shape and dxn.shape == xs[1].shape and dxi.shape == xs[2].shape
        _, total, _, _ = ops.grad_layout(shapes, 4)
        assert flat.numel() == total >= sum(torch.Size(s).numel() for s in shapes)


@pytest.mark.parametrize("K,vdt", [(25, 1), (2, 0)])
def test_capsule_routing_fakes(K, vdt, monkeypatch):
    from multimodalrouting_b200 import ops
    monkeypatch.delenv("MMR_RT_TC", raising=False)
    B = 6
    with FakeTensorMode():
        embs = torch.empty(10, B, 256, device="cuda")
        pw = [torch.empty(33, 256, device="cuda") for _ in range(10)]
        pb = [torch.empty(33, device="cuda") for _ in range(10)]
        caps_w = torch.empty(10, 32, K, 64, device="cuda")
        p2m, emb, bias = torch.empty(64, 32, device="cuda"), torch.empty(K, 64, device="cuda"), torch.empty(K, device="cuda")
        rm = torch.empty(B, 10, device="cuda")
        out = ops.capsule_routing_fwd(embs, B * 256, 256, None, None, None, rm, pw, pb, caps_w, p2m, emb, bias, B, 1, 3, False,
                                      1.0, 0.02, 0.98, vdt)
        logits, alpha, R, poses, acts, packed = out
        assert tuple(logits.shape) == (B, K) and tuple(alpha.shape) == (B, 10) and tuple(R.shape) == (B, 10, K)
        assert tuple(poses.shape) == (B, 10, 32) and tuple(acts.shape) == (B, 10)
        dims = ops._routing_dims(B, K, 1, 3, False, False, 1.0, 0.02, 0.98, B * 256, 256, vdt)
        # fp16 weight copies + the forward scratch of the split path (projector outputs, head matrix, fp16 votes)
        assert packed.dtype == torch.uint8 and packed.numel() == (ops.routing_state_bytes(dims) if vdt == ops.DTYPE_BF16 else 0)
        if vdt == ops.DTYPE_BF16:
            assert packed.numel() >= ops.routing_pack_bytes(K) + B * 10 * K * 64 * 2 + B * 330 * 4 + K * 32 * 4
        d_embs, d_poses, d_acts, flat = ops.capsule_routing_bwd(embs, B * 256, 256, None, None, None, rm, pw, pb, caps_w, p2m, emb,
                                                                bias, torch.empty(B, K, device="cuda"), None, B, 1, 3, False, 1.0,
                                                                0.02, 0.98, vdt, packed)
        assert tuple(d_embs.shape) == (10, B, 256) and d_poses.numel() == 0 and d_acts.numel() == 0
        assert flat.numel() == ops.routing_flat_layout(K)["total"]


def test_sanitize_fakes():
    from multimodalrouting_b200 import producers
    with FakeTensorMode():
        x = torch.empty(5, 7, 768, device="cuda", dtype=torch.bfloat16)
        y = producers.sanitize_rows_fwd(x, 0, 20.0)
        assert y.shape == x.shape and y.dtype == torch.float32
        dx = producers.sanitize_rows_bwd(x, torch.empty(5, 7, 768, device="cuda"), 0, 20.0)
        assert dx.shape == x.shape and dx.dtype == torch.float32


@pytest.mark.parametrize("variant,K,d_n,masks", [("pheno", 25, 256, "full"), ("mort", 2, 768, "full"), ("pheno", 3, 256, "none"),
                                                   ("mort", 2, 256, "rm1d")])
def test_full_forward_runs_on_fake_cuda_tensors(variant, K, d_n, masks):
    """The whole host side of the drop-in path -- module forward, mask plumbing, autograd.Function forward, custom ops with
    their fake kernels -- executes on fake CUDA tensors (no GPU, no arithmetic): what torch.compile / export trace, and a CPU
    check of every output's shape, dtype and device."""
    import multimodalrouting_b200 as mmr
    if variant == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    B, TL, TN, TI = 5, 48, 16, 49
    with FakeTensorMode(), torch.device("cuda"):
        mult = mmr.MULTModel(256, d_n, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
        proj, head = rh.RoutePrimaryProjector(256, 32), rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=K)
        x = [torch.randn(B, T, d, requires_grad=True) for T, d in ((TL, 256), (TN, d_n), (TI, 256))]
        m = [None, None, None] if masks == "none" else [torch.ones(B, T) for T in (TL, TN, TI)]
        rm = None if masks == "none" else (torch.ones(10) if masks == "rm1d" else torch.ones(B, 10))
        for autocast in (False, True):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                logits, alpha, routes, R = rh.forward_capsule_from_multmodel(
                    mult, x[0], x[1], x[2], proj, head, mL=m[0], mN=m[1], mI=m[2],
                    route_adapter=rh.RouteDimAdapter(256, 256, 256, 256), route_mask=rm, act_temperature=1.2)
            assert tuple(logits.shape) == (B, K) and logits.dtype == torch.float32 and logits.device.type == "cuda"
            assert tuple(alpha.shape) == (B, 10) and not alpha.requires_grad            # prim_acts come back detached
            assert tuple(R.shape) == (B, 10, K) and R.requires_grad and logits.requires_grad
            assert sorted(routes) == sorted(rh.ROUTES) and all(tuple(v.shape) == (B, 256) for v in routes.values())
            assert all(v.dtype == mult.final_lni.weight.dtype for v in routes.values())
        out = rh.forward_capsule_from_route_dict(routes, proj, head, route_mask=rm, return_routing=False)
        assert out[3] is None and tuple(out[0].shape) == (B, K)
