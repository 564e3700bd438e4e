"""-m gpu: the tcgen05/TMEM/TMA GEMM engine -- unit parity against fp64 matmul and end-to-end bf16
parity of the full path running on it (the default engine for bf16)."""
import pytest
import torch

from helpers import max_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (300, 512, 1024), (1000, 1024, 256), (4096, 256, 2048)])
def test_tc_gemm_tn(M, N, K):
    from multimodalrouting_b200 import ops
    g = torch.Generator().manual_seed(2)
    A = torch.randn(M, K, generator=g).cuda().bfloat16()
    B = torch.randn(N, K, generator=g).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    C = ops.debug_gemm(ops.GEMM_TC, ops.DTYPE_BF16, False, A, B, bias)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t() + bias.double()
    assert max_rel(C, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", [(512, 256, 256), (4096, 1024, 256), (40960, 256, 1024)])
def test_tc_gemm_tn_dynamic_tile_scheduler(M, N, K, monkeypatch):
    """The same GEMM with MMR_TC_CLC=1: one cluster per work item, resident CTA pairs cancel the pending clusters of the grid
    (clusterlaunchcontrol.try_cancel) and take over their items -- 160 items for 74 resident pairs at M = 40960; M must be a multiple of 256 for the pair path."""
    from multimodalrouting_b200 import ops
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g).cuda().bfloat16()
    B = torch.randn(N, K, generator=g).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    ref = A.double() @ B.double().t() + bias.double()
    res = {}
    for clc in ("0", "1"):
        monkeypatch.setenv("MMR_TC_CLC", clc)
        res[clc] = ops.debug_gemm(ops.GEMM_TC, ops.DTYPE_BF16, False, A, B, bias)
        torch.cuda.synchronize()
        assert max_rel(res[clc], ref) < 1e-5, clc
    assert torch.equal(res["0"], res["1"])      # same tiles, same arithmetic: only the assignment of items to CTA pairs differs


@pytest.mark.parametrize("Kr,M,N", [(64, 128, 256), (128, 128, 256), (1000, 256, 1024), (5000, 1024, 256), (777, 2048, 256)])
def test_tc_gemm_wgrad(Kr, M, N):
    from multimodalrouting_b200 import ops
    g = torch.Generator().manual_seed(3)
    Y = torch.randn(Kr, M, generator=g).cuda().bfloat16()
    X = torch.randn(Kr, N, generator=g).cuda().bfloat16()
    W = ops.debug_gemm(ops.GEMM_TC, ops.DTYPE_BF16, True, Y, X, None)
    torch.cuda.synchronize()
    ref = Y.double().t() @ X.double()
    assert max_rel(W, ref) < 1e-5
