"""Micro-benchmark of the loss tail (csrc/loss.cuh) on one GPU.

    python tools/bench_loss.py            # prints one JSON object

Workloads: BASELINE configs[1] (Pheno, B=512, K=25, bf16 R as the bf16 path returns it, pos_weight, both regularisers on)
and the per-GPU share of configs[2] (Mort, B=4096).  Timed: forward + backward of the loss (a) issued eagerly through
losses.py, (b) replayed from a CUDA graph (what a graphed training step pays), next to (c) the reference's eager
sequence -- coerce_rc_to_report with its `.item()` syncs, BCEWithLogits, regularisers, autograd backward -- restated in
oracle/tail_oracle.py and run on the same GPU.  The data is a few hundred KB: every number here is launch latency."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def graphed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def main():
    from multimodalrouting_b200 import _lib, losses
    from oracle import tail_oracle as to
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(5)
    lib = _lib.load()
    out = {}
    # ---- Pheno, BASELINE configs[1]
    B, K = 512, 25
    mask = (torch.rand(B, 10, generator=gen) < 0.8).float(); mask[:, 0] = 1.0
    q = (torch.rand(B, 10, K, generator=gen) + 1e-3) * mask.unsqueeze(-1)
    rc = (q / q.sum(1, keepdim=True)).to(torch.bfloat16).to(dev)
    logits = (torch.randn(B, K, generator=gen) * 3).to(dev)
    y = (torch.rand(B, K, generator=gen) < 0.2).float().to(dev)
    pw = (torch.rand(K, generator=gen) * 4 + 0.5).to(dev)
    pa = torch.rand(B, 10, generator=gen).to(dev)
    mask = mask.to(dev)
    st = losses.LossState(dev)

    def ours():
        lg = logits.detach().requires_grad_(True)
        losses.pheno_train_loss(lg, y, rc, pa, mask, pos_weight=pw, route_entropy_lambda=0.01, route_uniform_lambda=0.1,
                                state=st).loss.backward()

    def ref():
        lg = logits.detach().requires_grad_(True)
        to.pheno_train_loss(lg, y, rc, pa, mask, pw, 0.01, 0, 0.1, 0, 1.0)["loss"].backward()

    n0 = lib.mmr_launch_count()
    ours()
    launches = lib.mmr_launch_count() - n0
    g = graphed(ours)
    nbytes = B * K * 4 * 3 + B * 10 * K * (2 + 4) + B * 10 * 8       # logits, y, dlogits; R in (bf16) + rc_report out; alpha, mask
    out["pheno_B512_K25"] = {"our_kernels": int(launches), "eager_ms": timeit(ours), "cuda_graph_ms": timeit(g.replay),
                             "reference_eager_torch_ms": timeit(ref, iters=20), "algorithmic_bytes": nbytes}
    # ---- Mort, configs[2] per-GPU share at 2 GPUs
    B = 4096
    logits2 = (torch.randn(B, 2, generator=gen) * 2).to(dev)
    y2 = (torch.rand(B, generator=gen) < 0.15).long().to(dev)
    pa2 = torch.rand(B, 10, generator=gen).to(dev)
    st2 = losses.LossState(dev)

    def ours2():
        lg = logits2.detach().requires_grad_(True)
        losses.mort_train_loss(lg, y2, pa2, label_smoothing=0.02, route_entropy_lambda=0.01, route_uniform_lambda=0.1,
                               state=st2).loss.backward()

    def ref2():
        lg = logits2.detach().requires_grad_(True)
        to.mort_train_loss(lg, y2, pa2, 0.02, 0.01, 0, 0.1, 0, 1)["loss"].backward()

    g2 = graphed(ours2)
    out["mort_B4096"] = {"eager_ms": timeit(ours2), "cuda_graph_ms": timeit(g2.replay),
                         "reference_eager_torch_ms": timeit(ref2, iters=20)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
