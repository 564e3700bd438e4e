"""Micro-benchmark of the producer epilogue and the training tail (csrc/tail.cuh) against the HBM roofline.

    python tools/bench_tail.py            # prints one JSON object

Sizes: the route inputs of BASELINE configs[1] (B=512: L 48 / N 16 / I 49 tokens x 256, fp32) and the 20.4 M
parameters of the PhenoModel hot path.  Timing: CUDA events over ITERS back-to-back launches after warm-up; the
working sets (145-820 MB) exceed the 126 MB L2.  torch's own foreach clip + AdamW + EMA sequence (what the reference
runs, main.py:3143-3165) is timed next to it as the library reference point."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def timeit(fn, iters=20, warm=3):
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


def main():
    import bench
    from multimodalrouting_b200 import optim, producers
    from oracle import tail_oracle as to
    peak = bench.peaks()[1]
    dev = torch.device("cuda", 0)
    out = {"hbm_peak_gbs": peak}
    # ---- producer epilogue
    B = 512
    xs = [torch.randn(B, T, 256, device=dev) * 2.0 for T in (48, 16, 49)]
    nbytes = sum(x.numel() for x in xs) * 8          # read fp32 + write fp32
    ms = timeit(lambda: [producers.sanitize_rows(x, 0, 20.0) for x in xs])
    ms_ref = timeit(lambda: [to.sanitize_mort(x) for x in xs])     # the reference's eager sequence on the same GPU
    out["sanitize_fwd"] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                           "algorithmic_bytes": nbytes, "reference_eager_torch_ms": ms_ref}
    big = torch.randn(4096 * 48, 256, device=dev)
    nb = big.numel() * 8
    ms = timeit(lambda: producers.sanitize_rows(big, 0, 20.0))
    out["sanitize_fwd_201MB"] = {"ms": ms, "GBps": nb / ms / 1e6, "frac_of_hbm_peak": nb / ms / 1e6 / peak}
    # ---- training tail on the real parameter set
    rh, mult, proj, head, _ = bench.build_models(dev)
    params = [p for m in (mult, proj, head) for p in m.parameters()]
    n = sum(p.numel() for p in params)
    flat_g = torch.randn(n, device=dev) * 1e-3
    o = 0
    for p in params:
        p.grad = flat_g[o:o + p.numel()].view(p.shape)
        o += p.numel()
    opt = optim.FusedAdamW(params, lr=2e-4, weight_decay=1e-4)
    ema = optim.EMA((mult, proj, head), decay=0.999)
    ms = timeit(lambda: opt.step(max_norm=0.3, ema=ema))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step(max_norm=0.3, ema=ema)
    ms_g = timeit(graph.replay)
    algo = n * 4 * (1 + 5 + 4)        # norm pass reads g; update reads p,g,m,v,ema and writes p,m,v,ema
    out["clip_adamw_ema"] = {"params": n, "tensors": len(params), "ms_eager_issue": ms, "ms_cuda_graph": ms_g,
                             "GBps": algo / ms_g / 1e6, "frac_of_hbm_peak": algo / ms_g / 1e6 / peak,
                             "algorithmic_bytes": algo}
    # torch foreach sequence the reference runs (library reference point)
    ref_opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=1e-4)
    shadow = [p.detach().clone() for p in params]

    def ref_step():
        torch.nn.utils.clip_grad_norm_(params, max_norm=0.3)
        ref_opt.step()
        torch._foreach_mul_(shadow, 0.999)
        torch._foreach_add_(shadow, [p.detach() for p in params], alpha=0.001)
    out["clip_adamw_ema"]["torch_foreach_ms"] = timeit(ref_step)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
