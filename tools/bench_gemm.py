"""Micro-benchmark of the persistent tcgen05 GEMM at the hot path's shapes (run on the GPU box).
Prints ms, TFLOP/s and effective HBM GB/s per shape; MMR_TC_STAGES etc. are read by the library."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodalrouting_b200 import _lib  # noqa: E402

lib = _lib.load()
M = int(os.environ.get("M", 115712))
shapes = [("q/out proj   N256 K256", 0, 256, 256), ("fc1+relu     N1024 K256", 1, 1024, 256),
          ("fc2          N256 K1024", 0, 256, 1024), ("d_fc2 bits   N1024 K256", 2, 1024, 256),
          ("kv proj      N2048 K256", 0, 2048, 256), ("f32 out      N256 K2048", 4, 256, 2048)]
ONLY = os.environ.get("ONLY")
ITERS = int(os.environ.get("ITERS", 20))
for idx, (name, op, N, K) in enumerate(shapes):
    if ONLY is not None and str(idx) not in ONLY.split(","):
        continue
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if op == 4 else torch.bfloat16)
    bits = torch.zeros(M, N // 32, device="cuda", dtype=torch.int32)
    ms = C.c_float()
    rc = lib.mmr_bench_gemm(op, M, N, K, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(), bits.data_ptr(),
                            ITERS, C.byref(ms), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "bench")
    fl = 2.0 * M * N * K
    by = M * K * 2 + M * N * (4 if op == 4 else 2) + N * K * 2
    print(f"{name}: {ms.value*1e3:8.1f} us  {fl/ms.value/1e9:8.1f} TFLOP/s  {by/ms.value/1e6:8.1f} GB/s (algorithmic)")
    if os.environ.get("CUBLAS", "1") == "1":      # library reference point for the same shape (not on the product path)
        Bt = B.t()
        for _ in range(3):
            torch.matmul(A, Bt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(ITERS):
            torch.matmul(A, Bt)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / ITERS
        print(f"    cuBLAS bf16 (no epilogue): {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TFLOP/s")
    ref = (A[:256].float() @ B.float().t())
    if op in (0, 1):
        ref = ref + bias
    if op == 1:
        ref = ref.relu()
    if op != 2:
        err = (out[:256].float() - ref).abs().max() / ref.abs().max()
        print("    rel err", float(err))
