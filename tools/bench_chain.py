"""Micro-benchmark + check of the chained FFN kernel (fc1 + ReLU + fc2 in one launch) at B=512 shapes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodalrouting_b200 import _lib  # noqa: E402

lib = _lib.load()
M = int(os.environ.get("M", 115712))
ITERS = int(os.environ.get("ITERS", 20))
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
B1 = (torch.randn(1024, 256, device="cuda", generator=g) / 16).bfloat16()
B2 = (torch.randn(256, 1024, device="cuda", generator=g) / 32).bfloat16()
b1 = torch.randn(1024, device="cuda", generator=g) * 0.1
b2 = torch.randn(256, device="cuda", generator=g) * 0.1
mid = torch.empty(M, 1024, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
bits = torch.zeros(M, 32, device="cuda", dtype=torch.int32)
ms = C.c_float()
rc = lib.mmr_bench_chain(1, M, A.data_ptr(), B1.data_ptr(), B2.data_ptr(), b1.data_ptr(), b2.data_ptr(), None,
                         bits.data_ptr(), mid.data_ptr(), out.data_ptr(), ITERS, C.byref(ms),
                         torch.cuda.current_stream().cuda_stream)
_lib.check(rc, "chain")
fl = 2 * 2.0 * M * 1024 * 256
print(f"chain fwd: {ms.value*1e3:.1f} us  {fl/ms.value/1e9:.1f} TFLOP/s")
ref_mid = (A[:512].float() @ B1.float().t() + b1).relu()
ref_out = ref_mid.bfloat16().float() @ B2.float().t() + b2
print("mid err", float((mid[:512].float() - ref_mid).abs().max() / ref_mid.abs().max()),
      "out err", float((out[:512].float() - ref_out).abs().max() / ref_out.abs().max()))
