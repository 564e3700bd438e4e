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
# data-gradient pair: dF = (G W2) .* relu' (bits), dH = (dF W1) .* rowmask  -- B1 = fc2.weight^T [1024,256], B2 = fc1.weight^T [256,1024]
G = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
W2T = B2.t().contiguous()     # [1024, 256]
W1T = B1.t().contiguous()     # [256, 1024]
rowmask = (torch.rand(M, device="cuda", generator=g) < 0.9).float()
dF = torch.empty(M, 1024, device="cuda", dtype=torch.bfloat16)
dH = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
rc = lib.mmr_bench_chain(0, M, G.data_ptr(), W2T.data_ptr(), W1T.data_ptr(), None, None, rowmask.data_ptr(),
                         bits.data_ptr(), dF.data_ptr(), dH.data_ptr(), ITERS, C.byref(ms),
                         torch.cuda.current_stream().cuda_stream)
_lib.check(rc, "chain bwd")
print(f"chain bwd: {ms.value*1e3:.1f} us  {fl/ms.value/1e9:.1f} TFLOP/s")
keep = (mid[:512] > 0).float()
ref_dF = (G[:512].float() @ W2T.float().t()) * keep
ref_dH = (ref_dF.bfloat16().float() @ W1T.float().t()) * rowmask[:512, None]
print("dF err", float((dF[:512].float() - ref_dF).abs().max() / ref_dF.abs().max()),
      "dH err", float((dH[:512].float() - ref_dH).abs().max() / ref_dH.abs().max()))
