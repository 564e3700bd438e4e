"""Builds the sm_100a CUDA library in-tree (csrc/libmmr_b200.so) with nvcc.

The .so is a plain C-ABI shared object (include/mmr_b200.h); it has no torch or Python
dependency, so it is compiled directly with nvcc rather than through torch.utils.cpp_extension.
"""
from __future__ import annotations

import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libmmr_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["mmr_common.cuh", "epilogue.cuh", "gemm_simt.cuh", "gemm_tc.cuh", "plan.cuh", "rows.cuh",
           "attention.cuh", "attention_mma.cuh", "attention_tc.cuh", "routing.cuh", "routing_split.cuh", "projector.cuh", "producer.cuh", "tail.cuh", "loss.cuh", os.path.join("..", "..", "include", "mmr_b200.h")]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the sm_100a kernels")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
