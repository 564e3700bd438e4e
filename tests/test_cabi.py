"""CPU checks of the C-ABI boundary: the library builds/loads, exports every symbol declared in
include/mmr_b200.h, and the host-only planning entry points behave (no kernels are launched)."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT


def _lib():
    from multimodalrouting_b200 import _lib, build
    if not os.path.exists(build.LIB):
        build.build()
    return _lib.load(), _lib


def test_exports_match_header():
    lib, _ = _lib()
    hdr = open(os.path.join(ROOT, "include", "mmr_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mmr_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mmr_b200.h but not exported"


def test_plan_sizes_and_param_count():
    lib, L = _lib()
    assert lib.mmr_version() >= 100
    d = L.FusionDims(16, 48, 16, 49, 256, 768, 256, 4, 0, 0)
    assert lib.mmr_fusion_num_params(C.byref(d)) == 317
    s = [C.c_size_t() for _ in range(4)]
    assert lib.mmr_fusion_sizes(C.byref(d), *[C.byref(v) for v in s]) == 0
    packed, saved, sf, sb = [v.value for v in s]
    assert packed > 2 * 19_000_000 * 4 * 0.9      # forward + transposed fp32 copies of the GEMM weights
    assert saved > 0 and sf > 0 and sb > 0
    d16 = L.FusionDims(16, 48, 16, 49, 256, 768, 256, 4, 1, 0)
    s2 = [C.c_size_t() for _ in range(4)]
    assert lib.mmr_fusion_sizes(C.byref(d16), *[C.byref(v) for v in s2]) == 0
    assert s2[0].value < packed and s2[1].value < saved       # bf16 buffers are smaller


@pytest.mark.parametrize("bad", [dict(B=0), dict(layers=0), dict(layers=9), dict(dN=100), dict(dtype=7)])
def test_invalid_dims_rejected(bad):
    lib, L = _lib()
    kw = dict(B=4, TL=48, TN=16, TI=49, dL=256, dN=256, dI=256, layers=4, dtype=0, gemm_engine=0)
    kw.update(bad)
    d = L.FusionDims(*[kw[k] for k in ("B", "TL", "TN", "TI", "dL", "dN", "dI", "layers", "dtype", "gemm_engine")])
    s = [C.c_size_t() for _ in range(4)]
    assert lib.mmr_fusion_sizes(C.byref(d), *[C.byref(v) for v in s]) != 0
    assert len(lib.mmr_last_error_string()) > 0


def test_routing_dims_rejected_without_launch():
    lib, L = _lib()
    d = L.RoutingDims(4, 40, 0, 3, 0, 0, 1.0, 0.02, 0.98, 1024, 256)   # K too large
    p = L.RoutingParams()
    assert lib.mmr_capsule_routing_fwd(C.byref(d), C.byref(p), *([None] * 11)) != 0


def test_cpu_tensors_fail_loudly():
    import torch
    from oracle import synth
    from multimodalrouting_b200 import MULTModel
    m = MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    inp = synth.make_inputs(B=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(inp["x_l"], inp["x_n"], inp["x_i"], inp["mL"], inp["mN"], inp["mI"])


def test_state_dict_keys_match_reference_layout():
    from oracle import synth
    from multimodalrouting_b200 import MULTModel
    from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    m = MULTModel(256, 768, 256, 256, 256, 256, True, True, True, 8, 4, 0, 0., 0., 0., 0., 0., 0., 0., False)
    spec = synth.mult_param_spec(256, 768, 256)
    assert [n for n, _ in m.named_parameters()] == [n for n, _, _ in spec]
    assert [tuple(p.shape) for p in m.parameters()] == [s for _, s, _ in spec]
    assert [id(p) for p in m._param_list()] == [id(p) for p in m.parameters()]
    assert sum(p.numel() for p in m.parameters()) == 19_877_376
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=25)
    assert set(head.state_dict()) == {"embedding", "bias", "capsule.w", "capsule.beta_u", "capsule.beta_a",
                                      "pose_to_mc.weight"}
    proj = rh.RoutePrimaryProjector(256, 32)
    assert set(proj.state_dict()) == {f"proj.{r}.{k}" for r in synth.ROUTES for k in ("weight", "bias")}


def test_ctypes_mirrors_match_the_compiled_structs():
    """mmr_abi_struct_sizes: the binding refuses to load a library whose public structs differ from its mirrors."""
    lib, L = _lib()
    sizes = (C.c_size_t * 7)()
    assert lib.mmr_abi_struct_sizes(sizes, 7) == 7
    mirrors = (L.FusionDims, L.RoutingDims, L.RoutingParams, L.RoutingGrads, L.OptTensor, L.OptHyper)
    assert [C.sizeof(m) for m in mirrors] == list(sizes)[:6]
    assert sizes[6] == L.OPT_STATE_BYTES
    assert lib.mmr_abi_struct_sizes(sizes, 2) == 2 and lib.mmr_abi_struct_sizes(None, 7) == 0


def test_routing_pack_rejects_bad_arguments_without_launch():
    lib, L = _lib()
    p = L.RoutingParams()
    assert lib.mmr_routing_pack_weights(C.byref(p), 25, 16, 32, None, None) != 0      # caps_w missing
    p.caps_w = 64
    assert lib.mmr_routing_pack_weights(C.byref(p), 0, 16, 32, None, None) != 0       # K out of range
    assert lib.mmr_routing_pack_weights(C.byref(p), 25, None, 32, None, None) != 0    # output missing
    from multimodalrouting_b200 import ops
    assert ops.routing_pack_bytes(25) == 2 * 10 * 25 * 64 * 32 * 2 + 10 * 40 * 256 * 2


def test_header_is_plain_c99_and_a_c_host_links():
    """include/mmr_b200.h compiles as C99 (-pedantic -Werror) and examples/c_host.c links against the library and runs
    its host-only entry points."""
    import shutil
    import subprocess
    import tempfile
    _lib()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    from multimodalrouting_b200 import build
    csrc = os.path.dirname(build.LIB)
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "c_host")
        cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
               os.path.join(ROOT, "examples", "c_host.c"), "-L", csrc, "-lmmr_b200", f"-Wl,-rpath,{csrc}", "-o", exe]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert run.returncode == 0, run.stdout + run.stderr
        assert "params=317" in run.stdout and "packed=" in run.stdout
