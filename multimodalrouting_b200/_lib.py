"""ctypes binding of the C ABI declared in include/mmr_b200.h (no torch types cross it)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_fp = C.c_void_p   # device pointers travel as void*


class FusionDims(C.Structure):
    _fields_ = [("B", C.c_int32), ("TL", C.c_int32), ("TN", C.c_int32), ("TI", C.c_int32),
                ("dL", C.c_int32), ("dN", C.c_int32), ("dI", C.c_int32), ("layers", C.c_int32),
                ("dtype", C.c_int32), ("gemm_engine", C.c_int32)]


class RoutingDims(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("variant", C.c_int32), ("num_routing", C.c_int32),
                ("detach_priors", C.c_int32), ("from_poses", C.c_int32), ("act_temperature", C.c_float),
                ("prior_floor", C.c_float), ("prior_ceiling", C.c_float),
                ("emb_route_stride", C.c_int64), ("emb_batch_stride", C.c_int64),
                ("vote_dtype", C.c_int32), ("reserved", C.c_int32)]


class RoutingParams(C.Structure):
    _fields_ = [("proj_w", c_fp * 10), ("proj_b", c_fp * 10), ("caps_w", c_fp), ("pose_to_mc", c_fp),
                ("embedding", c_fp), ("bias", c_fp), ("caps_wt_f16", c_fp), ("caps_w_f16", c_fp), ("proj_w_f16", c_fp)]


class RoutingGrads(C.Structure):
    _fields_ = [("proj_w", c_fp * 10), ("proj_b", c_fp * 10), ("caps_w", c_fp), ("pose_to_mc", c_fp),
                ("embedding", c_fp), ("bias", c_fp)]


class OptTensor(C.Structure):
    _fields_ = [("p", c_fp), ("g", c_fp), ("m", c_fp), ("v", c_fp), ("ema", c_fp), ("n", C.c_int64)]


class OptHyper(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("max_norm", C.c_double), ("ema_decay", C.c_double),
                ("has_ema", C.c_int32), ("reserved", C.c_int32)]


class ProjDims(C.Structure):
    _fields_ = [("rows", C.c_int64), ("d_in", C.c_int32), ("d_out", C.c_int32), ("has_ln", C.c_int32),
                ("has_bias", C.c_int32), ("x_dtype", C.c_int32), ("dtype", C.c_int32), ("gemm_engine", C.c_int32),
                ("reserved", C.c_int32)]


OPT_STATE_BYTES = 40     # sizeof(mmr_opt_state): 3 doubles, 2 floats, 2 int32
LOSS_STATE_BYTES = 48    # sizeof(mmr_loss_state): 7 floats, 2 int32, ticket, 2 reserved


class LossArgs(C.Structure):
    _fields_ = [("variant", C.c_int32), ("B", C.c_int32), ("K", C.c_int32), ("rc_dtype", C.c_int32),
                ("logits", c_fp), ("y", c_fp), ("pos_weight", c_fp), ("prim_acts", c_fp), ("rc_raw", c_fp),
                ("route_mask", c_fp), ("label_smoothing", C.c_float), ("route_entropy_lambda", C.c_float),
                ("route_uniform_lambda", C.c_float), ("atol", C.c_float), ("dlogits", c_fp), ("rc_report", c_fp),
                ("state", c_fp), ("scratch", c_fp)]


EXPORTS = ["mmr_version", "mmr_last_error_string", "mmr_fusion_num_params", "mmr_fusion_sizes",
           "mmr_route_fusion_fwd", "mmr_route_fusion_bwd", "mmr_route_fusion_bwd_events", "mmr_route_fusion_bwd_ex", "mmr_routing_scratch_bytes",
           "mmr_capsule_routing_fwd", "mmr_capsule_routing_fwd_ex", "mmr_routing_fwd_scratch_bytes", "mmr_capsule_routing_bwd", "mmr_capsule_routing_bwd_ex", "mmr_debug_gemm", "mmr_bench_gemm", "mmr_launch_count",
           "mmr_prof_enable", "mmr_prof_collect", "mmr_sanitize_rows_fwd", "mmr_sanitize_rows_bwd",
           "mmr_grad_sqnorm", "mmr_opt_prepare", "mmr_opt_apply", "mmr_ema_update", "mmr_route_mask_from_presence", "mmr_routing_pack_weights", "mmr_abi_struct_sizes",
           "mmr_loss_scratch_bytes", "mmr_loss_fwd_bwd", "mmr_projector_fwd", "mmr_projector_bwd",
           "mmr_fusion_pack_weights", "mmr_route_fusion_fwd_packed",
           "mmr_producer_proj_sizes", "mmr_producer_proj_fwd", "mmr_producer_proj_bwd",
           "mmr_routing_stats_accumulate", "mmr_attention_fwd", "mmr_attention_bwd"]


def lib_path() -> str:
    """csrc/libmmr_b200.so; MMR_B200_LIB points at another build of the same sources (tuning experiments)."""
    return os.environ.get("MMR_B200_LIB") or _build.LIB


def load():
    """Loads csrc/libmmr_b200.so.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the sm_100a kernels first (python -m multimodalrouting_b200.build "
            "or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(path)
    lib.mmr_version.restype = C.c_int
    lib.mmr_last_error_string.restype = C.c_char_p
    lib.mmr_fusion_num_params.argtypes = [C.POINTER(FusionDims)]
    lib.mmr_fusion_num_params.restype = C.c_int
    lib.mmr_fusion_sizes.argtypes = [C.POINTER(FusionDims)] + [C.POINTER(C.c_size_t)] * 4
    lib.mmr_fusion_sizes.restype = C.c_int
    lib.mmr_route_fusion_fwd.argtypes = [C.POINTER(FusionDims), C.POINTER(c_fp)] + [c_fp] * 12
    lib.mmr_route_fusion_fwd.restype = C.c_int
    lib.mmr_route_fusion_fwd_packed.argtypes = [C.POINTER(FusionDims), C.POINTER(c_fp)] + [c_fp] * 12
    lib.mmr_route_fusion_fwd_packed.restype = C.c_int
    lib.mmr_fusion_pack_weights.argtypes = [C.POINTER(FusionDims), C.POINTER(c_fp), c_fp, c_fp]
    lib.mmr_fusion_pack_weights.restype = C.c_int
    lib.mmr_route_fusion_bwd.argtypes = ([C.POINTER(FusionDims), C.POINTER(c_fp)] + [c_fp] * 10 +
                                         [C.POINTER(c_fp)] + [c_fp] * 4)
    lib.mmr_route_fusion_bwd.restype = C.c_int
    lib.mmr_route_fusion_bwd_events.argtypes = ([C.POINTER(FusionDims), C.POINTER(c_fp)] + [c_fp] * 10 +
                                                [C.POINTER(c_fp)] + [c_fp] * 4 + [C.POINTER(c_fp)])
    lib.mmr_route_fusion_bwd_events.restype = C.c_int
    lib.mmr_route_fusion_bwd_ex.argtypes = ([C.POINTER(FusionDims), C.POINTER(c_fp)] + [c_fp] * 10 +
                                            [C.POINTER(c_fp)] + [c_fp] * 4 + [C.POINTER(c_fp), c_fp, C.POINTER(c_fp)])
    lib.mmr_route_fusion_bwd_ex.restype = C.c_int
    lib.mmr_routing_scratch_bytes.argtypes = [C.POINTER(RoutingDims)]
    lib.mmr_routing_scratch_bytes.restype = C.c_size_t
    lib.mmr_routing_pack_weights.argtypes = [C.POINTER(RoutingParams), C.c_int, c_fp, c_fp, c_fp, c_fp]
    lib.mmr_routing_pack_weights.restype = C.c_int
    lib.mmr_capsule_routing_fwd.argtypes = [C.POINTER(RoutingDims), C.POINTER(RoutingParams)] + [c_fp] * 11
    lib.mmr_capsule_routing_fwd.restype = C.c_int
    lib.mmr_capsule_routing_fwd_ex.argtypes = [C.POINTER(RoutingDims), C.POINTER(RoutingParams)] + [c_fp] * 12
    lib.mmr_capsule_routing_fwd_ex.restype = C.c_int
    lib.mmr_routing_fwd_scratch_bytes.argtypes = [C.POINTER(RoutingDims)]
    lib.mmr_routing_fwd_scratch_bytes.restype = C.c_size_t
    lib.mmr_capsule_routing_bwd.argtypes = ([C.POINTER(RoutingDims), C.POINTER(RoutingParams)] + [c_fp] * 8 +
                                            [C.POINTER(RoutingGrads)] + [c_fp] * 4)
    lib.mmr_capsule_routing_bwd.restype = C.c_int
    lib.mmr_capsule_routing_bwd_ex.argtypes = ([C.POINTER(RoutingDims), C.POINTER(RoutingParams)] + [c_fp] * 8 +
                                               [C.POINTER(RoutingGrads)] + [c_fp] * 5)
    lib.mmr_capsule_routing_bwd_ex.restype = C.c_int
    lib.mmr_projector_fwd.argtypes = [C.POINTER(RoutingParams), c_fp, C.c_int64, C.c_int64, C.c_int, c_fp, c_fp, c_fp]
    lib.mmr_projector_fwd.restype = C.c_int
    lib.mmr_projector_bwd.argtypes = [C.POINTER(RoutingParams), c_fp, C.c_int64, C.c_int64, C.c_int, c_fp, c_fp, c_fp,
                                      C.POINTER(RoutingGrads), c_fp, c_fp]
    lib.mmr_projector_bwd.restype = C.c_int
    lib.mmr_producer_proj_sizes.argtypes = [C.POINTER(ProjDims)] + [C.POINTER(C.c_size_t)] * 3
    lib.mmr_producer_proj_sizes.restype = C.c_int
    lib.mmr_producer_proj_fwd.argtypes = [C.POINTER(ProjDims)] + [c_fp] * 9
    lib.mmr_producer_proj_fwd.restype = C.c_int
    lib.mmr_producer_proj_bwd.argtypes = [C.POINTER(ProjDims)] + [c_fp] * 12
    lib.mmr_producer_proj_bwd.restype = C.c_int
    lib.mmr_routing_stats_accumulate.argtypes = [c_fp, C.c_int, c_fp, c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp]
    lib.mmr_routing_stats_accumulate.restype = C.c_int
    lib.mmr_attention_fwd.argtypes = [C.c_int] * 4 + [c_fp] * 6
    lib.mmr_attention_fwd.restype = C.c_int
    lib.mmr_attention_bwd.argtypes = [C.c_int] * 4 + [c_fp] * 10
    lib.mmr_attention_bwd.restype = C.c_int
    lib.mmr_debug_gemm.argtypes = [C.c_int] * 6 + [c_fp] * 5
    lib.mmr_debug_gemm.restype = C.c_int
    lib.mmr_bench_gemm.argtypes = [C.c_int] * 4 + [c_fp] * 5 + [C.c_int, C.POINTER(C.c_float), c_fp]
    lib.mmr_bench_gemm.restype = C.c_int
    lib.mmr_launch_count.restype = C.c_longlong
    lib.mmr_prof_enable.argtypes = [C.c_int]
    lib.mmr_prof_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.mmr_sanitize_rows_fwd.argtypes = [c_fp, C.c_int, c_fp, C.c_int64, C.c_int, C.c_int, C.c_float, c_fp, c_fp]
    lib.mmr_sanitize_rows_fwd.restype = C.c_int
    lib.mmr_sanitize_rows_bwd.argtypes = [c_fp, C.c_int, c_fp, c_fp, C.c_int64, C.c_int, C.c_int, C.c_float, c_fp]
    lib.mmr_sanitize_rows_bwd.restype = C.c_int
    lib.mmr_route_mask_from_presence.argtypes = [c_fp, c_fp, c_fp, C.c_int, C.c_int, c_fp, c_fp]
    lib.mmr_route_mask_from_presence.restype = C.c_int
    lib.mmr_grad_sqnorm.argtypes = [C.POINTER(OptTensor), C.c_int, c_fp, c_fp]
    lib.mmr_grad_sqnorm.restype = C.c_int
    lib.mmr_opt_prepare.argtypes = [C.POINTER(OptHyper), c_fp, c_fp]
    lib.mmr_opt_prepare.restype = C.c_int
    lib.mmr_opt_apply.argtypes = [C.POINTER(OptTensor), C.c_int, C.POINTER(OptHyper), c_fp, c_fp]
    lib.mmr_opt_apply.restype = C.c_int
    lib.mmr_ema_update.argtypes = [C.POINTER(OptTensor), C.c_int, C.c_double, c_fp]
    lib.mmr_ema_update.restype = C.c_int
    lib.mmr_abi_struct_sizes.argtypes = [C.POINTER(C.c_size_t), C.c_int]
    lib.mmr_abi_struct_sizes.restype = C.c_int
    # the ctypes mirrors above must match the structs this build was compiled with
    lib.mmr_loss_scratch_bytes.argtypes = [C.c_int]
    lib.mmr_loss_scratch_bytes.restype = C.c_size_t
    lib.mmr_loss_fwd_bwd.argtypes = [C.POINTER(LossArgs), c_fp]
    lib.mmr_loss_fwd_bwd.restype = C.c_int
    sizes = (C.c_size_t * 10)()
    n = lib.mmr_abi_struct_sizes(sizes, 10)
    mirrors = (FusionDims, RoutingDims, RoutingParams, RoutingGrads, OptTensor, OptHyper)
    for i, cls in enumerate(mirrors[:n]):
        if C.sizeof(cls) != sizes[i]:
            raise RuntimeError(f"ABI mismatch: ctypes {cls.__name__} is {C.sizeof(cls)} bytes, {path} has {sizes[i]} "
                               "(stale libmmr_b200.so? rebuild with python -m multimodalrouting_b200.build)")
    if n >= 7 and sizes[6] != OPT_STATE_BYTES:
        raise RuntimeError(f"ABI mismatch: mmr_opt_state is {sizes[6]} bytes, binding expects {OPT_STATE_BYTES}")
    if n < 9 or sizes[7] != LOSS_STATE_BYTES or sizes[8] != C.sizeof(LossArgs):
        raise RuntimeError(f"ABI mismatch: loss structs are {list(sizes[7:n])} bytes, binding expects "
                           f"[{LOSS_STATE_BYTES}, {C.sizeof(LossArgs)}] (stale libmmr_b200.so?)")
    if n < 10 or sizes[9] != C.sizeof(ProjDims):
        raise RuntimeError(f"ABI mismatch: mmr_proj_dims is {sizes[9] if n >= 10 else 'absent'} bytes, binding expects "
                           f"{C.sizeof(ProjDims)} (stale libmmr_b200.so?)")
    _LIB = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mmr_last_error_string().decode("utf-8", "replace")
        if rc in (1, 2):
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg}")
