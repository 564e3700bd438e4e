"""torch.library custom ops + autograd glue over the C ABI (include/mmr_b200.h).

PyTorch is plumbing here: it owns device memory, streams and autograd bookkeeping; all arithmetic
of the hot path runs in the hand-written sm_100a kernels of csrc/.  There is no fallback path: if
the shared library is missing or the tensors are not on a CUDA device the ops raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch._subclasses.fake_tensor import FakeTensor

from . import _lib
from ._lib import FusionDims, RoutingDims, RoutingGrads, RoutingParams, c_fp

DTYPE_F32, DTYPE_BF16 = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TC = 0, 1, 2
VARIANT = {"mort": 0, "pheno": 1}
N_ROUTES = 10


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "multimodalrouting_b200 runs only on a CUDA (B200, sm_100a) device; got a "
                f"{t.device} tensor.  There is no CPU fallback for the route-fusion path.")


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def resolve_dtype(mode: str = "auto") -> int:
    """fp32 unless running under torch.autocast (any reduced dtype -> bf16 kernels), like the
    reference's amp contexts (M/main.py:2633-2659).  MMR_B200_DTYPE=fp32|bf16 overrides."""
    mode = os.environ.get("MMR_B200_DTYPE", mode)
    if mode in ("bf16", "bfloat16"):
        return DTYPE_BF16
    if mode in ("fp32", "float32"):
        return DTYPE_F32
    if torch.is_autocast_enabled("cuda"):
        return DTYPE_BF16
    return DTYPE_F32


def resolve_engine() -> int:
    return {"auto": GEMM_AUTO, "simt": GEMM_SIMT, "tc": GEMM_TC}[os.environ.get("MMR_B200_GEMM", "auto")]


def _fusion_dims(x_l, x_n, x_i, layers, dtype, engine) -> FusionDims:
    B, TL, dL = x_l.shape
    _, TN, dN = x_n.shape
    _, TI, dI = x_i.shape
    return FusionDims(B, TL, TN, TI, dL, dN, dI, layers, dtype, engine)


def fusion_sizes(dims: FusionDims) -> Tuple[int, int, int, int]:
    lib = _lib.load()
    s = [C.c_size_t() for _ in range(4)]
    _lib.check(lib.mmr_fusion_sizes(C.byref(dims), *[C.byref(v) for v in s]), "mmr_fusion_sizes")
    return tuple(int(v.value) for v in s)


def _ptr_table(ts: Sequence[Optional[Tensor]]):
    arr = (c_fp * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = None if t is None else t.data_ptr()
    return arr


# --------------------------------------------------------------------------------------------
# raw ops (opaque to torch.compile, with fake impls)
def _pack_dims(layers: int, dtype: int, engine: int) -> FusionDims:
    # the packed-weight layout depends on the depth and the compute dtype only
    return FusionDims(1, 1, 1, 1, 256, 256, 256, layers, dtype, engine)


@torch.library.custom_op("mmr_b200::route_fusion_pack", mutates_args=())
def route_fusion_pack(params: Sequence[Tensor], layers: int, dtype: int, engine: int) -> Tensor:
    """Compute-type copies of the GEMM weights (forward and transposed, query scaling and the K/V LayerNorm affine folded
    in) that the forward and backward kernels read: mmr_fusion_pack_weights.  Valid for one parameter version."""
    _require_cuda(*params)
    lib = _lib.load()
    dims = _pack_dims(layers, dtype, engine)
    n_expected = lib.mmr_fusion_num_params(C.byref(dims))
    if n_expected != len(params):
        raise ValueError(f"route_fusion_pack expects {n_expected} parameter tensors, got {len(params)}")
    packed = torch.empty(fusion_sizes(dims)[0], dtype=torch.uint8, device=params[0].device)
    _lib.check(lib.mmr_fusion_pack_weights(C.byref(dims), _ptr_table(params), _ptr(packed), _stream()),
               "mmr_fusion_pack_weights")
    return packed


@route_fusion_pack.register_fake
def _(params, layers, dtype, engine):
    return params[0].new_empty(fusion_sizes(_pack_dims(layers, dtype, engine))[0], dtype=torch.uint8)


@torch.library.custom_op("mmr_b200::route_fusion_fwd", mutates_args=())
def route_fusion_fwd(x_l: Tensor, x_n: Tensor, x_i: Tensor, mL: Optional[Tensor], mN: Optional[Tensor],
                     mI: Optional[Tensor], pos: Tensor, params: Sequence[Tensor], packed: Tensor, layers: int,
                     dtype: int, engine: int) -> Tuple[Tensor, Tensor]:
    """MULTModel.forward from packed weights (route_fusion_pack).  Returns (routes [10,B,256], saved-for-backward)."""
    _require_cuda(x_l, x_n, x_i, pos, packed, *params)
    lib = _lib.load()
    dims = _fusion_dims(x_l, x_n, x_i, layers, dtype, engine)
    n_expected = lib.mmr_fusion_num_params(C.byref(dims))
    if n_expected != len(params):
        raise ValueError(f"route_fusion_fwd expects {n_expected} parameter tensors, got {len(params)}")
    packed_b, saved_b, sf_b, _ = fusion_sizes(dims)
    if packed.numel() != packed_b or packed.dtype != torch.uint8:
        raise ValueError(f"route_fusion_fwd: packed weights must be {packed_b} bytes (route_fusion_pack), got {packed.numel()}")
    dev = x_l.device
    saved = torch.empty(saved_b, dtype=torch.uint8, device=dev)
    scratch = torch.empty(sf_b, dtype=torch.uint8, device=dev)
    routes = torch.empty(N_ROUTES, x_l.shape[0], 256, dtype=torch.float32, device=dev)
    rc = lib.mmr_route_fusion_fwd_packed(C.byref(dims), _ptr_table(params), _ptr(x_l), _ptr(x_n), _ptr(x_i), _ptr(mL),
                                         _ptr(mN), _ptr(mI), _ptr(pos), _ptr(packed), _ptr(saved), _ptr(scratch),
                                         _ptr(routes), _stream())
    _lib.check(rc, "mmr_route_fusion_fwd_packed")
    return routes, saved


@route_fusion_fwd.register_fake
def _(x_l, x_n, x_i, mL, mN, mI, pos, params, packed, layers, dtype, engine):
    B = x_l.shape[0]
    # same metadata as the real op: the byte counts come from the (host-only) planner of the C ABI
    _, saved_b, _, _ = fusion_sizes(_fusion_dims(x_l, x_n, x_i, layers, dtype, engine))
    return (x_l.new_empty(N_ROUTES, B, 256, dtype=torch.float32), x_l.new_empty(saved_b, dtype=torch.uint8))


_GRAD_LAYOUTS = {}
# per-layer parameter slots (include/mmr_b200.h order) whose gradients are final as soon as the backward of
# their layer has run: out_proj.{weight,bias}, fc1.{weight,bias}, fc2.{weight,bias}, layer_norms.1.{weight,bias}
_EARLY_SLOTS = (2, 3, 4, 5, 6, 7, 10, 11)
_LATE_SLOTS = (0, 1, 8, 9)          # in_proj_{weight,bias}, layer_norms.0.*: finalised by the last kernels


def grad_layout(shapes: Sequence[Tuple[int, ...]], layers: int):
    """Placement of the parameter gradients inside the flat buffer returned by route_fusion_bwd.

    * gradients that are final once layer l's backward has run are stacked per (layer, kind) over the six
      encoders and laid out layer by layer in COMPLETION order (last layer first): a data-parallel caller can
      all-reduce block l while layers l-1 .. 0 are still being differentiated (`buckets`);
    * the late kinds are stacked over all 6 x layers instances; everything else sits in 16-byte aligned slots;
    * stacking lets the autograd side hand the views out with one ``unbind`` per stack instead of two tensor
      ops per parameter.
    Returns (offsets per parameter, total floats, groups = [(offset, [param indices], shape)],
             buckets = [(start, end)] * layers in completion order + [(start of the rest, total)])."""
    key = (tuple(shapes), layers)
    hit = _GRAD_LAYOUTS.get(key)
    if hit is not None:
        return hit
    n = len(shapes)
    numel = [int(torch.Size(sh).numel()) for sh in shapes]
    offs = [-1] * n
    groups, buckets = [], []
    o = rest = 0
    per_enc = layers * 12 + 2
    if n == 9 + 6 * per_enc + 8:        # MULTModel state_dict order (include/mmr_b200.h)
        stackable = all(numel[9 + s] % 4 == 0 for s in range(12))
        if stackable:
            for l in range(layers - 1, -1, -1):
                start = o
                for slot in _EARLY_SLOTS:
                    idx = [9 + d * per_enc + l * 12 + slot for d in range(6)]
                    groups.append((o, idx, tuple(shapes[idx[0]])))
                    for i in idx:
                        offs[i] = o
                        o += numel[i]
                buckets.append((start, o))
            rest = o                    # everything from here on is final only when the backward ends
            for slot in _LATE_SLOTS:
                idx = [9 + d * per_enc + l * 12 + slot for d in range(6) for l in range(layers)]
                groups.append((o, idx, tuple(shapes[idx[0]])))
                for i in idx:
                    offs[i] = o
                    o += numel[i]
    for i in range(n):
        if offs[i] < 0:
            offs[i] = o
            o += (numel[i] + 3) // 4 * 4
    buckets.append((rest, o))
    out = (offs, o, groups, buckets)
    _GRAD_LAYOUTS[key] = out
    return out


# Data-parallel overlap (multimodalrouting_b200.dist.OverlappedGradReducer): when set, the backward asks the
# library to record `events[l]` after layer l's early gradients and then calls hook(flat, buckets).
_OVERLAP = {"hook": None, "events": None}


def set_grad_overlap(hook, events) -> None:
    """One reducer per process: the hook is consulted by every RouteFusionFn backward.  Installing a second one while the
    first is active would silently redirect the first module's gradients, so that raises (close() the old one first)."""
    if hook is not None and _OVERLAP["hook"] is not None and _OVERLAP["hook"] != hook:
        raise RuntimeError("a gradient-overlap hook is already installed (another OverlappedGradReducer is active); "
                           "close() it before creating a new one")
    _OVERLAP["hook"], _OVERLAP["events"] = hook, events


# Second stream for the weight-gradient kernels of the backward (mmr_route_fusion_bwd_ex): one side stream and nine
# ordering events per device, created once (outside any CUDA-graph capture) and owned here, not by the library.
N_BWD_SYNC_EVENTS = 9
_WGRAD_SIDE = {}


def wgrad_side_stream(device):
    """(side stream, ctypes event table) for `device`, or (None, None) when MMR_WGRAD_STREAM=0."""
    if os.environ.get("MMR_WGRAD_STREAM", "1") == "0":
        return None, None
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    ent = _WGRAD_SIDE.get(key)
    if ent is None:
        if torch.cuda.is_current_stream_capturing():
            return None, None          # cannot create / synchronise during a capture: this call stays single-stream
        with torch.cuda.device(key):
            stream = torch.cuda.Stream()
            events = [torch.cuda.Event() for _ in range(N_BWD_SYNC_EVENTS)]
            for e in events:           # CUDA events are created lazily: force the handles to exist
                e.record(stream)
            stream.synchronize()
        table = (c_fp * N_BWD_SYNC_EVENTS)(*[int(e.cuda_event) for e in events])
        ent = (stream, events, table)
        _WGRAD_SIDE[key] = ent
    return ent[0], ent[2]


@torch.library.custom_op("mmr_b200::route_fusion_bwd", mutates_args=())
def route_fusion_bwd(x_l: Tensor, x_n: Tensor, x_i: Tensor, mL: Optional[Tensor], mN: Optional[Tensor],
                     mI: Optional[Tensor], params: Sequence[Tensor], packed: Tensor, saved: Tensor,
                     d_routes: Tensor, need: Sequence[bool], layers: int, dtype: int,
                     engine: int, events: Sequence[int]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Returns (dx_l, dx_n, dx_i, flat parameter gradients laid out by `grad_layout`)."""
    lib = _lib.load()
    dims = _fusion_dims(x_l, x_n, x_i, layers, dtype, engine)
    _, _, _, sb_b = fusion_sizes(dims)
    dev = x_l.device
    scratch = torch.empty(sb_b, dtype=torch.uint8, device=dev)
    offs, o, _, _ = grad_layout([tuple(p.shape) for p in params], layers)   # 16-byte aligned slots
    flat = torch.zeros(o, dtype=torch.float32, device=dev)
    base = flat.data_ptr()
    grads = (c_fp * len(params))()
    for i, p in enumerate(params):
        grads[i] = (base + 4 * offs[i]) if need[i] else None
    dx = [torch.empty_like(x) for x in (x_l, x_n, x_i)]
    evs = None
    if len(events) > 0:
        evs = (c_fp * layers)()
        for l in range(layers):
            evs[l] = events[l] if l < len(events) and events[l] else None
    side, sync = (None, None) if evs is not None else wgrad_side_stream(dev)
    # (the side stream is forked from and joined back into the current stream inside the call, so every buffer is
    # released in current-stream order: no record_stream needed, and the call stays CUDA-graph capturable)
    rc = lib.mmr_route_fusion_bwd_ex(C.byref(dims), _ptr_table(params), _ptr(x_l), _ptr(x_n), _ptr(x_i), _ptr(mL),
                                     _ptr(mN), _ptr(mI), _ptr(packed), _ptr(saved), _ptr(scratch), _ptr(d_routes),
                                     grads, _ptr(dx[0]), _ptr(dx[1]), _ptr(dx[2]), _stream(), evs,
                                     side.cuda_stream if side is not None else None, sync)
    _lib.check(rc, "mmr_route_fusion_bwd")
    return dx[0], dx[1], dx[2], flat


@route_fusion_bwd.register_fake
def _(x_l, x_n, x_i, mL, mN, mI, params, packed, saved, d_routes, need, layers, dtype, engine, events):
    _, n, _, _ = grad_layout([tuple(p.shape) for p in params], layers)
    return (torch.empty_like(x_l), torch.empty_like(x_n), torch.empty_like(x_i),
            x_l.new_empty(n, dtype=torch.float32))


def _common_base(ts: Sequence[Optional[Tensor]], shape) -> Optional[Tuple[int, int, int]]:
    """If the tensors are equally spaced fp32 [B,256] views of one buffer, return
    (base_ptr, route_stride_elems, batch_stride_elems)."""
    t0 = ts[0]
    if any(t is None or t.dtype != torch.float32 or tuple(t.shape) != tuple(shape) for t in ts):
        return None
    if any(isinstance(t, FakeTensor) for t in ts):
        return None     # tracing: fake tensors have no addresses; the caller falls back to the dense copy
    if t0.stride(1) != 1 or any(t.stride() != t0.stride() for t in ts):
        return None
    st0 = t0.untyped_storage().data_ptr()
    if any(t.untyped_storage().data_ptr() != st0 for t in ts):
        return None     # separately allocated tensors may be equally spaced by accident
    p0 = t0.data_ptr()
    if len(ts) == 1:
        return p0, 0, t0.stride(0)
    step = ts[1].data_ptr() - p0
    if step <= 0 or step % 4 or any(t.data_ptr() - p0 != i * step for i, t in enumerate(ts)):
        return None
    return p0, step // 4, t0.stride(0)


class RouteFusionFn(torch.autograd.Function):
    """MULTModel.forward as one autograd node: 3 sequences + masks + 317 parameters -> 10 routes."""

    @staticmethod
    def forward(ctx, x_l, x_n, x_i, mL, mN, mI, pos, layers, dtype, engine, packed, *params):
        xs = [_f32c(x.detach()) for x in (x_l, x_n, x_i)]
        ms = [_f32c(m.detach()) if m is not None else None for m in (mL, mN, mI)]
        ps = [p.detach() for p in params]
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("route fusion parameters must be contiguous fp32 tensors")
        if packed is None:
            packed = route_fusion_pack(ps, layers, dtype, engine)
        routes, saved = route_fusion_fwd(xs[0], xs[1], xs[2], ms[0], ms[1], ms[2], pos, ps, packed, layers, dtype, engine)
        ctx.save_for_backward(*xs, *[m for m in ms if m is not None], packed, saved, *ps)
        ctx.mask_present = [m is not None for m in ms]
        ctx.cfg = (layers, dtype, engine)
        ctx.n_params = len(ps)
        outs = tuple(routes[r] for r in range(N_ROUTES))
        return outs

    @staticmethod
    def backward(ctx, *d_routes):
        sv = list(ctx.saved_tensors)
        xs = sv[:3]
        k = 3
        ms = []
        for present in ctx.mask_present:
            if present:
                ms.append(sv[k]); k += 1
            else:
                ms.append(None)
        packed, saved = sv[k], sv[k + 1]
        ps = sv[k + 2:]
        layers, dtype, engine = ctx.cfg
        B = xs[0].shape[0]
        base = _common_base(d_routes, (B, 256))
        if base is not None and base[1] == B * 256 and base[2] == 256:
            d_all = torch.as_strided(d_routes[0], (N_ROUTES, B, 256), (B * 256, 256, 1))
        else:
            d_all = torch.stack([g.float() if g is not None else torch.zeros(B, 256, device=xs[0].device)
                                 for g in d_routes], dim=0).contiguous()
        need = [bool(n) for n in ctx.needs_input_grad[11:]]
        hook, events = _OVERLAP["hook"], _OVERLAP["events"]
        handles = [int(e.cuda_event) for e in events[:layers]] if (hook is not None and events) else []
        dxl, dxn, dxi, flat = route_fusion_bwd(xs[0], xs[1], xs[2], ms[0], ms[1], ms[2], ps, packed, saved, d_all,
                                               need, layers, dtype, engine, handles)
        offs, _, groups, buckets = grad_layout([tuple(p.shape) for p in ps], layers)
        if hook is not None:
            hook(flat, buckets)
        grads = [None] * len(ps)
        grouped = set()
        for start, idx, sh in groups:           # one unbind per parameter kind
            cnt, numel = len(idx), ps[idx[0]].numel()
            views = flat[start:start + cnt * numel].view(cnt, *sh).unbind(0)
            for i, v in zip(idx, views):
                grouped.add(i)
                if need[i]:
                    grads[i] = v
        for i, p in enumerate(ps):
            if i not in grouped and need[i]:
                grads[i] = flat[offs[i]:offs[i] + p.numel()].view(p.shape)
        gi = ctx.needs_input_grad
        return (dxl if gi[0] else None, dxn if gi[1] else None, dxi if gi[2] else None,
                None, None, None, None, None, None, None, None, *grads)


def route_fusion(x_l, x_n, x_i, mL, mN, mI, pos, params: List[Tensor], layers: int, dtype: int,
                 engine: int, packed: Optional[Tensor] = None) -> Tuple[Tensor, ...]:
    """`packed`: output of route_fusion_pack for the CURRENT parameter values, or None to pack inside this call."""
    _require_cuda(x_l, x_n, x_i)
    return RouteFusionFn.apply(x_l, x_n, x_i, mL, mN, mI, pos, layers, dtype, engine, packed, *params)


# --------------------------------------------------------------------------------------------
# capsule routing
def _routing_dims(B, K, variant, num_routing, detach_priors, from_poses, temp, floor, ceil, rs, bs, vdt) -> RoutingDims:
    return RoutingDims(B, K, variant, num_routing, int(detach_priors), int(from_poses), float(temp), float(floor),
                       float(ceil), int(rs), int(bs), int(vdt), 0)


PROJ_PACK_BYTES = 10 * 40 * 256 * 2


def routing_pack_bytes(K: int) -> int:
    """bytes of the fp16 weight copies the tensor-core routing paths read (caps_wt | caps_w16 | proj_wb)."""
    return 2 * (10 * K * 64 * 32 * 2) + PROJ_PACK_BYTES


def _pack_bytes_aligned(K: int) -> int:
    return (routing_pack_bytes(K) + 255) // 256 * 256


def routing_state_bytes(dims: RoutingDims) -> int:
    """bytes of the buffer the forward hands to the backward in the reduced-precision mode: the fp16 weight copies followed by
    the forward scratch of the split path (projector outputs, head matrix, fp16 votes; csrc/routing_split.cuh)."""
    return _pack_bytes_aligned(dims.K) + int(_lib.load().mmr_routing_fwd_scratch_bytes(C.byref(dims)))


def _routing_params(proj_w, proj_b, caps_w, pose_to_mc, embedding, bias, packed=None) -> RoutingParams:
    rp = RoutingParams()
    if packed is not None and packed.numel() > 0:
        K = embedding.shape[0]
        caps = 10 * K * 64 * 32 * 2
        rp.caps_wt_f16 = packed.data_ptr()
        rp.caps_w_f16 = packed.data_ptr() + caps
        rp.proj_w_f16 = (packed.data_ptr() + 2 * caps) if proj_w else None
    for r in range(N_ROUTES):
        rp.proj_w[r] = proj_w[r].data_ptr() if proj_w else None
        rp.proj_b[r] = proj_b[r].data_ptr() if proj_b else None
    rp.caps_w = caps_w.data_ptr()
    rp.pose_to_mc = pose_to_mc.data_ptr()
    rp.embedding = embedding.data_ptr()
    rp.bias = bias.data_ptr()
    return rp


@torch.library.custom_op("mmr_b200::capsule_routing_fwd", mutates_args=())
def capsule_routing_fwd(embs: Optional[Tensor], rs: int, bs: int, poses_in: Optional[Tensor],
                        acts_in: Optional[Tensor], acts_override: Optional[Tensor], route_mask: Optional[Tensor],
                        proj_w: Sequence[Tensor], proj_b: Sequence[Tensor], caps_w: Tensor, pose_to_mc: Tensor,
                        embedding: Tensor, bias: Tensor, B: int, variant: int, num_routing: int,
                        detach_priors: bool, temp: float, floor: float, ceil: float, vdt: int
                        ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Returns (logits [B,K], alpha [B,10], R [B,10,K], poses [B,10,32], acts [B,10], packed fp16 weights | empty)."""
    _require_cuda(caps_w, embs, poses_in)
    lib = _lib.load()
    K = embedding.shape[0]
    from_poses = embs is None
    dims = _routing_dims(B, K, variant, num_routing, detach_priors, from_poses, temp, floor, ceil, rs, bs, vdt)
    dev = caps_w.device
    # reduced-precision mode: projector and vote contraction run on tensor cores from fp16 weight copies
    use_tc = vdt == DTYPE_BF16 and os.environ.get("MMR_RT_TC", "1") != "0"
    packed = torch.empty(routing_state_bytes(dims) if use_tc else 0, dtype=torch.uint8, device=dev)
    rp = _routing_params(proj_w, proj_b, caps_w, pose_to_mc, embedding, bias, packed)
    if use_tc:
        rc = lib.mmr_routing_pack_weights(C.byref(rp), K, rp.caps_wt_f16, rp.caps_w_f16, rp.proj_w_f16, _stream())
        _lib.check(rc, "mmr_routing_pack_weights")
    logits = torch.empty(B, K, dtype=torch.float32, device=dev)
    alpha = torch.empty(B, N_ROUTES, dtype=torch.float32, device=dev)
    R = torch.empty(B, N_ROUTES, K, dtype=torch.float32, device=dev)
    poses = torch.empty(B, N_ROUTES, 32, dtype=torch.float32, device=dev)
    acts = torch.empty(B, N_ROUTES, dtype=torch.float32, device=dev)
    # the split path (csrc/routing_split.cuh) stages the projector outputs and the fp16 votes of the batch behind the packed
    # weights; the backward reads them from there instead of recomputing them
    scratch = (packed.data_ptr() + _pack_bytes_aligned(K)) if use_tc else None
    rc = lib.mmr_capsule_routing_fwd_ex(C.byref(dims), C.byref(rp), _ptr(embs), _ptr(poses_in), _ptr(acts_in),
                                        _ptr(acts_override), _ptr(route_mask), _ptr(logits), _ptr(alpha), _ptr(R),
                                        None if from_poses else _ptr(poses), None if from_poses else _ptr(acts),
                                        scratch, _stream())
    _lib.check(rc, "mmr_capsule_routing_fwd_ex")
    return logits, alpha, R, poses, acts, packed


@capsule_routing_fwd.register_fake
def _(embs, rs, bs, poses_in, acts_in, acts_override, route_mask, proj_w, proj_b, caps_w, pose_to_mc, embedding,
      bias, B, variant, num_routing, detach_priors, temp, floor, ceil, vdt):
    K = embedding.shape[0]
    e = caps_w
    use_tc = vdt == DTYPE_BF16 and os.environ.get("MMR_RT_TC", "1") != "0"
    dims = _routing_dims(B, K, variant, num_routing, detach_priors, embs is None, temp, floor, ceil, rs, bs, vdt)
    return (e.new_empty(B, K), e.new_empty(B, N_ROUTES), e.new_empty(B, N_ROUTES, K),
            e.new_empty(B, N_ROUTES, 32), e.new_empty(B, N_ROUTES),
            e.new_empty(routing_state_bytes(dims) if use_tc else 0, dtype=torch.uint8))


@torch.library.custom_op("mmr_b200::capsule_routing_bwd", mutates_args=())
def capsule_routing_bwd(embs: Optional[Tensor], rs: int, bs: int, poses_in: Optional[Tensor],
                        acts_in: Optional[Tensor], acts_override: Optional[Tensor], route_mask: Optional[Tensor],
                        proj_w: Sequence[Tensor], proj_b: Sequence[Tensor], caps_w: Tensor, pose_to_mc: Tensor,
                        embedding: Tensor, bias: Tensor, d_logits: Tensor, d_R: Optional[Tensor], B: int,
                        variant: int, num_routing: int, detach_priors: bool, temp: float, floor: float,
                        ceil: float, vdt: int, packed: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Returns (d_embs [10,B,256] | empty, d_poses [B,10,32] | empty, d_acts [B,10] | empty, flat grads)
    flat grads layout: proj_w[10] (33*256 each) | proj_b[10] (36 each, 33 used) | caps_w | pose_to_mc |
    embedding | bias."""
    lib = _lib.load()
    K = embedding.shape[0]
    from_poses = embs is None
    dims = _routing_dims(B, K, variant, num_routing, detach_priors, from_poses, temp, floor, ceil, rs, bs, vdt)
    rp = _routing_params(proj_w, proj_b, caps_w, pose_to_mc, embedding, bias, packed)
    dev = caps_w.device
    scratch = torch.empty(int(lib.mmr_routing_scratch_bytes(C.byref(dims))), dtype=torch.uint8, device=dev)
    n_flat = routing_flat_layout(K)["total"]
    flat = torch.zeros(n_flat, dtype=torch.float32, device=dev)
    lay = routing_flat_layout(K)
    base = flat.data_ptr()
    g = RoutingGrads()
    for r in range(N_ROUTES):
        g.proj_w[r] = None if from_poses else base + 4 * (lay["proj_w"] + r * 33 * 256)
        g.proj_b[r] = None if from_poses else base + 4 * (lay["proj_b"] + r * 36)
    g.caps_w = base + 4 * lay["caps_w"]
    g.pose_to_mc = base + 4 * lay["pose_to_mc"]
    g.embedding = base + 4 * lay["embedding"]
    g.bias = base + 4 * lay["bias"]
    if from_poses:
        d_embs = torch.empty(0, dtype=torch.float32, device=dev)
        d_poses = torch.empty(B, N_ROUTES, 32, dtype=torch.float32, device=dev)
        d_acts = torch.empty(B, N_ROUTES, dtype=torch.float32, device=dev)
    else:
        d_embs = torch.empty(N_ROUTES, B, 256, dtype=torch.float32, device=dev)
        d_poses = torch.empty(0, dtype=torch.float32, device=dev)
        # gradient wrt acts_override (the prior chain ends there, routing_and_heads.py:314); zero for masked routes
        d_acts = torch.zeros(B if acts_override is not None else 0, N_ROUTES, dtype=torch.float32, device=dev)
    # d_route_embs shares the (route, batch) strides of the input embeddings in the C ABI; RoutingFn
    # normalises the inputs to the dense [10,B,256] layout, so the output is dense as well.
    fwd_state = None
    if packed is not None and packed.numel() >= routing_state_bytes(dims):
        fwd_state = packed.data_ptr() + _pack_bytes_aligned(K)      # forward scratch of the same call (split path)
    rc = lib.mmr_capsule_routing_bwd_ex(C.byref(dims), C.byref(rp), _ptr(embs), _ptr(poses_in), _ptr(acts_in),
                                        _ptr(acts_override), _ptr(route_mask), _ptr(d_logits), _ptr(d_R),
                                        _ptr(scratch), C.byref(g), None if from_poses else _ptr(d_embs),
                                        _ptr(d_poses) if from_poses else None,
                                        _ptr(d_acts) if (from_poses or acts_override is not None) else None,
                                        fwd_state, _stream())
    _lib.check(rc, "mmr_capsule_routing_bwd_ex")
    return d_embs, d_poses, d_acts, flat


@capsule_routing_bwd.register_fake
def _(embs, rs, bs, poses_in, acts_in, acts_override, route_mask, proj_w, proj_b, caps_w, pose_to_mc, embedding,
      bias, d_logits, d_R, B, variant, num_routing, detach_priors, temp, floor, ceil, vdt, packed=None):
    K = embedding.shape[0]
    e = caps_w
    n = routing_flat_layout(K)["total"]
    if embs is None:
        return e.new_empty(0), e.new_empty(B, N_ROUTES, 32), e.new_empty(B, N_ROUTES), e.new_empty(n)
    return (e.new_empty(N_ROUTES, B, 256), e.new_empty(0), e.new_empty(B if acts_override is not None else 0, N_ROUTES),
            e.new_empty(n))


def routing_flat_layout(K: int):
    lay, o = {}, 0
    lay["proj_w"] = o; o += N_ROUTES * 33 * 256
    lay["proj_b"] = o; o += N_ROUTES * 36
    lay["caps_w"] = o; o += N_ROUTES * 32 * K * 64
    lay["pose_to_mc"] = o; o += 64 * 32
    lay["embedding"] = o; o += K * 64
    lay["bias"] = o; o += (K + 3) // 4 * 4
    lay["total"] = o
    return lay


class RoutingFn(torch.autograd.Function):
    """forward_capsule_from_route_dict (embs given) or CapsuleMortalityHead.forward (poses given)."""

    @staticmethod
    def forward(ctx, cfg, acts_override, route_mask, caps_w, pose_to_mc, embedding, bias, *rest):
        variant, num_routing, detach_priors, temp, floor, ceil, from_poses = cfg
        vdt = resolve_dtype()      # votes in bf16 under autocast (like the reference's einsums), else fp32
        if from_poses:
            poses_in, acts_in = rest
            B = poses_in.shape[0]
            poses_c, acts_c = _f32c(poses_in.detach()), _f32c(acts_in.detach())
            embs_dense, proj_w, proj_b = None, [], []
            rs = bs = 0
        else:
            embs = rest[:N_ROUTES]
            proj_w = [p.detach() for p in rest[N_ROUTES:2 * N_ROUTES]]
            proj_b = [p.detach() for p in rest[2 * N_ROUTES:3 * N_ROUTES]]
            B = embs[0].shape[0]
            base = _common_base([e.detach() for e in embs], (B, 256))
            if base is not None and base[1] == B * 256 and base[2] == 256:
                embs_dense = torch.as_strided(embs[0].detach(), (N_ROUTES, B, 256), (B * 256, 256, 1))
            else:
                embs_dense = torch.stack([_f32c(e.detach()) for e in embs], dim=0)
            rs, bs = B * 256, 256
            poses_c = acts_c = None
        rm = _f32c(route_mask.detach()) if route_mask is not None else None
        ao = _f32c(acts_override.detach()) if acts_override is not None else None
        logits, alpha, R, poses, acts, packed = capsule_routing_fwd(
            embs_dense, rs, bs, poses_c, acts_c, ao, rm, proj_w, proj_b, caps_w.detach(), pose_to_mc.detach(),
            embedding.detach(), bias.detach(), B, variant, num_routing, detach_priors, temp, floor, ceil, vdt)
        ctx.cfg = cfg
        ctx.vdt = vdt
        ctx.B = B
        ctx.has = (embs_dense is not None, rm is not None, ao is not None)
        ctx.ao_shape = tuple(acts_override.shape) if acts_override is not None else None
        tensors = [t for t in (embs_dense, poses_c, acts_c, ao, rm) if t is not None]
        ctx.n_in = len(tensors)
        ctx.save_for_backward(*tensors, caps_w.detach(), pose_to_mc.detach(), embedding.detach(), bias.detach(),
                              *proj_w, *proj_b, packed)
        ctx.mark_non_differentiable(alpha, poses, acts)
        return logits, alpha, R, poses, acts

    @staticmethod
    def backward(ctx, d_logits, _d_alpha, d_R, _d_poses, _d_acts):
        variant, num_routing, detach_priors, temp, floor, ceil, from_poses = ctx.cfg
        sv = list(ctx.saved_tensors)
        has_embs, has_rm, has_ao = ctx.has
        k = 0
        embs_dense = poses_c = acts_c = ao = rm = None
        if has_embs:
            embs_dense = sv[k]; k += 1
        else:
            poses_c, acts_c = sv[k], sv[k + 1]; k += 2
        if has_ao:
            ao = sv[k]; k += 1
        if has_rm:
            rm = sv[k]; k += 1
        caps_w, pose_to_mc, embedding, bias = sv[k:k + 4]
        k += 4
        proj_w = sv[k:k + N_ROUTES] if has_embs else []
        proj_b = sv[k + N_ROUTES:k + 2 * N_ROUTES] if has_embs else []
        packed = sv[-1]
        B, K = ctx.B, embedding.shape[0]
        if d_logits is None:
            d_logits = torch.zeros(B, K, device=caps_w.device)
        rs, bs = (B * 256, 256) if has_embs else (0, 0)
        d_embs, d_poses, d_acts, flat = capsule_routing_bwd(
            embs_dense, rs, bs, poses_c, acts_c, ao, rm, proj_w, proj_b, caps_w, pose_to_mc, embedding, bias,
            _f32c(d_logits), _f32c(d_R) if d_R is not None else None, B, variant, num_routing, detach_priors, temp,
            floor, ceil, ctx.vdt, packed)
        lay = routing_flat_layout(K)
        g_caps = flat[lay["caps_w"]:lay["caps_w"] + caps_w.numel()].view(caps_w.shape)
        g_mc = flat[lay["pose_to_mc"]:lay["pose_to_mc"] + 64 * 32].view(64, 32)
        g_emb = flat[lay["embedding"]:lay["embedding"] + K * 64].view(K, 64)
        g_bias = flat[lay["bias"]:lay["bias"] + K]
        # Mort never consumes the priors (current_act = ones * mask, d = sum_r R pose: Mort routing_and_heads.py:208-265), so
        # the reference leaves acts_override.grad = None there; Pheno routes with alpha
        g_ao = None
        if has_ao and not from_poses and ctx.needs_input_grad[1] and variant == VARIANT["pheno"]:
            g_ao = d_acts.view(ctx.ao_shape)
        head = (None, g_ao, None, g_caps, g_mc, g_emb, g_bias)
        if from_poses:
            return head + (d_poses, d_acts)
        g_pw = [flat[lay["proj_w"] + r * 33 * 256: lay["proj_w"] + (r + 1) * 33 * 256].view(33, 256)
                for r in range(N_ROUTES)]
        g_pb = [flat[lay["proj_b"] + r * 36: lay["proj_b"] + r * 36 + 33] for r in range(N_ROUTES)]
        return head + tuple(d_embs[r] for r in range(N_ROUTES)) + tuple(g_pw) + tuple(g_pb)


# --------------------------------------------------------------------------------------------
# standalone attention core (csrc/api.cu: mmr_attention_fwd / bwd) -- used by partial_fusion.py
@torch.library.custom_op("mmr_b200::attention_fwd", mutates_args=())
def attention_fwd(q: Tensor, kv: Tensor, kmask: Optional[Tensor], B: int, Tq: int, Tk: int, dtype: int) -> Tuple[Tensor, Tensor]:
    """q [B*Tq,256] (pre-scaled), kv [B*Tk,512] = K | V, kmask fp32 [B,Tk] | None -> (o [B*Tq,256], ml [B*Tq,8,2])."""
    _require_cuda(q, kv)
    lib = _lib.load()
    o = torch.empty_like(q)
    ml = torch.empty(B * Tq, 8, 2, dtype=torch.float32, device=q.device)
    rc = lib.mmr_attention_fwd(dtype, B, Tq, Tk, _ptr(q), _ptr(kv), _ptr(kmask), _ptr(o), _ptr(ml), _stream())
    _lib.check(rc, "mmr_attention_fwd")
    return o, ml


@attention_fwd.register_fake
def _(q, kv, kmask, B, Tq, Tk, dtype):
    return torch.empty_like(q), q.new_empty(B * Tq, 8, 2, dtype=torch.float32)


@torch.library.custom_op("mmr_b200::attention_bwd", mutates_args=())
def attention_bwd(q: Tensor, kv: Tensor, kmask: Optional[Tensor], o: Tensor, ml: Tensor, d_o: Tensor, B: int, Tq: int,
                  Tk: int, dtype: int) -> Tuple[Tensor, Tensor]:
    lib = _lib.load()
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    dvec = torch.empty(B * Tq, 8, dtype=torch.float32, device=q.device)
    rc = lib.mmr_attention_bwd(dtype, B, Tq, Tk, _ptr(q), _ptr(kv), _ptr(kmask), _ptr(o), _ptr(ml), _ptr(d_o), _ptr(dq),
                               _ptr(dkv), _ptr(dvec), _stream())
    _lib.check(rc, "mmr_attention_bwd")
    return dq, dkv


@attention_bwd.register_fake
def _(q, kv, kmask, o, ml, d_o, B, Tq, Tk, dtype):
    return torch.empty_like(q), torch.empty_like(kv)


class AttentionFn(torch.autograd.Function):
    """o [B,Tq,256] = per-head softmax(q k^T + key mask) v for q [B,Tq,256] (already scaled), kv [B,Tk,512] = K | V."""

    @staticmethod
    def forward(ctx, q, kv, kmask):
        dtype = resolve_dtype()
        ct = torch.bfloat16 if dtype == DTYPE_BF16 else torch.float32
        B, Tq, _ = q.shape
        Tk = kv.shape[1]
        qc = q.detach().to(ct).reshape(B * Tq, 256).contiguous()
        kvc = kv.detach().to(ct).reshape(B * Tk, 512).contiguous()
        km = _f32c(kmask.detach()) if kmask is not None else None
        o, ml = attention_fwd(qc, kvc, km, B, Tq, Tk, dtype)
        ctx.save_for_backward(qc, kvc, o, ml, *([km] if km is not None else []))
        ctx.cfg = (B, Tq, Tk, dtype, q.dtype, kv.dtype)
        return o.view(B, Tq, 256)

    @staticmethod
    def backward(ctx, d_o):
        B, Tq, Tk, dtype, qdt, kvdt = ctx.cfg
        sv = ctx.saved_tensors
        qc, kvc, o, ml = sv[:4]
        km = sv[4] if len(sv) > 4 else None
        dq, dkv = attention_bwd(qc, kvc, km, o, ml, d_o.to(o.dtype).reshape(B * Tq, 256).contiguous(), B, Tq, Tk, dtype)
        return dq.view(B, Tq, 256).to(qdt), dkv.view(B, Tk, 512).to(kvdt), None


# --------------------------------------------------------------------------------------------
# standalone RoutePrimaryProjector.forward (routing_and_heads.py:111-121)
@torch.library.custom_op("mmr_b200::projector_fwd", mutates_args=())
def projector_fwd(embs: Tensor, proj_w: Sequence[Tensor], proj_b: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """embs fp32 [10,B,256] -> (poses [B,10,32], acts [B,10,1])."""
    _require_cuda(embs, *proj_w)
    lib = _lib.load()
    B = embs.shape[1]
    rp = RoutingParams()
    for r in range(N_ROUTES):
        rp.proj_w[r] = proj_w[r].data_ptr()
        rp.proj_b[r] = proj_b[r].data_ptr()
    poses = torch.empty(B, N_ROUTES, 32, dtype=torch.float32, device=embs.device)
    acts = torch.empty(B, N_ROUTES, 1, dtype=torch.float32, device=embs.device)
    _lib.check(lib.mmr_projector_fwd(C.byref(rp), _ptr(embs), B * 256, 256, B, _ptr(poses), _ptr(acts), _stream()),
               "mmr_projector_fwd")
    return poses, acts


@projector_fwd.register_fake
def _(embs, proj_w, proj_b):
    B = embs.shape[1]
    return embs.new_empty(B, N_ROUTES, 32), embs.new_empty(B, N_ROUTES, 1)


@torch.library.custom_op("mmr_b200::projector_bwd", mutates_args=())
def projector_bwd(embs: Tensor, proj_w: Sequence[Tensor], proj_b: Sequence[Tensor], d_poses: Optional[Tensor],
                  d_acts: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """Returns (d_embs [10,B,256], flat grads: proj_w[10] (33*256 each) | proj_b[10] (36 each, 33 used))."""
    lib = _lib.load()
    B = embs.shape[1]
    dev = embs.device
    rp = RoutingParams()
    g = RoutingGrads()
    flat = torch.zeros(N_ROUTES * (33 * 256 + 36), dtype=torch.float32, device=dev)
    base = flat.data_ptr()
    for r in range(N_ROUTES):
        rp.proj_w[r] = proj_w[r].data_ptr()
        rp.proj_b[r] = proj_b[r].data_ptr()
        g.proj_w[r] = base + 4 * (r * 33 * 256)
        g.proj_b[r] = base + 4 * (N_ROUTES * 33 * 256 + r * 36)
    scratch = torch.empty(B * 330, dtype=torch.float32, device=dev)
    d_embs = torch.empty(N_ROUTES, B, 256, dtype=torch.float32, device=dev)
    _lib.check(lib.mmr_projector_bwd(C.byref(rp), _ptr(embs), B * 256, 256, B, _ptr(d_poses), _ptr(d_acts), _ptr(scratch),
                                     C.byref(g), _ptr(d_embs), _stream()), "mmr_projector_bwd")
    return d_embs, flat


@projector_bwd.register_fake
def _(embs, proj_w, proj_b, d_poses, d_acts):
    return torch.empty_like(embs), embs.new_empty(N_ROUTES * (33 * 256 + 36))


class ProjectorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *rest):
        embs = rest[:N_ROUTES]
        proj_w = [p.detach() for p in rest[N_ROUTES:2 * N_ROUTES]]
        proj_b = [p.detach() for p in rest[2 * N_ROUTES:3 * N_ROUTES]]
        embs_dense = torch.stack([_f32c(e.detach()) for e in embs], dim=0)
        poses, acts = projector_fwd(embs_dense, proj_w, proj_b)
        ctx.save_for_backward(embs_dense, *proj_w, *proj_b)
        return poses, acts

    @staticmethod
    def backward(ctx, d_poses, d_acts):
        sv = list(ctx.saved_tensors)
        embs_dense, proj_w, proj_b = sv[0], sv[1:1 + N_ROUTES], sv[1 + N_ROUTES:]
        d_embs, flat = projector_bwd(embs_dense, proj_w, proj_b, _f32c(d_poses) if d_poses is not None else None,
                                     _f32c(d_acts) if d_acts is not None else None)
        nb = N_ROUTES * 33 * 256
        g_pw = [flat[r * 33 * 256:(r + 1) * 33 * 256].view(33, 256) for r in range(N_ROUTES)]
        g_pb = [flat[nb + r * 36: nb + r * 36 + 33] for r in range(N_ROUTES)]
        return tuple(d_embs[r] for r in range(N_ROUTES)) + tuple(g_pw) + tuple(g_pb)


def debug_gemm(engine: int, dtype: int, trans: bool, A: Tensor, B: Tensor, bias: Optional[Tensor]) -> Tensor:
    """Unit-test hook: C = A @ B^T (+bias)  or, trans, C = A^T @ B (reduction over rows)."""
    _require_cuda(A, B)
    lib = _lib.load()
    if trans:
        Kr, M = A.shape
        N = B.shape[1]
        K = Kr
    else:
        M, K = A.shape
        N = B.shape[0]
    Cc = torch.empty(M, N, dtype=torch.float32, device=A.device)
    rc = lib.mmr_debug_gemm(engine, dtype, int(trans), M, N, K, _ptr(A), _ptr(B), _ptr(bias), _ptr(Cc), _stream())
    _lib.check(rc, "mmr_debug_gemm")
    return Cc
