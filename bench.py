"""bench.py -- patients/sec of route fusion + capsule routing, forward + backward (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU port of the reference path

Workload (N=1): BASELINE.json configs[1] -- MIMIC-IV PhenoModel, 25 labels, random-init, synthetic
batch of 512 patients (L 48 x 256, N 16 x 256, I 49 x 256), bf16 fwd+bwd.  For N>1 every rank runs
the same per-GPU batch (weak scaling) and the step ends with the gradient all-reduce over NCCL.
One "step" = MULTModel forward + capsule routing + BCE loss + full backward (parameter and input
gradients) + (N>1) gradient all-reduce.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

LAYERS = 4
METRIC = "patients/sec route-fusion+capsule routing fwd+bwd at 1/2/4/8 B200; % roofline"
UNIT = "patients/sec"
# --config: the default is the configuration BASELINE.json's metric is quoted on (configs[1]); the other two are the
# remaining multi-GPU configs of BASELINE.json, kept as secondary lines (profiles/), not the headline.
WORKLOADS = {
    "pheno512": dict(variant="pheno", K=25, TL=48, TN=16, TI=49, batch=512, scaling="weak",
                     name="BASELINE configs[1]: MIMIC-IV PhenoModel 25-label, random-init, synthetic batch=512/GPU, "
                          "L48/N16/I49 x 256, fwd+bwd bf16"),
    "mort8192": dict(variant="mort", K=2, TL=48, TN=16, TI=49, batch=8192, scaling="strong",
                     name="BASELINE configs[2]: MIMIC-IV MortModel (K=2), random-init, synthetic GLOBAL batch=8192 split over "
                          "the GPUs, L48/N16/I49 x 256, fwd+bwd bf16 + gradient all-reduce"),
    "inspect": dict(variant="pheno", K=3, TL=512, TN=128, TI=196, batch=256, scaling="weak",
                    name="BASELINE configs[4]: INSPECT token counts L512/N128/I196 x 256, 3-label head, synthetic "
                         "batch=256/GPU, fwd+bwd bf16"),
}
WL = dict(WORKLOADS["pheno512"])          # the selected workload (set in main)


def peaks():
    """(burst bf16 TF/s, sustained bf16 TF/s, HBM GB/s, source).  A kernel class timed alone in a short region is held
    against the BURST peak, the whole step inside a long run against the SUSTAINED one (MEASURED_PEAKS.json "how")."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1663.5), d.get("bf16_tflops_sustained", 1402.1), d.get("hbm_gbs", 6548.5), "measured"
    return 1663.5, 1402.1, 6548.5, "fallback (B200_PROFILING.md / the pool's last measured peaks)"


def gemm_flops_per_step(B, q_rows=None):
    """FLOPs of the tensor-core GEMMs by class.  q_rows = {"l": n, "n": n, "i": n}: query rows actually processed per
    modality (packed layout: the valid tokens); default: every token (the algorithmic count of SURVEY.md section 8d)."""
    d, f, L = 256, 1024, LAYERS
    T = {"l": WL["TL"], "n": WL["TN"], "i": WL["TI"]}
    dirs = [("l", "n"), ("l", "i"), ("n", "l"), ("n", "i"), ("i", "l"), ("i", "n")]
    mq = sum((q_rows[q] if q_rows else B * T[q]) for q, _ in dirs)
    mk = sum(B * T[k] for _, k in dirs)
    fwd_tn = L * mq * (2 * d * d + 2 * d * d + 2 * d * f + 2 * f * d) + mk * 2 * d * (L * 2 * d)
    bwd_tn = fwd_tn            # data gradients mirror the forward GEMMs
    wgrad = fwd_tn             # weight gradients: same M*N*K products
    return fwd_tn + bwd_tn, wgrad


def gemm_bytes_per_step(B):
    """Algorithmic HBM bytes of the tensor-core GEMM launches of one step (operands once + outputs once, valid rows;
    weights excluded: 13 MB): per layer q / out projection 2 x (512 + 512) B per query row, fc1 512 + 2048 + 128 (ReLU bits),
    fc2 2048 + 512, the same again for the four data gradients; K/V of all layers 512 + 4096 B per key row forward and
    4096 + 1024 (fp32) backward.  profiles/r1_step_bytes.md."""
    T = {"l": WL["TL"], "n": WL["TN"], "i": WL["TI"]}
    dirs = [("l", "n"), ("l", "i"), ("n", "l"), ("n", "i"), ("i", "l"), ("i", "n")]
    mq = sum(B * T[q] for q, _ in dirs)
    mk = sum(B * T[k] for _, k in dirs)
    per_layer = 1024 + 1024 + 2688 + 2560
    return 2 * LAYERS * mq * per_layer + mk * (4608 + 5120)


def total_flops_per_patient():
    d, L = 256, LAYERS
    T = {"l": WL["TL"], "n": WL["TN"], "i": WL["TI"]}
    dirs = [("l", "n"), ("l", "i"), ("n", "l"), ("n", "i"), ("i", "l"), ("i", "n")]
    fwd = sum(L * (20 * T[q] * d * d + 4 * T[k] * d * d + 4 * T[q] * T[k] * d) for q, k in dirs)
    return 3 * fwd   # fwd+bwd = 3x forward


class ClockSampler:
    """Samples SM clocks and throttle reasons DURING the timed region (NVML every ~5 ms; nvidia-smi fallback)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mx, self.reasons, self.stop = [], 0.0, set(), False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                          ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                              "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        f = [x.strip() for x in out.split(",")]
        self.sm.append(float(f[1])); self.mx = max(self.mx, float(f[2]))
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def _run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.005 if self.nvml is not None else 0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_batch(B, seed):
    from multimodalrouting_b200 import synth
    return synth.make_inputs(B=B, K=WL["K"], seed=seed, TL=WL["TL"], TN=WL["TN"], TI=WL["TI"])


def build_models(device, seed=42):
    from multimodalrouting_b200 import synth
    from multimodalrouting_b200 import MULTModel
    if WL["variant"] == "mort":
        from multimodalrouting_b200.MortModel import routing_and_heads as rh
    else:
        from multimodalrouting_b200.PhenoModel import routing_and_heads as rh
    sdm, sdp, sdh = synth.make_state(K=WL["K"], seed=seed, sharp=1.0)
    mult = MULTModel(256, 256, 256, 256, 256, 256, True, True, True, 8, LAYERS, 0, 0., 0., 0., 0., 0., 0., 0., False)
    proj = rh.RoutePrimaryProjector(256, 32)
    head = rh.CapsuleMortalityHead(32, 64, 3, 0.0, "EM", num_classes=WL["K"])
    mult.load_state_dict(sdm); proj.load_state_dict(sdp); head.load_state_dict(sdh)
    return rh, mult.to(device), proj.to(device), head.to(device), (sdm, sdp, sdh)


def oracle_stepper(batch, sds, device="cpu", autocast=False, seed=43):
    """One fwd + loss + bwd step of the oracle restatement of the reference path (eager PyTorch) on `device`."""
    from oracle import route_fusion_oracle as orc
    from multimodalrouting_b200 import synth
    sdm, sdp, sdh = [{k: v.detach().to(device).clone().requires_grad_(True) for k, v in sd.items()} for sd in sds]
    inp = {k: v.to(device) for k, v in make_batch(batch, seed).items()}

    def step():
        for sd in (sdm, sdp, sdh):
            for v in sd.values():
                v.grad = None
        xs = [inp[k].clone().requires_grad_(True) for k in ("x_l", "x_n", "x_i")]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            logits, _, _, _ = orc.full_forward(sdm, sdp, sdh, xs[0], xs[1], xs[2], inp["mL"], inp["mN"], inp["mI"],
                                               variant=WL["variant"], route_mask=inp["route_mask"])
        loss = synth.loss_fn(logits, inp["y"], WL["variant"])
        loss.backward()
        return loss
    return step


def cpu_oracle_rate(budget_s, batch, threads, sds, min_iters=2):
    """Times the CPU port of the reference path (oracle, fp32) fwd+bwd on `threads` host threads."""
    torch.set_num_threads(threads)
    step = oracle_stepper(batch, sds)
    step()
    times = []
    t_end = time.time() + budget_s
    while len(times) < min_iters or (time.time() < t_end and len(times) < 50):
        t0 = time.time(); step(); times.append(time.time() - t0)
    return batch / statistics.median(times), len(times)


def gpu_eager_reference(batch, sds, device, iters=10, warmup=3):
    """The reference's eager path ON THIS GPU ("the number to beat"): the oracle restatement of the reference
    modules (same ATen op families: F.linear / bmm / einsum / layer_norm / fp32 softmax) on CUDA under
    torch.autocast(bfloat16) with TF32 enabled, exactly the switches the reference drivers set
    (MIMIC-IV/MortModel/Paired_Cross_Attention/main.py:2613-2621 TF32, :2633-2659 autocast).  Same batch, CUDA events."""
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        step = oracle_stepper(batch, sds, device=device, autocast=True)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return {"value": batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": batch, "iters": iters,
            "what": "reference eager path on the same B200: oracle restatement of the reference modules (eager PyTorch/ATen, "
                    "cuBLAS GEMMs) under torch.autocast(bfloat16), TF32 on (main.py:2613-2621,2633-2659), device-resident "
                    "inputs, CUDA events; host launch overhead included, as a user of the reference pays it"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own (CPU, eager PyTorch fp32) implementation of the path, as restated in oracle/
    (the reference is a Python script tree that cannot be installed or shipped to the GPU box), on
    all host threads, on the SAME per-GPU batch as the product arm."""
    if rank != 0:
        return
    from multimodalrouting_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cap = 512 if WL["TL"] <= 64 else 32        # bounded sample: the whole --steps/--warmup run must end within minutes
    batch = min(args.batch, cap)
    sds = synth.make_state(K=WL["K"], seed=42)
    step = oracle_stepper(batch, sds)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        float(step().detach())
    dt = time.time() - t0
    value = batch * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": WL["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WL["name"], "batch_per_gpu": batch, "global_batch": batch * args.gpus,
                       "parallelism": f"dp{args.gpus}", "reference_sample_per_step": batch,
                       "torch_num_threads": torch.get_num_threads()},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "torch_num_threads": torch.get_num_threads(),
                             "sample": f"{args.steps} steps x {batch} patients ("
                                       + ("the product arm's full per-GPU batch" if batch == args.batch else
                                          f"bounded sample of the {args.batch}-patient per-GPU batch")
                                       + "), oracle port of the reference eager path, fp32, rank 0 only"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _teardown(dist, reducer):
    """Leaves the process group.  With collectives captured inside a CUDA graph the communicator cannot be
    destroyed while the graph is alive (destroy_process_group then waits forever), so that configuration skips the
    explicit destroy: the work is complete and synchronised, the process simply exits."""
    torch.cuda.synchronize()
    if reducer is not None:
        reducer.close()
        sys.stderr.flush()
        os._exit(0)       # only with --overlap (opt-in); the default path below leaves the group cleanly
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="pheno512", choices=sorted(WORKLOADS),
                    help="pheno512 = BASELINE configs[1] (the headline); mort8192 = configs[2] (global batch 8192, strong "
                         "scaling); inspect = configs[4] token counts (256 patients per GPU)")
    ap.add_argument("--batch", type=int, default=0, help="patients per GPU (default: the workload's)")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference-eager-on-this-GPU leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused-loss", action="store_true",
                    help="opt-in: compute the loss with the device-side loss tail (losses.pheno_train_loss) instead of torch BCE")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--overlap", action="store_true",
                    help="N>1: all-reduce the gradients layer block by layer block on a side stream during the backward "
                         "(captured into the step graph) instead of after it; measured equal at N=2 (5.91 vs 5.89 ms), "
                         "so the simpler after-the-backward reduction is the default")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    WL.clear(); WL.update(WORKLOADS[args.config])
    if args.batch <= 0:
        args.batch = WL["batch"] // world if WL["scaling"] == "strong" else WL["batch"]
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)                 # library banners (e.g. NCCL's version line) must not pollute the one JSON line
    import torch.distributed as dist
    from multimodalrouting_b200 import _lib
    from multimodalrouting_b200.dist import OverlappedGradReducer, allreduce_gradients
    from multimodalrouting_b200 import synth
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev, timeout=__import__("datetime").timedelta(seconds=180))
    lib = _lib.load()
    rh, mult, proj, head, sds = build_models(dev)
    modules = (mult, proj, head)
    # the benchmark step has no optimizer update, so the compute-type weight copies are packed once (outside the captured
    # step) instead of once per forward; MMR_BENCH_REPACK=1 keeps the packing inside every step (what a training loop
    # whose optimizer lives in the same graph pays: +0.09 ms, profiles/r2_launches.txt)
    mult.static_weights = os.environ.get("MMR_BENCH_REPACK") != "1"
    B = args.batch
    inp = make_batch(B, 43 + rank)
    keys = ("x_l", "x_n", "x_i", "mL", "mN", "mI", "route_mask", "y")
    # The batch travels as ONE contiguous arena (256-byte aligned slots): one pinned host buffer, one device staging
    # buffer, one set of static device inputs -- so a step's H2D copy and its move into the static inputs are one
    # copy each instead of eight (the eight small device-to-device copies cost 0.15 ms per step, measured).
    def arena(device=None, pin=False):
        offs, o = {}, 0
        for k in keys:
            offs[k] = o
            o += (inp[k].numel() * inp[k].element_size() + 255) // 256 * 256
        buf = torch.empty(o, dtype=torch.uint8, device=device)
        if pin:
            buf = buf.pin_memory()
        views = {k: buf[offs[k]:offs[k] + inp[k].numel() * inp[k].element_size()].view(inp[k].dtype).view(inp[k].shape)
                 for k in keys}
        return buf, views

    host_buf, host = arena(pin=True)
    for k in keys:
        host[k].copy_(inp[k])
    dev_buf, devb = arena(device=dev)
    dev_buf.copy_(host_buf)
    adapter = rh.RouteDimAdapter(256, 256, 256, 256)
    if args.fused_loss and WL["variant"] != "pheno":
        raise SystemExit("--fused-loss is wired for the Pheno workloads")
    if args.fused_loss:
        # opt-in: the reference's Pheno training loss through the device-side loss tail (csrc/loss.cuh; same BCE, plus
        # coerce_rc_to_report on R) instead of torch's BCE kernels -- not part of the default measurement
        from multimodalrouting_b200 import losses as _losses
        _loss_state = _losses.LossState(dev)

    def fwd_bwd():
        for m in modules:
            m.zero_grad(set_to_none=True)
        xs = [devb[k].detach().requires_grad_(True) for k in ("x_l", "x_n", "x_i")]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, alpha, routes, R = rh.forward_capsule_from_multmodel(
                mult, xs[0], xs[1], xs[2], proj, head, mL=devb["mL"], mN=devb["mN"], mI=devb["mI"],
                route_adapter=adapter, route_mask=devb["route_mask"])
        if args.fused_loss:
            loss = _losses.pheno_train_loss(logits, devb["y"], R, alpha, devb["route_mask"], state=_loss_state).loss
        else:
            loss = synth.loss_fn(logits, devb["y"], WL["variant"])
        loss.backward()
        return loss

    # N>1: the gradient all-reduce is part of the step: a handful of NCCL AVG all-reduces over the flat gradient
    # buffers right after the (graph-replayed) backward.  --overlap instead reduces each layer's block on a side
    # stream as soon as the library signals that it is final (collectives captured into the same CUDA graph).
    reducer = None
    if world > 1 and args.overlap:
        reducer = OverlappedGradReducer(mult, (proj, head), LAYERS)

    def train_step():
        loss = fwd_bwd()
        if reducer is not None:
            reducer.finish()
        return loss

    # The step is ~130 launches of static shape: capture it once (CUDA graph) so the host cost of issuing it
    # (Python + autograd + launches, about as long as the GPU work at B=512) disappears from the step.
    graphed = None
    launches_per_replay = 0
    if not args.no_graph:
        from multimodalrouting_b200.graphs import GraphedStep
        try:
            lc0 = lib.mmr_launch_count()
            graphed = GraphedStep(train_step, warmup=2)
            launches_per_replay = (lib.mmr_launch_count() - lc0) // 3     # 2 warm-up calls + 1 capture
        except Exception as exc:            # noqa: BLE001  -- fall back, and say so
            print(f"[bench] CUDA graph capture failed ({exc!r})", file=sys.stderr)
            graphed = None
            if reducer is not None:         # retry with the collectives outside the graph
                reducer.close()
                reducer = None
                torch.cuda.synchronize()
                try:
                    lc0 = lib.mmr_launch_count()
                    graphed = GraphedStep(train_step, warmup=2)
                    launches_per_replay = (lib.mmr_launch_count() - lc0) // 3
                    print("[bench] captured without the overlapped all-reduce", file=sys.stderr)
                except Exception as exc2:   # noqa: BLE001
                    print(f"[bench] CUDA graph capture failed again ({exc2!r}); running eagerly", file=sys.stderr)
                    graphed = None

    # End-to-end input pipeline: the NEXT step's host batch is copied (pinned host -> device staging buffer) on a
    # side stream while the current step computes; at the start of a step the staged batch is moved into the
    # static input tensors (device-to-device).  Every step still pays for its own H2D copy inside the timed
    # region -- it is just overlapped, as a prefetching data loader would do.
    stage_buf, stage = arena(device=dev)
    stage_i32, dev_i32 = stage_buf.view(torch.int32), dev_buf.view(torch.int32)
    copy_stream = torch.cuda.Stream()
    ev_staged, ev_consumed = torch.cuda.Event(), torch.cuda.Event()
    pipe = {"primed": False, "pending": None, "slot": 0, "last_loss": None}
    # the step's result (the loss) is read back EVERY step: D2H into pinned memory right behind the step, collected
    # one step later so that the host can already issue the next step while the GPU finishes this one
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def collect_loss():
        if pipe["pending"] is not None:
            i = pipe["pending"]
            loss_ev[i].synchronize()
            pipe["last_loss"] = float(loss_host[i])
            pipe["pending"] = None
        return pipe["last_loss"]

    def prefetch_host_batch():
        copy_stream.wait_event(ev_consumed)
        with torch.cuda.stream(copy_stream):
            stage_buf.copy_(host_buf, non_blocking=True)
            ev_staged.record(copy_stream)
        pipe["primed"] = True

    parts = set(os.environ.get("MMR_E2E_PARTS", "h2d,d2d,loss").split(","))   # diagnostic: drop parts of the e2e path

    def step(from_host: bool, last: bool = False):
        if from_host and "h2d" in parts:
            cur = torch.cuda.current_stream()
            if not pipe["primed"]:
                ev_consumed.record(cur)
                prefetch_host_batch()
            cur.wait_event(ev_staged)
            if "d2d" in parts:
                # SM copy kernel (one pass at HBM speed, ~20 us for 60 MB); cudaMemcpyAsync D2D runs on a copy engine
                # at ~1.2 TB/s and cost 0.1 ms per step here
                torch.bitwise_or(stage_i32, 0, out=dev_i32)
            ev_consumed.record(cur)
            pipe["primed"] = False
            if not last:
                prefetch_host_batch()
        loss = graphed() if graphed is not None else train_step()
        if world > 1 and reducer is None:
            allreduce_gradients(modules, world)
        if from_host and "loss" not in parts:
            return loss
        if from_host:
            i = pipe["slot"]
            loss_host[i].copy_(loss.detach(), non_blocking=True)      # D2H read of the step's result
            loss_ev[i].record()
            collect_loss()                                            # result of the PREVIOUS step
            pipe["pending"], pipe["slot"] = i, i ^ 1
            if last:
                return collect_loss()                                 # drain: every step's loss reached the host
            return pipe["last_loss"]
        return loss

    def timed(n, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(from_host, last=(i == n - 1))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    l0 = lib.mmr_launch_count()
    with ClockSampler(local_rank) as cs:
        ms = timed(args.steps, False)
    launches = lib.mmr_launch_count() - l0
    if graphed is not None:     # replayed kernels are not re-issued through the library: count the captured nodes
        launches = launches_per_replay * args.steps
    clocks = cs.summary()
    value = world * B * args.steps / (ms / 1e3)
    # end-to-end through the public API with host buffers (H2D of the inputs + D2H of the loss per step)
    for _ in range(2):
        step(True, last=True)
    ms_e2e = timed(args.steps, True)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = host_buf.numel()

    # per-class device time of OUR kernels (CUDA events on the launching stream), separate pass
    prof = None
    nprof = 3
    if rank == 0:
        lib.mmr_prof_enable(1)
    g_saved, graphed = graphed, None   # the per-class CUDA-event pass issues the kernels eagerly ...
    ws_saved = os.environ.get("MMR_WGRAD_STREAM")
    os.environ["MMR_WGRAD_STREAM"] = "0"      # ... and on ONE stream, so that every kernel class is timed alone
    for _ in range(nprof):          # every rank steps (the gradient all-reduce is a collective)
        # Issuing a step eagerly takes the host longer than the GPU needs to run it, so the events around a launch would
        # also time the GPU waiting for the host.  A ~25 ms device-side spin in front of every step lets the host run ahead:
        # each (event, kernels, event) group is already queued when the GPU reaches it and the interval is device time only.
        torch.cuda._sleep(int(0.025 * 1.9e9))
        step(False)
    graphed = g_saved
    if ws_saved is None:
        os.environ.pop("MMR_WGRAD_STREAM", None)
    else:
        os.environ["MMR_WGRAD_STREAM"] = ws_saved
    torch.cuda.synchronize()
    if rank == 0:
        msc = (C.c_double * 8)(); nc = (C.c_longlong * 8)()
        lib.mmr_prof_collect(msc, nc)
        lib.mmr_prof_enable(0)
        names = ["gemm_tc", "wgrad_tc", "attn_fwd", "attn_bwd", "gemm_simt", "routing", "fusion_fwd_call", "fusion_bwd_call"]
        prof = {n: {"ms_per_step": msc[i] / nprof, "launch_groups_per_step": nc[i] / nprof} for i, n in enumerate(names)}
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            _teardown(dist, reducer)
        return
    tf_burst, tf_sust, hbm_peak, src = peaks()
    packed = os.environ.get("MMR_VARLEN", "1") != "0" and os.environ.get("MMR_ATTN") != "tc"
    q_rows = {k: (int((inp[m] != 0).sum()) if packed else B * WL[t]) for k, m, t in (("l", "mL", "TL"), ("n", "mN", "TN"), ("i", "mI", "TI"))}
    q_dense = {"l": B * WL["TL"], "n": B * WL["TN"], "i": B * WL["TI"]}
    fl_tn, fl_wg = gemm_flops_per_step(B, q_rows)           # FLOPs the kernels execute (valid query rows)
    t_tn = prof["gemm_tc"]["ms_per_step"] / 1e3
    n_tn = max(prof["gemm_tc"]["launch_groups_per_step"], 1)
    achieved = fl_tn / t_tn / 1e12 if t_tn > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 TN GEMM, all six directions per launch)",
                "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
                # dram__bytes_read.sum + dram__bytes_write.sum averaged over the 34 gemm_tc launches of one step,
                # ncu --set full (profiles/r2_gemm_step.md, packed query rows; algorithmic bytes ~195 MB per launch).
                # Other workloads / the dense layout were not captured: null there.
                "traffic": 161.0e6 if (args.config == "pheno512" and packed and B == 512) else None,
                "traffic_unit": "bytes per launch (ncu dram read+write, mean of the 34 launches of a step)",
                "peak_source": f"{src} BURST bf16 (MEASURED_PEAKS.json bf16_tflops): the class is timed alone, eagerly, in a "
                               f"~0.1 s region; frac_of_sustained uses bf16_tflops_sustained",
                "frac_of_sustained": achieved / tf_sust,
                "avg_launch_ms": 1e3 * t_tn / n_tn, "launches_per_step": n_tn,
                "flops_per_launch": fl_tn / n_tn,
                "flops_counted": "executed by the kernels: packed query rows (valid tokens), dense key/value rows",
                "wgrad_tc_tflops": (fl_wg / (prof["wgrad_tc"]["ms_per_step"] / 1e3) / 1e12) if prof["wgrad_tc"]["ms_per_step"] > 0 else None,
                "whole_step_frac_of_tensor_roofline": (value / world) * total_flops_per_patient() / 1e12 / tf_sust,
                "whole_step_peak": tf_sust,
                # the same launches seen from the HBM side: they are HBM-shaped (128-230 FLOP/B against a machine balance
                # of ~214), so this fraction explains the tensor fraction above (profiles/r1_step_bytes.md)
                "hbm_view": {"algorithmic_bytes_per_step": gemm_bytes_per_step(B),
                             "achieved_GBps": (gemm_bytes_per_step(B) / t_tn / 1e9) if t_tn > 0 else 0.0,
                             "frac_of_hbm_peak": (gemm_bytes_per_step(B) / t_tn / 1e9 / hbm_peak) if t_tn > 0 else 0.0}}
    # second roofline the north-star names: capsule routing against HBM bandwidth.  Algorithmic bytes per patient
    # (SURVEY.md section 8d, K=25, fp32 route embeddings): forward 11,420 B + backward 20,600 B.
    rt_ms = prof["routing"]["ms_per_step"]
    Kl = WL["K"]
    rt_bytes = ((10240 + 40 + 4 * Kl + 40 + 40 * Kl) + (10240 + 4 * Kl + 40 * Kl + 10240)) * B
    rt_gbs = rt_bytes / (rt_ms / 1e3) / 1e9 if rt_ms > 0 else 0.0
    split = os.environ.get("MMR_RT_SPLIT", "1") != "0"
    roofline_routing = {"bound": "hbm",
                        "kernel": ("rs_project + rs_votes + rs_iterate_fwd | rs_iterate_bwd + rs_dpose (csrc/routing_split.cuh) "
                                   "+ head / vote-weight gradient launches") if split else
                                  "routing_fwd_kernel + routing_bwd_kernel (+ head / vote-weight gradient launches)",
                        "achieved": rt_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": rt_gbs / hbm_peak, "traffic": None,
                        "ms_per_step": rt_ms, "algorithmic_bytes_per_step": rt_bytes,
                        "note": "timed eagerly between CUDA events (launch gaps included); latency / issue bound, not byte bound: "
                                "the algorithmic bytes of a step take 2.5 us (K=25, B=512) / 39 us (K=2, B=8192) at the HBM peak, "
                                "the per-patient agreement iterations ~9k dependent warp instructions (profiles/r2_routing.md)"}
    cpu = None
    gpu_ref = None
    if not args.no_gpu_reference and world == 1:
        try:
            gpu_ref = gpu_eager_reference(B, sds, dev)
            gpu_ref["speedup_device_resident"] = value / gpu_ref["value"]
            gpu_ref["speedup_e2e"] = e2e / gpu_ref["value"]
        except Exception as exc:      # noqa: BLE001 -- a reported baseline must not take the product line down
            gpu_ref = {"unavailable": repr(exc)}
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        cb = min(B, 128)
        rate, iters = cpu_oracle_rate(15.0, cb, threads, sds)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "torch_num_threads": torch.get_num_threads(),
               "sample": f"{iters} fwd+bwd iterations x {cb} patients of the same workload, oracle port (eager PyTorch fp32)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": WL["scaling"],
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WL["name"], "config_key": args.config, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "cuda_graph": graphed is not None,
                       "query_rows": (f"packed: {2 * sum(q_rows.values())} of {2 * sum(q_dense.values())} query rows hold a valid "
                                      f"token (padded tokens contribute nothing to outputs or gradients and get no row); "
                                      f"the whole-step roofline fraction still counts the dense FLOPs of SURVEY 8d"
                                      if packed else "dense (MMR_VARLEN=0)"),
                       "packed_weights": ("packed once per parameter version (no optimizer update inside the benchmark step)"
                                          if mult.static_weights else "re-packed inside every step"),
                       "wgrad_side_stream": os.environ.get("MMR_WGRAD_STREAM", "1") != "0",
                       "grad_allreduce": (None if world == 1 else
                                          "overlapped with the backward, per layer block" if reducer is not None
                                          else "after the backward"),
                       "l2_policy": "per-step working set (activations saved for backward ~3 GB) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_routing": roofline_routing,
            "cpu_baseline": cpu, "reference_gpu_eager": gpu_ref,
            "kernel_time_ms_per_step": prof}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        _teardown(dist, reducer)


if __name__ == "__main__":
    main()
