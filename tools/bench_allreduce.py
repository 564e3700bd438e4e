"""Times the gradient all-reduce of the hot path's flat buffers (run under torchrun, one rank per GPU)."""
import os
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
bufs = [torch.randn(19746304, device="cuda"), torch.randn(600000, device="cuda"), torch.randn(90000, device="cuda")]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def sum_then_scale():
    for b in bufs:
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        b.mul_(1.0 / world)


def avg():
    for b in bufs:
        dist.all_reduce(b, op=dist.ReduceOp.AVG)


def avg_big_only():
    dist.all_reduce(bufs[0], op=dist.ReduceOp.AVG)


for name, fn in (("sum + mul_", sum_then_scale), ("AVG", avg), ("AVG, 79 MB buffer only", avg_big_only)):
    ms = timed(fn)
    if rank == 0:
        print(f"{name}: {ms*1e3:.0f} us per step-worth of gradients ({world} ranks)")
dist.destroy_process_group()
