"""probe: host<->device copy bandwidth with all ranks copying at once (torchrun): what the box's host side gives N GPUs"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from ctopprm_learn_b200 import capi

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
d1, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
h1, h2 = torch.from_numpy(capi.pinned_empty((n,), np.uint8)), torch.from_numpy(capi.pinned_empty((n,), np.uint8))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, reps=6):
    for it in range(reps + 1):
        if it == 1:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
        if mode in ("h2d", "duplex"):
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if mode in ("d2h", "duplex"):
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return n / float(t.item()) / 1e9
for mode in ("h2d", "d2h", "duplex"):
    gbs = run(mode)
    if rank == 0:
        print(f"{world} ranks at once, {mode}: {gbs:6.1f} GB/s per rank per direction, {gbs * world:7.1f} GB/s aggregate per direction")
