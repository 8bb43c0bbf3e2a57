"""probe: host<->device copy bandwidth with page-locked memory (torch pinned and ctp_host_alloc)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ctopprm_learn_b200 import capi

n = 512 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hp = torch.empty(n, dtype=torch.uint8).pin_memory()
mine = torch.from_numpy(capi.pinned_empty((n,), np.uint8))
pageable = torch.empty(n, dtype=torch.uint8)
for name, h in (("torch pinned", hp), ("ctp_host_alloc", mine), ("pageable", pageable)):
    for direction in ("h2d", "d2h"):
        for _ in range(2):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name:16s} {direction}: {n / dt / 1e9:6.1f} GB/s")
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)
hp2 = torch.empty(n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2):
        hp2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"duplex: {n / dt / 1e9:6.1f} GB/s each way")
