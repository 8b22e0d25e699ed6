#!/usr/bin/env python
"""What the host side of a GPU box can move when all ranks copy at once: every rank runs pinned H2D and D2H copies of 52 MB
(the per-step buffers of l2hmc_transition_host on config 2) on two streams concurrently, nothing else.  The aggregate is the
ceiling of the end-to-end (host-buffer) number of bench.py at that GPU count.
    torchrun --nproc-per-node N tools/host_link_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world)
size = 52 << 20
h_in = torch.empty(size, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(size, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(size, dtype=torch.uint8, device="cuda")
d_out = torch.empty(size, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("h2d", "d2h", "both"):
    for it in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 40
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    gbs = reps * size * (2 if mode == "both" else 1) / dt / 1e9
    t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
    if world > 1:
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        print("%d ranks, %-4s: aggregate %.1f GB/s (slowest rank %.1f GB/s)" % (world, mode, float(t[0]), float(lo[0]) if world > 1 else gbs), flush=True)
if world > 1:
    dist.destroy_process_group()
