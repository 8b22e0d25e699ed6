"""Time the one collective of the path (all-gather of x' shards) on this box: python -m torch.distributed.run ... tools/nccl_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from l2hmc_b200.sharding import all_gather_chains, init_distributed  # noqa: E402

rank, world, local = init_distributed()
torch.cuda.set_device(local)
n, D = 262144, 50
x = torch.randn(n, D, device="cuda")
out = torch.empty(n * world, D, device="cuda")
for i in range(6):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    all_gather_chains(x, n * world, out=out)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("all_gather %d: %.3f ms (events) %.3f ms (wall), %.1f MB per rank" % (i, e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, n * D * 4 / 1e6), flush=True)
ok = torch.equal(out[rank * n:(rank + 1) * n], x)
if rank == 0:
    print("own shard intact:", ok, "| can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1) if world > 1 else None)
dist.barrier()
dist.destroy_process_group()
