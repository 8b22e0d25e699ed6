"""world_size-2 gloo test (CPU) of the multi-GPU host logic: chain sharding, global-chain-id Philox
keying, and the single all-gather that reassembles the samples."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401  (sys.path)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, d, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from l2hmc_b200 import philox
    from l2hmc_b200.sharding import all_gather_chains, init_distributed, shard_bounds
    r, w, _ = init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = shard_bounds(n_total, rank, world)
    # each rank produces its shard's "samples" from the global-id-keyed generator
    local = torch.as_tensor(philox.normals(7, 3, hi - lo, d, chain_offset=lo))
    px = torch.as_tensor(philox.direction_and_uniform(7, 3, hi - lo, chain_offset=lo)[1])
    full = all_gather_chains(local, n_total)
    full_px = all_gather_chains(px, n_total)
    if rank == 0:
        out.put((full.numpy(), full_px.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(n_total, world=2, d=5):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, full_px = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from l2hmc_b200 import philox
    assert np.array_equal(full, philox.normals(7, 3, n_total, d))
    assert np.array_equal(full_px, philox.direction_and_uniform(7, 3, n_total)[1])


def test_allgather_even_shards():
    _run(64)


def test_allgather_ragged_shards():
    _run(37)
