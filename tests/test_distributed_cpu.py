"""world_size-2 gloo test (CPU) of the multi-GPU host logic: chain sharding, global-chain-id Philox
keying, and the single all-gather that reassembles the samples."""
import os
import socket

import numpy as np
import pytest
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


# ---- data-parallel training: sharded chains, one all-reduce of the gradients ------------------------------------------
@pytest.fixture(scope="module")
def emu_lib():
    import test_train_emu as E
    return E.build_emu()   # builds tests/emu/_build/libtrain_emu.so if needed and loads it


def _train_worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import ctypes as C
    import test_train_emu as E
    from l2hmc_b200 import training
    from l2hmc_b200.sharding import init_distributed, shard_bounds
    init_distributed(backend="gloo")
    lib = C.CDLL(E.OUT)
    lib.emu_loss_grad.restype = C.c_int
    P, x, d, v = _train_problem(n_total)
    lo, hi = shard_bounds(n_total, rank, world)
    # this rank's shard through the (emulated) kernels, the means running over the GLOBAL chain count
    loss, d_eps, gx, gv, _, _ = E.run_emu(lib, P, x[lo:hi], v[lo:hi], d[lo:hi], 0.1, 1.0 / n_total)
    grads = {"loss": torch.tensor([loss], dtype=torch.float32), "eps": torch.tensor([d_eps], dtype=torch.float32),
             "XNet": {k: torch.as_tensor(gx[E_ABI(k)]) for k in training.NAMES},
             "VNet": {k: torch.as_tensor(gv[E_ABI(k)]) for k in training.NAMES}}
    training.allreduce_grads(grads)
    if rank == 0:
        out.put({"loss": float(grads["loss"][0]), "eps": float(grads["eps"][0]),
                 "XNet": {k: grads["XNet"][k].numpy() for k in training.NAMES},
                 "VNet": {k: grads["VNet"][k].numpy() for k in training.NAMES}})
    dist.barrier()
    dist.destroy_process_group()


def E_ABI(k):
    return {"ls": "scale_s", "lq": "scale_q"}.get(k, k)


def _train_problem(n_total):
    P = util.Problem(regime="stress", kind="gaussian", D=3, H=6, T=2, eps=0.1)
    rng = np.random.default_rng(17)
    x = P.x0(n_total, rng)
    d = rng.integers(0, 2, n_total).astype(np.uint8)
    v = rng.standard_normal((n_total, P.D)).astype(np.float32)
    return P, x, d, v


def test_sharded_training_gradients_allreduce_to_the_full_batch(request):
    """Two ranks, uneven shards: per-shard l2hmc_loss_grad (the kernel source on the host, tests/emu) with inv_count =
    1 / N_global, then training.allreduce_grads over gloo == one full-batch call."""
    import ctypes as C
    import test_train_emu as E
    lib = request.getfixturevalue("emu_lib")
    n_total, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    P, x, d, v = _train_problem(n_total)
    loss, d_eps, gx, gv, _, _ = E.run_emu(lib, P, x, v, d, 0.1, 1.0 / n_total)
    assert got["loss"] == pytest.approx(loss, rel=1e-5)
    assert got["eps"] == pytest.approx(d_eps, rel=1e-4, abs=1e-4)
    from l2hmc_b200 import training
    for key, ref in (("XNet", gx), ("VNet", gv)):
        for k in training.NAMES:
            r_ = ref[E_ABI(k)]
            assert np.allclose(got[key][k], r_, rtol=1e-4, atol=1e-5 * max(1e-6, np.abs(r_).max())), (key, k)
