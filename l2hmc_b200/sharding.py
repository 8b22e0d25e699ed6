"""Multi-GPU execution: independent chains sharded over ranks, one all-gather of samples at the end.

The reference has no distributed code at all; the only scalable axis of the
sampling path is the chain axis N, and every op on it is row-wise.  So: one
process per GPU (torchrun), contiguous N/G chains per rank, parameters replicated (each rank builds the
same ``Dynamics``), Philox keyed by the GLOBAL chain id (``chain_offset``) so results do not depend on
G, and a single ``all_gather`` of (x', p) when the caller wants the full sample set.  No collective
runs inside the leapfrog loop.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of chains owned by ``rank``; the first n % world ranks hold one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the torchrun job described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*; returns
    (rank, world, local_rank).  Single-process runs (no env) return (0, 1, 0) without a group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def all_gather_chains(local: torch.Tensor, n_total: int, group=None, out=None) -> torch.Tensor:
    """Reassemble a chain-sharded tensor ([n_local, ...] per rank, shard_bounds layout) into
    [n_total, ...] on every rank with ONE all_gather (uneven shards are padded to the largest).
    ``out``: optional preallocated [n_total, ...] result (used when the shards are even)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local.shape[0] != n_total:
            raise ValueError("single-rank gather: local has %d chains, expected %d" % (local.shape[0], n_total))
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(n_total, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d chains, shard_bounds says %d" % (rank, local.shape[0], hi - lo))
    biggest = shard_bounds(n_total, 0, world)
    pad = biggest[1] - biggest[0]
    buf = local
    if local.shape[0] != pad:
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    if out is None or tuple(out.shape) != (world * pad,) + tuple(local.shape[1:]):
        out = torch.empty((world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if n_total == world * pad:
        return out
    parts = []
    for r in range(world):
        l, h = shard_bounds(n_total, r, world)
        parts.append(out[r * pad: r * pad + (h - l)])
    return torch.cat(parts, dim=0)


class ShardedSampler(object):
    """Runs ``propose`` on this rank's shard of a global batch of chains."""

    def __init__(self, dynamics, n_total: int, rank: int = 0, world: int = 1):
        self.dynamics = dynamics
        self.n_total = int(n_total)
        self.rank, self.world = int(rank), int(world)
        self.lo, self.hi = shard_bounds(self.n_total, self.rank, self.world)

    def local_slice(self, x_global: torch.Tensor) -> torch.Tensor:
        return x_global[self.lo:self.hi].contiguous()

    def step(self, x_local: torch.Tensor, n_transitions: int = 1, counter: Optional[int] = None):
        """One (or K fused) MH transitions on the local shard; Philox uses global chain ids."""
        from . import _lib
        return self.dynamics._transition(x_local, dir_mode=_lib.DIR_RANDOM, do_mh=True, n_transitions=n_transitions,
                                         counter=counter, chain_offset=self.lo, want_v=False)

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        return all_gather_chains(local, self.n_total)
