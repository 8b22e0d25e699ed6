"""Training path: the SCG notebook's objective, its gradient and its optimiser loop on the GPU.

SURVEY section 8(f)3; DESIGN.md section 7.1: the gradient the reference gets from TF1 autodiff --
``optimizer.minimize(loss)`` (SCGExperiment.ipynb:183-188) through ``propose`` (utils/sampler.py:28-51), ``p_accept``
(utils/dynamics.py:302-309), the unrolled leapfrog (:246-300) and ``tf.gradients(energy, x)`` inside it (:217-218) -- is
computed by ``l2hmc_loss_grad``: for the notebook's small nets (x_dim <= 4, width <= 16) ONE launch of a fused kernel
(csrc/train_small.cuh: one chain per thread, recorded forward sweep and hand-written reverse sweep on chip); for larger
nets a launch sequence (csrc/train.cuh: plain fp32 FMA GEMMs -- correct, not fast).  Gaussian, GMM, RoughWell and funnel
targets, no aux.  No CPU fallback.

    loss, grads, Lx, px = loss_and_grads(dynamics, x)                  # one propose batch
    state = train_step(dynamics, opt, samples)                         # one iteration of SCGExperiment.ipynb:254-270
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .dynamics import TORCH_FLOAT
from .layers import load_stq_net
from .sampler import tf_accept

LOSSES = {"mixed": 0, "standard": 1, "inverse": 2, "logsumexp": 3}   # get_loss names, utils/losses.py:26-34
NAMES = ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "Ws", "bs", "Wt", "bt", "Wq", "bq", "ls", "lq")
_ABI = {"ls": "scale_s", "lq": "scale_q"}


def _draw(dynamics, n, device, want_u=False, chain_offset=0):
    """Direction bits, momenta (and accept uniforms) from the in-kernel Philox stream (l2hmc_philox_fill), keyed by the
    GLOBAL chain id ``chain_offset + i`` so that a sharded batch draws what the single-rank batch draws."""
    v = torch.empty((n, dynamics.x_dim), dtype=TORCH_FLOAT, device=device)
    d = torch.empty((n,), dtype=torch.uint8, device=device)
    u = torch.empty((n,), dtype=TORCH_FLOAT, device=device) if want_u else None
    dynamics._chk(dynamics._lib.l2hmc_philox_fill(dynamics._ctx, n, int(chain_offset), dynamics.seed, dynamics.next_counter(), v.data_ptr(),
                                                  d.data_ptr(), u.data_ptr() if want_u else None, dynamics._stream()))
    return d, v, u


def zero_grads(dynamics, device) -> Dict[str, object]:
    """Gradient accumulators shaped like the parameters: {'XNet': {...}, 'VNet': {...}, 'eps': [1], 'loss': [1]}."""
    if dynamics.hmc:
        raise ValueError("an HMC-mode Dynamics has no parameters to train")

    def like(p):
        return {k: torch.zeros(tuple(np.shape(p[k])), dtype=TORCH_FLOAT, device=device) for k in NAMES}
    return {"XNet": like(dynamics._net_params[0]), "VNet": like(dynamics._net_params[1]),
            "eps": torch.zeros(1, dtype=TORCH_FLOAT, device=device), "loss": torch.zeros(1, dtype=TORCH_FLOAT, device=device)}


def accumulate_loss_grads(dynamics, x, acc, *, rng: Optional[dict] = None, scale: float = 0.1, count: Optional[int] = None,
                          loss: str = "mixed", chain_offset: int = 0):
    """Add one propose batch to ``acc`` (from zero_grads): loss += scale mean(1/v) - mean(v)/scale with
    v = |x - Lx|^2 px + 1e-4 (SCGExperiment.ipynb:171-181; utils/losses.py:36-59), gradients likewise.
    rng: optional {'direction' uint8 [N], 'v' [N, D]}; drawn from the Philox stream otherwise.  ``count``: the number of
    chains the means run over (default N).  ``loss``: 'mixed' (the notebook's; ``scale`` applies), 'standard', 'inverse'
    or 'logsumexp' -- ``get_loss(name)`` of utils/losses.py:26-59; the last two are means of one batch and do not add
    over calls.  ``chain_offset``: global index of this shard's first chain (data-parallel training: internal draws are
    keyed by global chain id, as sharding.ShardedSampler does).  Returns (Lx, px)."""
    if loss not in LOSSES:
        raise ValueError("loss must be one of %s" % (sorted(LOSSES),))
    if loss in ("inverse", "logsumexp") and count is not None and int(count) != int(x.shape[0]):
        raise ValueError("loss %r weighs every chain by a statistic of the whole batch: it cannot be accumulated over "
                         "shards or calls (count must be the batch size)" % loss)
    if dynamics.hmc:
        raise ValueError("an HMC-mode Dynamics has no parameters to train")
    dynamics._ensure_ctx()
    dynamics._sync_temperature()
    x = dynamics._prep(x, "x", dynamics.x_dim)
    n, dev = x.shape[0], x.device
    rng = rng or {}
    if "direction" in rng and "v" in rng:
        d = rng["direction"].detach().to(device=dev, dtype=torch.uint8).contiguous()
        v = dynamics._prep(rng["v"], "v", dynamics.x_dim)
    else:
        d, v, _ = _draw(dynamics, n, dev, chain_offset=chain_offset)
        d = rng["direction"].detach().to(device=dev, dtype=torch.uint8).contiguous() if "direction" in rng else d
        v = dynamics._prep(rng["v"], "v", dynamics.x_dim) if "v" in rng else v
    if d.numel() != n or v.shape[0] != n:
        raise ValueError("direction and v must have N rows")
    Lx = torch.empty((n, dynamics.x_dim), dtype=TORCH_FLOAT, device=dev)
    px = torch.empty((n,), dtype=TORCH_FLOAT, device=dev)
    a = _lib.LossGradArgs()
    a.n, a.x, a.v, a.dir = n, x.data_ptr(), v.data_ptr(), d.data_ptr()
    a.scale, a.inv_count = float(scale), 1.0 / float(count if count is not None else n)
    a.loss, a.d_eps = acc["loss"].data_ptr(), acc["eps"].data_ptr()
    for field, key in (("grad_xnet", "XNet"), ("grad_vnet", "VNet")):
        g = getattr(a, field)
        for k in NAMES:
            t = acc[key][k]
            if not (t.device == dev and t.dtype == TORCH_FLOAT and t.is_contiguous()):
                raise TypeError("gradient accumulators must be contiguous fp32 tensors on the device of x")
            setattr(g, _ABI.get(k, k), t.data_ptr())
    a.x_out, a.px_out, a.stream = Lx.data_ptr(), px.data_ptr(), dynamics._stream()
    a.loss_kind = LOSSES[loss]
    dynamics._chk(dynamics._lib.l2hmc_loss_grad(dynamics._ctx, C.byref(a)))
    return Lx, px


def loss_and_grads(dynamics, x, *, rng=None, scale=0.1, loss="mixed", count=None, chain_offset=0):
    """Value and gradient of one propose batch.  Returns (loss [1], grads, Lx, px); grads['alpha'] is the gradient for
    the reference's trainable ``alpha = log(eps)`` (utils/dynamics.py:50-58)."""
    acc = zero_grads(dynamics, x.device)
    Lx, px = accumulate_loss_grads(dynamics, x, acc, rng=rng, scale=scale, loss=loss, count=count, chain_offset=chain_offset)
    acc["alpha"] = acc["eps"] * dynamics.eps
    return acc["loss"], acc, Lx, px


def notebook_loss_and_grads(dynamics, x, z, *, rng_x=None, rng_z=None, scale=0.1, count=None, chain_offset=0):
    """The notebook objective (SCGExperiment.ipynb:159-181): proposals from the samples x and from noise z.
    Returns (loss [1], grads, Lx, px) with Lx, px those of the x batch (what the training loop feeds back)."""
    acc = zero_grads(dynamics, x.device)
    Lx, px = accumulate_loss_grads(dynamics, x, acc, rng=rng_x, scale=scale, count=count, chain_offset=chain_offset)
    accumulate_loss_grads(dynamics, z, acc, rng=rng_z, scale=scale, count=count, chain_offset=chain_offset)
    acc["alpha"] = acc["eps"] * dynamics.eps
    return acc["loss"], acc, Lx, px


def allreduce_grads(grads, group=None):
    """Data-parallel training (SURVEY section 8e applied to the training path): chains are sharded over the ranks, every
    rank accumulates its shard with ``count`` = the GLOBAL number of chains, and this sums loss and gradients over the
    ranks in ONE all-reduce of the flattened vector (NCCL over NVLink / NVSwitch on GPUs; about 72 k floats for config 2,
    so latency- not bandwidth-bound).  The exchange the reference never had to make: it trains on one device."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    tensors = [grads["loss"], grads["eps"]] + [grads[key][k] for key in ("XNet", "VNet") for k in NAMES]
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for t in tensors:
        t.copy_(flat[o:o + t.numel()].reshape(t.shape))
        o += t.numel()
    return grads   # derive grads["alpha"] = grads["eps"] * eps after the reduction


class Adam(object):
    """tf.train.AdamOptimizer(learning_rate) with the notebook's schedule
    tf.train.exponential_decay(1e-3, global_step, 1000, 0.96, staircase=True) (SCGExperiment.ipynb:183-186).

    The S/T/Q nets are small (541 parameters in the notebook, 71.7 k at width 100): the update itself is host arithmetic
    in float32 (TF's formulation), fed by ONE device-to-host read of the flattened gradients per step; the master
    parameters live in the Dynamics' layer objects, one ``refresh()`` pushes them back to the library."""

    def __init__(self, dynamics, learning_rate=1e-3, decay_steps=1000, decay_rate=0.96, beta1=0.9, beta2=0.999, epsilon=1e-8,
                 eps_trainable=None):
        self.lr0, self.decay_steps, self.decay_rate = float(learning_rate), int(decay_steps), float(decay_rate)
        self.b1, self.b2, self.epsilon = float(beta1), float(beta2), float(epsilon)
        self.global_step = 0
        # Dynamics(eps_trainable=...) decides whether alpha is a variable (utils/dynamics.py:49-56)
        self.eps_trainable = bool(dynamics.eps_trainable if eps_trainable is None else eps_trainable)
        self._m, self._v = {}, {}

    @property
    def learning_rate(self):
        return self.lr0 * self.decay_rate ** (self.global_step // self.decay_steps)

    def _update(self, name, p, g):
        """p, g: float32 numpy arrays of one variable; returns the new value (float32)."""
        f32 = np.float32
        m = self._m.get(name)
        if m is None:
            m = self._m[name] = np.zeros_like(g, dtype=f32)
            self._v[name] = np.zeros_like(g, dtype=f32)
        v = self._v[name]
        m *= f32(self.b1)
        m += f32(1.0 - self.b1) * g
        v *= f32(self.b2)
        v += f32(1.0 - self.b2) * (g * g)
        t = self.global_step + 1
        lr_t = f32(self.learning_rate * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t))   # TF's formulation
        return (p - lr_t * m / (np.sqrt(v) + f32(self.epsilon))).astype(f32)

    def apply(self, dynamics, grads):
        tensors = [grads["eps"].reshape(-1)] + [grads[key][k].reshape(-1) for key in ("XNet", "VNet") for k in NAMES]
        flat = torch.cat(tensors).detach().to(torch.float32).cpu().numpy()   # the one device -> host read of the step
        g_eps = flat[0]
        o = 1
        for i, key in enumerate(("XNet", "VNet")):
            new = {}
            for k in NAMES:
                p = np.asarray(dynamics._net_params[i][k], np.float32)
                g = flat[o:o + p.size].reshape(p.shape)
                o += p.size
                new[k] = self._update(key + "/" + k, p, g)
            load_stq_net(dynamics.XNet if i == 0 else dynamics.VNet, new)
        dynamics.refresh()
        if self.eps_trainable:
            # alpha is the stored variable (utils/dynamics.py:50-54): update it directly, no log(exp(.)) round trip
            alpha = np.asarray([dynamics.alpha], np.float32)
            g_alpha = np.asarray([g_eps * np.float32(dynamics.eps)], np.float32)
            dynamics.set_alpha(float(self._update("alpha", alpha, g_alpha)[0]))
        self.global_step += 1


def train_step(dynamics, opt, samples, *, scale=0.1, rng_x=None, rng_z=None, z=None, u=None, count=None, chain_offset=0,
               group=None):
    """One iteration of the notebook's loop (SCGExperiment.ipynb:254-270): loss on the current samples and on fresh noise,
    Adam update, and the Metropolis-Hastings output ``tf_accept(x, Lx, px)`` as the next samples (:159-160).
    Data-parallel: pass this rank's shard as ``samples``, ``count`` = the global number of chains and ``chain_offset`` =
    the global index of the shard's first chain; loss and gradients are summed over the ranks (allreduce_grads) before
    the (replicated) Adam update.  Returns dict(loss, px, samples, learning_rate)."""
    x = dynamics._prep(samples, "samples", dynamics.x_dim)
    if z is None:
        dynamics._ensure_ctx()
        _, z, _ = _draw(dynamics, x.shape[0], x.device, chain_offset=chain_offset)   # tf.random_normal(tf.shape(x)) (:161)
    lr = opt.learning_rate
    loss, grads, Lx, px = notebook_loss_and_grads(dynamics, x, z, rng_x=rng_x, rng_z=rng_z, scale=scale, count=count,
                                                  chain_offset=chain_offset)
    nxt = tf_accept(x, Lx, px, u=u, seed=dynamics.seed, counter=dynamics.next_counter(), chain_offset=chain_offset)
    if count is not None:
        allreduce_grads(grads, group=group)
    opt.apply(dynamics, grads)
    return {"loss": float(loss[0]), "px": px, "samples": nxt, "learning_rate": lr}
