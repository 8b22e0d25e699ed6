"""Sampling operators (mirror of the reference's utils/sampler.py).

``propose`` (:28-51), ``tf_accept`` (:53-55) and ``chain_operator`` (:57-85) keep the reference's
signatures and return values; keyword-only arguments after them are additions.

What differs underneath: the reference runs BOTH directions for every chain and blends them with the
direction bit (:35-44).  The fused kernel runs only the selected direction of each chain with one
momentum draw -- the same transition kernel in distribution, half the work, and a non-finite value in
the discarded direction can no longer leak into the result through ``mask * a + (1 - mask) * b``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .dynamics import Dynamics, TORCH_FLOAT


def propose(x, dynamics, init_v=None, aux=None, do_mh_step=False, log_jac=False, *, rng=None, n_transitions=1, stats=None):
    """One L2HMC (or HMC) proposal; returns ``(Lx, Lv, px, outputs)`` like the reference.

    rng: optional dict with explicit randomness -- 'direction' uint8 [N] (1 = forward), 'v' [N, D]
    momentum for the selected direction, 'u' [N] accept uniforms.  Missing entries are drawn in-kernel
    (Philox keyed by dynamics.seed and its call counter).
    n_transitions > 1 iterates the MH chain on-chip (requires do_mh_step); outputs are the last ones.
    stats: optional CUDA float64 [2] accumulator: += sum of px, += number of accepted proposals, reduced in the kernel
    (what the notebook prints as np.mean(px_), SCGExperiment.ipynb:268, without a second pass over px).
    """
    rng = rng or {}
    if dynamics.hmc:
        # utils/sampler.py:29-31 -- forward only, init_v forwarded, MH output always appended
        v = init_v if init_v is not None else rng.get("v")
        o = dynamics._transition(x, v=v, dir_mode=_lib.DIR_FORWARD, u=rng.get("u"), do_mh=True,
                                 n_transitions=n_transitions, aux=aux, stats=stats)
        return o["Lx"], o["Lv"], o["px"], [o["x_next"]]
    o = dynamics._transition(x, v=rng.get("v"), dir_mode=_lib.DIR_RANDOM, direction=rng.get("direction"),
                             u=rng.get("u"), log_jac=log_jac, do_mh=do_mh_step, n_transitions=n_transitions, aux=aux,
                             stats=stats)
    Lv = o["Lv"] if init_v is not None else None  # utils/sampler.py:40-42
    outputs = []
    if do_mh_step:
        outputs.append(o["x_next"])
    return o["Lx"], Lv, o["px"], outputs


_util_ctx = {}


def _util(x_dim, device_index):
    """A parameter-free context for the elementwise helper kernels."""
    key = (int(x_dim), int(device_index))
    ctx = _util_ctx.get(key)
    if ctx is None:
        lib = _lib.load()
        cfg = _lib.Config(key[0], 1, 1, 1, key[1], _lib.KERNEL_AUTO, 0.1)
        ctx = C.c_void_p()
        rc = lib.l2hmc_create(C.byref(cfg), C.byref(ctx))
        if rc != 0:
            raise _lib.L2HMCError(rc, (lib.l2hmc_last_error(None) or b"?").decode())
        _util_ctx[key] = ctx
    return _lib.load(), ctx


def tf_accept(x, Lx, px, *, u=None, seed=0, counter=0, chain_offset=0):
    """where(px - u >= 0, Lx, x) row-wise (utils/sampler.py:53-55); u drawn in-kernel when None (Philox keyed by the global
    chain id chain_offset + i)."""
    if not (x.is_cuda and Lx.is_cuda and px.is_cuda):
        raise TypeError("tf_accept works on CUDA tensors")
    x = x.detach().to(TORCH_FLOAT).contiguous()
    Lx = Lx.detach().to(TORCH_FLOAT).contiguous()
    px = px.detach().to(TORCH_FLOAT).contiguous()
    n, d = x.shape
    lib, ctx = _util(d, x.device.index)
    out = torch.empty_like(x)
    up = None
    if u is not None:
        u = u.detach().to(device=x.device, dtype=TORCH_FLOAT).contiguous()
        up = u.data_ptr()
    stream = C.c_void_p(torch.cuda.current_stream(x.device.index).cuda_stream)
    _lib.check(lib, ctx, lib.l2hmc_accept(ctx, n, int(chain_offset), x.data_ptr(), Lx.data_ptr(), px.data_ptr(), up, int(seed),
                                          int(counter), out.data_ptr(), None, stream))
    return out


def randn_like(x, *, seed=0, counter=0):
    """N(0, I) with the library generator (tf.random_normal(tf.shape(x)) in the reference)."""
    n, d = x.shape
    lib, ctx = _util(d, x.device.index)
    out = torch.empty((n, d), dtype=TORCH_FLOAT, device=x.device)
    stream = C.c_void_p(torch.cuda.current_stream(x.device.index).cuda_stream)
    _lib.check(lib, ctx, lib.l2hmc_philox_fill(ctx, n, 0, int(seed), int(counter), out.data_ptr(), None, None, stream))
    return out


def chain_operator(init_x, dynamics, nb_steps, aux=None, init_v=None, do_mh_step=False, *, rng=None):
    """Compose nb_steps proposals, accumulate log|J|, one MH at the end (utils/sampler.py:57-85).

    Kept quirk of the reference: sub-proposals draw fresh momentum (init_v only seeds the first
    Hamiltonian and makes ``propose`` return Lv); the final p_accept pairs init_v with the last Lv.
    rng: optional list (one dict per sub-proposal, see ``propose``) plus rng_final={'u': ...} as the
    last element for the closing MH step.

    ONE kernel launch where the fused kernels cover the problem (``l2hmc_transition_args.chain``: the sub-proposals, the
    accumulated log|J|, p_accept and the Metropolis step all stay on chip); on the layered engine (VAE target, wide nets)
    and the generic tensor-core kernel the same composition runs as a loop of launches.
    """
    nb_steps = int(nb_steps)
    fused = _chain_in_one_launch(init_x, dynamics, nb_steps, aux, init_v, do_mh_step, rng)
    if fused is not None:
        return fused
    if init_v is None:
        init_v = randn_like(init_x, seed=dynamics.seed ^ 0x5EED, counter=dynamics.next_counter())
    x, v = init_x, init_v
    log_jac = torch.zeros((init_x.shape[0],), dtype=TORCH_FLOAT, device=init_x.device)
    for t in range(nb_steps):
        r = rng[t] if rng is not None else None
        x, v, px, _ = propose(x, dynamics, init_v=v, aux=aux, log_jac=True, do_mh_step=False, rng=r)
        log_jac = log_jac + px
    p_accept = dynamics.p_accept(init_x, init_v, x, v, log_jac, aux=aux)
    outputs = []
    if do_mh_step:
        u = rng[nb_steps].get("u") if rng is not None and len(rng) > nb_steps else None
        outputs.append(tf_accept(init_x, x, p_accept, u=u, seed=dynamics.seed, counter=dynamics.next_counter()))
    return x, v, p_accept, outputs


def _chain_in_one_launch(init_x, dynamics, nb_steps, aux, init_v, do_mh_step, rng):
    """chain_operator through l2hmc_transition_args.chain, or None when this Dynamics runs on an engine without it."""
    if dynamics.hmc or aux is not None or nb_steps < 1 or dynamics.aux_dim:
        return None
    dynamics._ensure_ctx()
    if not dynamics.kernel_name.startswith(("small", "tile", "tc_")):
        return None
    v = d = u = None
    if rng is not None:
        steps = list(rng[:nb_steps])
        if len(steps) != nb_steps:
            raise ValueError("rng needs one dict per sub-proposal")
        have_v = [("v" in r) for r in steps]
        have_d = [("direction" in r) for r in steps]
        if any(have_v) != all(have_v) or any(have_d) != all(have_d):
            return None   # partially injected randomness: the loop of launches handles it
        if all(have_v):
            v = torch.stack([dynamics._prep(r["v"], "v", dynamics.x_dim) for r in steps])
        if all(have_d):
            d = torch.stack([r["direction"].detach().to(device=init_x.device, dtype=torch.uint8) for r in steps])
        if len(rng) > nb_steps:
            u = rng[nb_steps].get("u")
    try:
        o = dynamics._transition(init_x, v=v, dir_mode=_lib.DIR_RANDOM, direction=d, u=u, do_mh=True, n_transitions=nb_steps,
                                 chain=True, v0=init_v)
    except _lib.L2HMCError as e:
        if e.code == 3:   # L2HMC_EUNSUPPORTED: this shape runs on the generic tensor-core kernel
            return None
        raise
    return o["Lx"], o["Lv"], o["px"], ([o["x_next"]] if do_mh_step else [])
