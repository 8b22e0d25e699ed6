"""Annealed importance sampling with HMC-mode Dynamics (mirror of /root/reference/utils/ais.py:30-82).

``ais_estimate(init_energy, final_energy, anneal_steps, initial_x, ...)`` keeps the reference's positional
arguments.  Each annealing step runs on the GPU through the same library as the sampler: the energies, the
HMC-mode leapfrog transition (``Dynamics(hmc=True).forward``) and the Metropolis select (``tf_accept``).

What is covered:
  * both energies closed-form single Gaussians (``distributions.Gaussian``), for which ``(1 - beta) U0 + beta U1`` is again
    a Gaussian energy (precision ``(1-beta) S0 + beta S1``) up to an additive constant that cancels in every Hamiltonian
    difference -- one HMC context whose energy parameters are re-sent per beta;
  * the reference's own use (eval_vae.py:52-65): init_energy = standard normal prior, final_energy = the decoder posterior
    ``vae.DecoderEnergy`` with ``aux`` = the images.  ``(1-beta) 0.5|z|^2 + beta (BCE + 0.5|z|^2) = beta BCE + 0.5|z|^2``: the
    decoder energy with a likelihood weight (``l2hmc_set_likelihood_scale``), HMC mode on the layered engine.
  * any other pair of closed-form energies of ``distributions`` (Gaussian, GMM, RoughWell, funnel): the annealed energy
    is evaluated per chain inside the fused kernels as a mixed energy (``distributions.MixedEnergy``,
    ``l2hmc_set_energy_mixed``; beta moved per step with ``l2hmc_set_mix_beta``).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .distributions import EnergyFunction, MixedEnergy
from .dynamics import Dynamics, TORCH_FLOAT
from .sampler import randn_like, tf_accept


def _mixed_gaussian(e0: EnergyFunction, e1: EnergyFunction, beta: float) -> EnergyFunction:
    """(1 - beta) * 0.5 (x-m0) S0 (x-m0)^T + beta * 0.5 (x-m1) S1 (x-m1)^T = 0.5 (x-m) S (x-m)^T + const."""
    S0, S1 = e0.S[0].astype(np.float64), e1.S[0].astype(np.float64)
    S0, S1 = 0.5 * (S0 + S0.T), 0.5 * (S1 + S1.T)
    m0, m1 = e0.mu[0].astype(np.float64), e1.mu[0].astype(np.float64)
    S = (1.0 - beta) * S0 + beta * S1
    m = np.linalg.solve(S, (1.0 - beta) * S0.dot(m0) + beta * S1.dot(m1))
    return EnergyFunction(_lib.ENERGY_GAUSSIAN, e0.dim, mu=m[None, :], S=S[None, :, :])


def ais_estimate(init_energy, final_energy, anneal_steps, initial_x, aux=None, step_size=0.5, leapfrogs=25, x_dim=5,
                 num_splits=1, refresh=False, refreshment=0.1, *, rng: Optional[dict] = None, seed: int = 0,
                 return_state: bool = False):
    """utils/ais.py:30-82.  rng: optional explicit randomness {'v0' [N,D], 'v' [steps,N,D], 'u' [steps,N]} (parity
    tests); otherwise the library generator (Philox keyed by seed and step).  Returns (estimate, mean accept prob)
    like the reference, plus (x, w) when return_state."""
    decoder = getattr(final_energy, "kind", None) == _lib.ENERGY_DECODER
    if decoder:
        e0 = init_energy
        std = (isinstance(e0, EnergyFunction) and e0.kind == _lib.ENERGY_GAUSSIAN and e0.n_comp == 1 and
               np.allclose(e0.mu, 0.0) and np.allclose(e0.S[0], np.eye(e0.dim)))
        if not std:
            raise NotImplementedError("with a decoder final energy the initial energy must be the standard normal prior "
                                      "(eval_vae.py:49-50)")
        if aux is None:
            raise TypeError("the decoder posterior needs aux (the images)")
    else:
        for e in (init_energy, final_energy):
            if not isinstance(e, EnergyFunction):
                raise TypeError("ais_estimate needs closed-form energies from l2hmc_b200.distributions (got %r)" % (e,))
        closed = (_lib.ENERGY_GAUSSIAN, _lib.ENERGY_GMM, _lib.ENERGY_ROUGHWELL, _lib.ENERGY_FUNNEL)
        if not (init_energy.kind in closed and final_energy.kind in closed):
            raise NotImplementedError("ais_estimate anneals between closed-form energies of l2hmc_b200.distributions, or from "
                                      "the standard normal prior to the decoder posterior (eval_vae.py)")
        if aux is not None:
            raise NotImplementedError("aux is only consumed by the decoder posterior")
    x = initial_x.detach().to(TORCH_FLOAT).contiguous()
    if not x.is_cuda:
        raise TypeError("ais_estimate works on CUDA tensors")
    n, D = x.shape
    if D != int(x_dim):
        raise ValueError("initial_x is %d-d but x_dim=%d" % (D, x_dim))
    dev = x.device
    anneal_steps = int(anneal_steps)
    beta = np.linspace(0.0, 1.0, anneal_steps + 1, dtype=np.float32)[1:]          # tf.linspace(0., 1., steps+1)[1:]
    beta_diff = float(beta[1] - beta[0]) if anneal_steps > 1 else float(beta[0])  # beta[1] - beta[0]
    w = torch.zeros((n,), dtype=TORCH_FLOAT, device=dev)
    if rng is not None and "v0" in rng:
        v = torch.as_tensor(np.asarray(rng["v0"], dtype=np.float32)).to(dev)
    else:
        v = randn_like(x, seed=seed, counter=0)
    dyn = None
    both_gaussian = (not decoder and init_energy.kind == _lib.ENERGY_GAUSSIAN and final_energy.kind == _lib.ENERGY_GAUSSIAN and
                     init_energy.n_comp == 1 and final_energy.n_comp == 1)
    alpha_sum = torch.zeros((), dtype=torch.float64, device=dev)
    for s in range(anneal_steps):
        if rng is not None and "v" in rng:
            z = torch.as_tensor(np.asarray(rng["v"][s], dtype=np.float32)).to(dev)
        else:
            z = randn_like(x, seed=seed, counter=2 * s + 1)
        rv = v * math.sqrt(1.0 - refreshment) + z * math.sqrt(refreshment) if refresh else z
        if decoder:
            w = w + beta_diff * (-final_energy(x, aux=aux) + init_energy(x))
            if dyn is None:
                dyn = Dynamics(int(x_dim), final_energy, T=int(leapfrogs), eps=float(step_size), hmc=True, device=dev.index)
            dyn.set_likelihood_scale(float(beta[s]))   # curr_energy = beta * BCE + 0.5 |z|^2
            Lx, Lv, px = dyn.forward(x, init_v=rv, aux=aux)
        elif both_gaussian:
            w = w + beta_diff * (-final_energy(x) + init_energy(x))
            mixed = _mixed_gaussian(init_energy, final_energy, float(beta[s]))
            if dyn is None:
                dyn = Dynamics(int(x_dim), mixed, T=int(leapfrogs), eps=float(step_size), hmc=True, device=dev.index)
            else:
                dyn.set_energy_function(mixed)
            Lx, Lv, px = dyn.forward(x, init_v=rv)
        else:   # any two closed-form energies: (1 - beta) U0 + beta U1 evaluated inside the kernel
            w = w + beta_diff * (-final_energy(x) + init_energy(x))
            if dyn is None:
                dyn = Dynamics(int(x_dim), MixedEnergy(init_energy, final_energy, float(beta[s])), T=int(leapfrogs),
                               eps=float(step_size), hmc=True, device=dev.index)
            else:
                dyn.set_mix_beta(float(beta[s]))
            Lx, Lv, px = dyn.forward(x, init_v=rv)
        if rng is not None and "u" in rng:
            u = torch.as_tensor(np.asarray(rng["u"][s], dtype=np.float32)).to(dev)
        else:
            u = torch.rand((n,), device=dev, dtype=TORCH_FLOAT,
                           generator=torch.Generator(device=dev).manual_seed(int(seed) * 1000003 + 2 * s + 2))
        x = tf_accept(x, Lx, px, u=u)          # updated_x = where(mask, Lx, last_x)
        v = tf_accept(-Lv, Lv, px, u=u)        # updated_v = where(mask, Lv, -Lv)
        alpha_sum += px.double().mean()
    parts = torch.chunk(w.double(), int(num_splits), dim=0)
    est = sum(torch.logsumexp(p, 0) - math.log(p.shape[0]) for p in parts)
    out = (float(est), float(alpha_sum / max(anneal_steps, 1)))
    return out + (x, w) if return_state else out
