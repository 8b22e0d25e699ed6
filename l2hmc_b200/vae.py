"""The MNIST-VAE posterior target of the reference as an energy the CUDA path can take.

/root/reference/mnist_vae.py builds, inside its training graph,

    decoder = Sequential([Linear(latent, 1024), softplus, Linear(1024, 1024), softplus, Linear(1024, 784)])   # :104-111
    def energy(z, aux=None):                                                                                  # :122-126
        logits = decoder(z)
        log_posterior = -sum(sigmoid_cross_entropy_with_logits(labels=aux, logits=logits), axis=1)
        log_prior = -0.5 * sum(z**2, axis=1)
        return -log_posterior - log_prior

and hands ``energy`` to ``Dynamics`` together with a ``net_factory`` whose nets add an encoding of ``aux`` (the image
batch) to their first stage (:134-167).  ``DecoderEnergy(decoder)`` is that closure as a descriptor: the decoder
weights go to libl2hmc.so once (``l2hmc_set_energy_decoder``) and U / grad U are evaluated by the layered engine
(csrc/layered.cuh) -- forward through the decoder, the Bernoulli log-likelihood, then the hand-written reverse pass.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .layers import compile_softplus_mlp


class DecoderEnergy(object):
    """``energy(z, aux)`` of mnist_vae.py:122-126 for a Linear/softplus ``decoder`` Sequential."""

    kind = _lib.ENERGY_DECODER
    accepts_aux = True

    def __init__(self, decoder):
        self.decoder = decoder
        self.widths, self.Ws, self.bs = compile_softplus_mlp(decoder, "decoder")
        self.dim = int(self.widths[0])
        self.aux_dim = int(self.widths[-1])
        self.mu = self.S = self.logc = self.scalars = None
        self.n_comp = 1
        self._dyn = {}

    def refresh(self):
        """Re-read the decoder weights from the layer objects."""
        self.widths, self.Ws, self.bs = compile_softplus_mlp(self.decoder, "decoder")
        self._dyn = {}

    def _evaluator(self, device_index: int):
        from .dynamics import Dynamics
        d = self._dyn.get(device_index)
        if d is None:
            d = Dynamics(self.dim, self, T=1, eps=0.1, hmc=True, device=device_index)
            self._dyn[device_index] = d
        return d

    def __call__(self, z, aux=None):
        if aux is None:
            raise TypeError("energy(z, aux): the decoder target needs the aux rows (labels of the Bernoulli likelihood)")
        if not isinstance(z, torch.Tensor) or not z.is_cuda:
            raise TypeError("energy functions evaluate on CUDA tensors (got %r)" % (type(z),))
        return self._evaluator(z.device.index).energy(z, aux=aux)


def bernoulli_aux(n: int, aux_dim: int = 784, rng=None) -> np.ndarray:
    """Synthetic stand-in for a batch of binarised MNIST images (there is no dataset in this environment)."""
    rng = np.random.default_rng(0) if rng is None else rng
    return (rng.random((n, aux_dim)) < 0.5).astype(np.float32)
