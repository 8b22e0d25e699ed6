"""Target distributions (mirror of the reference's utils/distributions.py).

Same constructors and methods as /root/reference/utils/distributions.py: ``Gaussian`` (:41-68),
``TiltedGaussian`` (:70-82), ``random_tilted_gaussian`` (:34-39), ``RoughWell`` (:84-101), ``GMM``
(:104-150), ``GaussianFunnel`` (:155-199), ``gen_ring`` (:201-213).  ``get_energy_function()`` still
returns a callable ``fn(x[, aux=])`` -> [N]; the callable additionally carries the closed-form
descriptor (kind + parameters) that ``Dynamics`` hands to the CUDA library, because a fused kernel
cannot call back into Python the way ``tf.gradients`` walks a TF closure (utils/dynamics.py:217-218).
Calling it evaluates the energy on the GPU through the same library.
"""
from __future__ import annotations

import collections
from typing import List, Optional, Sequence

import numpy as np
import torch
from scipy.stats import multivariate_normal, ortho_group

from . import _lib


class EnergyFunction(object):
    """Callable energy closure with its analytic descriptor.

    kind: one of _lib.ENERGY_*; mu [K,D], S [K,D,D] (fp32 precision matrices, exactly the
    ``i_sigma.astype('float32')`` constants of the reference), logc [K], scalars (kind specific).
    """

    def __init__(self, kind: int, dim: int, mu=None, S=None, logc=None, scalars=None, accepts_aux=True):
        self.kind = int(kind)
        self.dim = int(dim)
        self.mu = None if mu is None else np.ascontiguousarray(mu, dtype=np.float32)
        self.S = None if S is None else np.ascontiguousarray(S, dtype=np.float32)
        self.logc = None if logc is None else np.ascontiguousarray(logc, dtype=np.float32)
        self.scalars = None if scalars is None else np.ascontiguousarray(scalars, dtype=np.float32)
        self.n_comp = 1 if self.mu is None else int(self.mu.shape[0])
        self.accepts_aux = accepts_aux
        self._dyn = {}

    def _evaluator(self, device_index: int):
        from .dynamics import Dynamics  # local import: dynamics imports this module
        d = self._dyn.get(device_index)
        if d is None:
            d = Dynamics(self.dim, self, T=1, eps=0.1, hmc=True, device=device_index)
            self._dyn[device_index] = d
        return d

    def __call__(self, x, *args, **kwargs):
        if not self.accepts_aux and (args or kwargs):
            raise TypeError("fn() takes exactly 1 argument")  # GMM's fn(x) (utils/distributions.py:126)
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise TypeError("energy functions evaluate on CUDA tensors (got %r)" % (type(x),))
        return self._evaluator(x.device.index).energy(x)


class MixedEnergy(EnergyFunction):
    """``curr_energy`` of utils/ais.py:44-45: (1 - beta) * init_energy(z) + beta * final_energy(z) for two closed-form
    energies of this module; evaluated per chain inside the fused kernels (l2hmc_set_energy_mixed)."""

    def __init__(self, a: EnergyFunction, b: EnergyFunction, beta: float = 0.0):
        closed = (_lib.ENERGY_GAUSSIAN, _lib.ENERGY_GMM, _lib.ENERGY_ROUGHWELL, _lib.ENERGY_FUNNEL)
        if not (isinstance(a, EnergyFunction) and isinstance(b, EnergyFunction) and a.kind in closed and b.kind in closed):
            raise TypeError("MixedEnergy mixes two closed-form energies of l2hmc_b200.distributions")
        if a.dim != b.dim:
            raise ValueError("the two energies live in different dimensions (%d, %d)" % (a.dim, b.dim))
        EnergyFunction.__init__(self, _lib.ENERGY_MIXED, a.dim)
        self.a, self.b, self.beta = a, b, float(beta)


def random_tilted_gaussian(dim, log_min=-2., log_max=2.):
    mu = np.zeros((dim,))
    R = ortho_group.rvs(dim)
    sigma = np.diag(np.exp(np.log(10.) * np.random.uniform(log_min, log_max, size=(dim,)))) + 1e-6 * np.eye(dim)
    S = R.T.dot(sigma).dot(R)
    return Gaussian(mu, S)


class Gaussian(object):
    def __init__(self, mu, sigma):
        self.mu = np.asarray(mu)
        self.sigma = np.asarray(sigma)
        self.i_sigma = np.linalg.inv(np.copy(self.sigma))  # fp64, cast to fp32 on use (:48,52)

    def get_energy_function(self):
        return EnergyFunction(_lib.ENERGY_GAUSSIAN, self.sigma.shape[0],
                              mu=self.mu.astype('float32')[None, :],
                              S=self.i_sigma.astype('float32')[None, :, :])

    def get_samples(self, n):
        C = np.linalg.cholesky(self.sigma)
        X = np.random.randn(n, self.sigma.shape[0])
        return X.dot(C.T)

    def log_density(self, X):
        return multivariate_normal(mean=self.mu, cov=self.sigma).logpdf(X)


class TiltedGaussian(Gaussian):
    def __init__(self, dim, log_min, log_max):
        self.R = ortho_group.rvs(dim)
        self.diag = np.diag(np.exp(np.log(10.) * np.random.uniform(log_min, log_max, size=(dim,)))) + 1e-8 * np.eye(dim)
        S = self.R.T.dot(self.diag).dot(self.R)
        self.dim = dim
        Gaussian.__init__(self, np.zeros((dim,)), S)

    def get_samples(self, n):
        # the reference draws a fixed 200 rows here regardless of n (utils/distributions.py:79)
        X = np.random.randn(n, self.dim)
        X = X.dot(np.sqrt(self.diag))
        X = X.dot(self.R)
        return X


class RoughWell(object):
    def __init__(self, dim, eps, easy=False):
        self.dim = dim
        self.eps = eps
        self.easy = easy

    def get_energy_function(self):
        den = self.eps if self.easy else self.eps * self.eps  # python double, rounded to fp32 once (:92-96)
        return EnergyFunction(_lib.ENERGY_ROUGHWELL, self.dim, scalars=[self.eps, den])

    def get_samples(self, n):
        # we can approximate by a gaussian for eps small enough
        return np.random.randn(n, self.dim)


class GMM(object):
    def __init__(self, mus, sigmas, pis):
        assert len(mus) == len(sigmas)
        assert sum(pis) == 1.0

        self.mus = mus
        self.sigmas = sigmas
        self.pis = pis
        self.nb_mixtures = len(pis)
        self.k = mus[0].shape[0]
        self.i_sigmas = []
        self.constants = []
        for i, sigma in enumerate(sigmas):
            self.i_sigmas.append(np.linalg.inv(sigma).astype('float32'))
            det = np.sqrt((2 * np.pi) ** self.k * np.linalg.det(sigma)).astype('float32')
            self.constants.append((pis[i] / det).astype('float32'))

    def get_energy_function(self):
        return EnergyFunction(_lib.ENERGY_GMM, self.k,
                              mu=np.stack([np.asarray(m, dtype=np.float32) for m in self.mus]),
                              S=np.stack(self.i_sigmas),
                              logc=np.log(np.asarray(self.constants, dtype=np.float32)),  # tf.log of the fp32 constant (:129)
                              accepts_aux=False)

    def get_samples(self, n):
        categorical = np.random.choice(self.nb_mixtures, size=(n,), p=self.pis)
        counter_samples = collections.Counter(categorical)
        samples = []
        for k, v in counter_samples.items():
            samples.append(np.random.multivariate_normal(self.mus[k], self.sigmas[k], size=(v,)))
        samples = np.concatenate(samples, axis=0)
        np.random.shuffle(samples)
        return samples

    def log_density(self, X):
        return np.log(sum([self.pis[i] * multivariate_normal(mean=self.mus[i], cov=self.sigmas[i]).pdf(X)
                           for i in range(self.nb_mixtures)]))


class GaussianFunnel(object):
    def __init__(self, dim=2, clip=6.):
        self.dim = dim
        self.sigma = 2.0
        self.clip = 4 * self.sigma  # the ctor argument is ignored by the reference too (:156-159)

    def get_energy_function(self):
        return EnergyFunction(_lib.ENERGY_FUNNEL, self.dim, scalars=[self.sigma, self.clip], accepts_aux=False)

    def get_samples(self, n):
        samples = np.zeros((n, self.dim))
        for t in range(n):
            v = self.sigma * np.random.randn()
            s = np.exp(v / 2)
            samples[t, 0] = v
            samples[t, 1:] = s * np.random.randn(self.dim - 1)
        return samples

    def log_density(self, x):
        v = x[:, 0]
        log_p_v = np.square(v / self.sigma)
        s = np.exp(v)
        sum_sq = np.square(x[:, 1:]).sum(axis=1)
        n = x.shape[1] - 1
        return 0.5 * (log_p_v + sum_sq / s + (n / 2) * np.log(2 * np.pi * s))


def gen_ring(r=1.0, var=1.0, nb_mixtures=2):
    base_points = []
    for t in range(nb_mixtures):
        c = np.cos(2 * np.pi * t / nb_mixtures)
        s = np.sin(2 * np.pi * t / nb_mixtures)
        base_points.append(np.array([r * c, r * s]))
    sigmas = [var * np.eye(2) for t in range(nb_mixtures)]
    pis = [1. / nb_mixtures] * nb_mixtures
    pis[0] += 1 - sum(pis)
    return GMM(base_points, sigmas, pis)
