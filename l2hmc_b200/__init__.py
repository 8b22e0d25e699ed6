"""l2hmc_b200 -- B200-native L2HMC augmented-leapfrog sampling path.

Drop-in names for the reference's ``utils`` package on this path:
    from l2hmc_b200.dynamics import Dynamics
    from l2hmc_b200.sampler import propose, tf_accept, chain_operator
    from l2hmc_b200.distributions import Gaussian, GMM, RoughWell, GaussianFunnel, gen_ring
    from l2hmc_b200.layers import Linear, Sequential, Zip, Parallel, ScaleTanh, relu
"""
from . import _lib, layers, distributions, philox  # noqa: F401
from .dynamics import Dynamics  # noqa: F401
from .sampler import propose, tf_accept, chain_operator  # noqa: F401

__all__ = ["Dynamics", "propose", "tf_accept", "chain_operator", "layers", "distributions", "philox"]
