"""l2hmc_b200 -- B200-native L2HMC augmented-leapfrog sampling path.

Drop-in names for the reference's ``utils`` package on this path:
    from l2hmc_b200.dynamics import Dynamics
    from l2hmc_b200.sampler import propose, tf_accept, chain_operator
    from l2hmc_b200.distributions import Gaussian, GMM, RoughWell, GaussianFunnel, gen_ring
    from l2hmc_b200.layers import Linear, Sequential, Zip, Parallel, ScaleTanh, relu, softplus
    from l2hmc_b200.vae import DecoderEnergy                      # energy(z, aux) of mnist_vae.py
    from l2hmc_b200.diagnostics import acl_spectrum, ESS, autocovariance, sample_trace   # utils/func_utils.py
    from l2hmc_b200.ais import ais_estimate                       # utils/ais.py (Gaussian -> Gaussian annealing)
    from l2hmc_b200.training import loss_and_grads, Adam, train_step   # the notebook's objective / optimiser loop (first-correct)
    from l2hmc_b200.losses import get_loss, loss_mixed              # utils/losses.py names (values)
"""
from . import _lib, layers, distributions, philox, vae, diagnostics, ais, training, losses  # noqa: F401
from .dynamics import Dynamics  # noqa: F401
from .sampler import propose, tf_accept, chain_operator  # noqa: F401

__all__ = ["Dynamics", "propose", "tf_accept", "chain_operator", "layers", "distributions", "philox", "vae",
           "diagnostics", "ais", "training", "losses"]
