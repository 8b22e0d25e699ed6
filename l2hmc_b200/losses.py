"""The reference's ``utils/losses.py`` names, value only.

In the reference these build TF graph nodes that autodiff later differentiates; here the gradient of a loss through the
sampler comes from ``training.loss_and_grads(dynamics, x, loss=name)`` (one C-ABI call, csrc/train.cuh), and these
functions give the VALUE of the same objective for given proposals ``(x, Lx, px)`` -- tensors on any device, a few
elementwise torch ops, not a hot path (monitoring, and the check that the two agree in tests/test_gpu_training.py).
"""
from __future__ import annotations

import math

import torch


def loss_vec(x, X, p):
    """utils/losses.py:36-37: expected squared jump distance per chain, + 1e-4."""
    return ((X - x) ** 2).sum(1) * p + 1e-4


def loss_logsumexp(x, X, p):
    """utils/losses.py:39-42."""
    v = loss_vec(x, X, p)
    return torch.logsumexp(-v, 0) - math.log(v.shape[0])


def loss_inverse(x, X, p):
    """utils/losses.py:44-47."""
    v = loss_vec(x, X, p)
    return -1.0 / (1.0 / (v + 1e-4)).mean()


def loss_std(x, X, p):
    """utils/losses.py:49-51."""
    return -loss_vec(x, X, p).mean(0)


def loss_mixed(x, Lx, px, scale=1.0):
    """utils/losses.py:53-59 (the notebook's objective is this with scale = 0.1 on two batches, SCGExperiment.ipynb:171-181)."""
    v1 = loss_vec(x, Lx, px) / scale
    return (1.0 / v1).mean() - v1.mean()


def get_loss(name):
    """utils/losses.py:26-34."""
    assoc = {"mixed": loss_mixed, "standard": loss_std, "inverse": loss_inverse, "logsumexp": loss_logsumexp}
    return assoc[name]
