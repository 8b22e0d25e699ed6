"""Sampler diagnostics on the GPU (mirror of the reference's utils/func_utils.py:45-54,114-120).

The notebook collects 2000 x [200, 2] samples on the host with one ``sess.run`` per transition and then calls
``acl_spectrum`` / ``ESS`` in numpy (SCGExperiment.ipynb:291-298,331-334,388).  Here the trace stays in HBM:
``sample_trace`` writes every transition's samples straight into one device tensor (transition t reads trace[t-1],
writes trace[t]; no host synchronisation in the loop) and ``acl_spectrum`` reduces it on the device
(``l2hmc_acl_spectrum``), so only n_steps-1 doubles ever reach the host.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .sampler import _util


def sample_trace(x, dynamics, n_steps, aux=None, stats=None):
    """Run ``n_steps`` transitions (propose + MH) and return the device-resident trace [n_steps, N, x_dim] of the
    chain states after each one -- the loop of SCGExperiment.ipynb:291-298 without leaving the GPU: ONE launch of the
    fused kernel iterates the chain on-chip and writes every transition's Metropolis output into the trace
    (``l2hmc_transition_args.trace``).  ``stats``: optional CUDA float64 [2] accumulator of (sum of accept
    probabilities, number accepted) over the whole trace."""
    n, d = x.shape
    n_steps = int(n_steps)
    trace = torch.empty((n_steps, n, d), dtype=torch.float32, device=x.device)
    if n_steps == 0:
        return trace
    dir_mode = _lib.DIR_FORWARD if dynamics.hmc else _lib.DIR_RANDOM
    dynamics._transition(x, dir_mode=dir_mode, do_mh=True, want_v=False, n_transitions=n_steps, aux=aux, trace=trace,
                         stats=stats)
    return trace


def acl_spectrum(X, scale, n_lags=None):
    """[autocovariance(X / scale, tau) for tau in range(n - 1)] (utils/func_utils.py:114-116) for a CUDA fp32 trace
    X [n_steps, N, x_dim]; returns a CUDA float64 tensor.  ``n_lags`` (default n_steps - 1, like the reference)
    truncates the spectrum."""
    if not isinstance(X, torch.Tensor) or not X.is_cuda or X.dim() != 3:
        raise TypeError("acl_spectrum works on a CUDA tensor [n_steps, N, x_dim]")
    X = X.detach().to(torch.float32).contiguous()
    S, n, d = X.shape
    L = (S - 1) if n_lags is None else int(n_lags)
    if L < 1:
        return torch.empty((0,), dtype=torch.float64, device=X.device)
    lib, ctx = _util(d, X.device.index)
    out = torch.empty((L,), dtype=torch.float64, device=X.device)
    stream = C.c_void_p(torch.cuda.current_stream(X.device.index).cuda_stream)
    _lib.check(lib, ctx, lib.l2hmc_acl_spectrum(ctx, S, n, X.data_ptr(), float(scale), L, out.data_ptr(), stream))
    return out


def autocovariance(X, tau=0):
    """mean_t( sum(X[t] * X[t + tau]) / N ) (utils/func_utils.py:45-54) of a CUDA trace, as a python float."""
    if not isinstance(X, torch.Tensor) or not X.is_cuda or X.dim() != 3:
        raise TypeError("autocovariance works on a CUDA tensor [n_steps, N, x_dim]")
    X = X.detach().to(torch.float32).contiguous()
    S = X.shape[0]
    tau = int(tau)
    if not 0 <= tau < S:
        raise ValueError("tau must be in [0, n_steps)")
    # only lag tau is needed: the lag-0 spectrum of the trace against its shifted self
    a, b = X[: S - tau].to(torch.float64), X[tau:].to(torch.float64)
    return float((a * b).sum() / X.shape[1] / (S - tau))


def ESS(A):
    """1 / (1 + 2 sum(A[1:] * (A[1:] > 0.05))) (utils/func_utils.py:118-120); A: tensor or array-like spectrum."""
    A = torch.as_tensor(A)
    A = A * (A > 0.05)
    return float(1. / (1. + 2 * A[1:].sum()))
