"""CPU oracle for the L2HMC augmented-leapfrog sampling path.  TEST INFRASTRUCTURE ONLY.

This is an op-for-op CPU restatement (torch-CPU, fp32 twin and fp64 twin) of the
reference's hot path.  It is the checker: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``l2hmc_b200/`` imports it and the
product path never falls back to it.

PARITY PINNED TO THE REFERENCE'S OWN CODE (round 2): real TensorFlow 1.x / Python 2 are absent here, but the
reference's hot-path files run UNMODIFIED on an eager torch-backed TensorFlow stand-in (oracle/tf_shim,
oracle/ref_loader.py, oracle/ref_runner.py).  tests/golden/make_ref_golden.py runs them (float64 and float32) on
injected parameters and randomness and commits the outputs (tests/golden/ref_*.npz, tests/golden/ref/*.npz);
tests/test_reference_pin.py asserts this oracle reproduces every one of them -- propose on all BASELINE targets,
each Dynamics method, chain_operator, the notebook's training objective with tf.gradients w.r.t. every variable,
utils/losses.py, the numpy diagnostics, utils/ais.py, and mnist_vae.py's own decoder / energy / sampler-net text --
to 1e-9 in fp64 (fp32 twin: bit-equal samples on most cases).  The reference itself holds no tests, golden vectors
or seeds; further anchors: the algebraic properties its code implies (exact inverse, log-det
vs autograd Jacobian, HMC limit; tests/test_oracle.py) and a second, independently written plain-C restatement
(oracle/l2hmc_oracle.c).

All randomness is injected (the reference draws v, direction bits and accept
uniforms from unseeded TF RNGs: utils/dynamics.py:248,276, utils/sampler.py:34,54).

Reference citations are into /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# utils/dynamics.py:101 -- a python float meeting an fp32 tensor: TF1 makes it an fp32 constant.  The fp64 twin is the
# fp32 graph in wider arithmetic, so it keeps the fp32-rounded value (as the reference run on oracle/tf_shim does).
TWO_PI = float(np.float32(2 * np.pi))


# --------------------------------------------------------------------------------------
# S/T/Q net (SCGExperiment.ipynb:51-77, utils/layers.py:29-37,81-95)
# --------------------------------------------------------------------------------------
NET_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4",
            "Ws", "bs", "Wt", "bt", "Wq", "bq", "ls", "lq")


def net_cast(p: Dict[str, torch.Tensor], dtype) -> Dict[str, torch.Tensor]:
    return {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in p.items()}


def net_apply(p, a, b, tau, aux_embed=None):
    """net([a, b, t, aux]) -> [S, T, Q].

    Zip of three Linear layers (+ the aux branch), python ``sum`` (order
    (((0 + e1) + e2) + e3) + e_aux, SCGExperiment.ipynb:54-60), relu, Linear, relu,
    Parallel[Linear+ScaleTanh, Linear, Linear+ScaleTanh] (SCGExperiment.ipynb:62-74).
    Linear is ``tf.add(tf.matmul(x, W), b)`` (utils/layers.py:37); ScaleTanh is
    ``exp(scale) * tanh(x)`` (utils/layers.py:83-86).  tau is [N, 2].
    """
    e1 = a @ p["W1"] + p["b1"]
    e2 = b @ p["W2"] + p["b2"]
    e3 = tau @ p["W3"] + p["b3"]
    e4 = 0.0 if aux_embed is None else aux_embed  # `lambda _: 0.` in the notebook
    h = (((0 + e1) + e2) + e3) + e4
    h = torch.relu(h)
    h = torch.relu(h @ p["W4"] + p["b4"])
    S = torch.exp(p["ls"]) * torch.tanh(h @ p["Ws"] + p["bs"])
    T = h @ p["Wt"] + p["bt"]
    Q = torch.exp(p["lq"]) * torch.tanh(h @ p["Wq"] + p["bq"])
    return S, T, Q


def net_zero(p, a, b, tau, aux_embed=None):
    """hmc=True nets: three zero tensors shaped like the first input (utils/dynamics.py:73-76)."""
    z = torch.zeros_like(a)
    return z, z, z


# --------------------------------------------------------------------------------------
# Energies (utils/distributions.py)
# --------------------------------------------------------------------------------------
class Energy:
    """energy(x) -> [N]; grad(x) -> [N, D] (what tf.gradients(energy, x) yields)."""

    def energy(self, x):  # pragma: no cover - interface
        raise NotImplementedError

    def grad(self, x):  # pragma: no cover - interface
        raise NotImplementedError

    def to(self, dtype):  # pragma: no cover - interface
        raise NotImplementedError

    def grad_autodiff(self, x):
        """Mirror of Dynamics.grad_energy (utils/dynamics.py:217-218): reverse mode through energy()."""
        xr = x.detach().clone().requires_grad_(True)
        e = self.energy(xr)
        (g,) = torch.autograd.grad(e.sum(), xr)
        return g


def _quad(x, mu, S):
    """quadratic_gaussian (utils/distributions.py:31-32) without the N x N intermediate:
    diag(0.5 * ((x-mu) S) (x-mu)^T) == 0.5 * rowsum(((x-mu) S) * (x-mu))."""
    d = x - mu
    return 0.5 * ((d @ S) * d).sum(1)


class GaussianEnergy(Energy):
    """Gaussian.get_energy_function (utils/distributions.py:50-57). S = inv(sigma) computed in
    fp64 and cast to fp32 (utils/distributions.py:48,52)."""

    def __init__(self, mu, S, dtype=torch.float32):
        self.mu = torch.as_tensor(np.asarray(mu, dtype=np.float32)).to(dtype)
        self.S = torch.as_tensor(np.asarray(S, dtype=np.float32)).to(dtype)

    def to(self, dtype):
        return GaussianEnergy(self.mu.numpy(), self.S.numpy(), dtype)

    def energy(self, x):
        return _quad(x, self.mu, self.S)

    def grad(self, x):
        # reverse mode of 0.5 * (dS) d^T: 0.5 * d S^T + 0.5 * d S
        d = x - self.mu
        return 0.5 * (d @ self.S.T) + 0.5 * (d @ self.S)


class GMMEnergy(Energy):
    """GMM.get_energy_function (utils/distributions.py:125-134): U = -logsumexp_i(-q_i(x) + log c_i),
    c_i = pi_i / sqrt((2 pi)^k det Sigma_i) as fp32 (utils/distributions.py:120-123)."""

    def __init__(self, mus, Ss, cs, dtype=torch.float32):
        self.mus = [torch.as_tensor(np.asarray(m, dtype=np.float32)).to(dtype) for m in mus]
        self.Ss = [torch.as_tensor(np.asarray(s, dtype=np.float32)).to(dtype) for s in Ss]
        self.cs = [torch.as_tensor(np.float32(c)).to(dtype) for c in cs]
        self.dtype = dtype

    def to(self, dtype):
        return GMMEnergy([m.numpy() for m in self.mus], [s.numpy() for s in self.Ss],
                         [float(c) for c in self.cs], dtype)

    def _V(self, x):
        return torch.stack([-_quad(x, m, S) + torch.log(c)
                            for m, S, c in zip(self.mus, self.Ss, self.cs)], dim=1)

    def energy(self, x):
        return -torch.logsumexp(self._V(x), dim=1)

    def grad(self, x):
        w = torch.softmax(self._V(x), dim=1)  # d(-logsumexp)/dV_i = -softmax_i ; dV_i/dx = -grad q_i
        g = torch.zeros_like(x)
        for i, (m, S) in enumerate(zip(self.mus, self.Ss)):
            d = x - m
            g = g + w[:, i:i + 1] * (0.5 * (d @ S.T) + 0.5 * (d @ S))
        return g


class RoughWellEnergy(Energy):
    """RoughWell.get_energy_function (utils/distributions.py:90-97)."""

    def __init__(self, eps, easy=False, dtype=torch.float32):
        self.eps = float(eps)
        self.easy = bool(easy)
        self.dtype = dtype

    def to(self, dtype):
        return RoughWellEnergy(self.eps, self.easy, dtype)

    def _scale(self, x):
        # python floats meet fp32 tensors in the reference (utils/distributions.py:92-96): eps and
        # eps*eps (evaluated in double) are each rounded to fp32 once; the fp64 twin keeps those values
        e = torch.tensor(np.float32(self.eps)).to(x.dtype)
        den = e if self.easy else torch.tensor(np.float32(self.eps * self.eps)).to(x.dtype)
        return e, den

    def energy(self, x):
        e, den = self._scale(x)
        n = (x * x).sum(1)
        return 0.5 * n + e * torch.cos(x / den).sum(1)

    def grad(self, x):
        e, den = self._scale(x)
        return x - e * torch.sin(x / den) / den


class FunnelEnergy(Energy):
    """GaussianFunnel.get_energy_function (utils/distributions.py:161-180); sigma=2, clip=4*sigma
    (the ctor's clip argument is ignored, utils/distributions.py:156-159)."""

    def __init__(self, sigma=2.0, clip=8.0, dtype=torch.float32):
        self.sigma = float(sigma)
        self.clip = float(clip)
        self.dtype = dtype

    def to(self, dtype):
        return FunnelEnergy(self.sigma, self.clip, dtype)

    def energy(self, x):
        dt = x.dtype
        v = x[:, 0]
        log_p_v = (v / self.sigma) ** 2
        s = torch.exp(v)
        sum_sq = (x[:, 1:] ** 2).sum(1)
        n = torch.tensor(float(x.shape[1] - 1), dtype=dt)
        two_pi = torch.tensor(TWO_PI, dtype=dt)  # 2.0 * np.pi * s (utils/distributions.py:167): fp32 constant
        E = 0.5 * (log_p_v + sum_sq / s + n * torch.log(two_pi * s))
        s_min = torch.exp(torch.tensor(-self.clip, dtype=dt))
        s_max = torch.exp(torch.tensor(self.clip, dtype=dt))
        E1 = 0.5 * (log_p_v + sum_sq / s_max + n * torch.log(two_pi * s_max))
        E2 = 0.5 * (log_p_v + sum_sq / s_min + n * torch.log(two_pi * s_min))
        E_ = torch.where(v > self.clip, E1, E)
        E_ = torch.where(-self.clip > v, E2, E_)
        return E_

    def grad(self, x):
        dt = x.dtype
        v = x[:, 0]
        s = torch.exp(v)
        sum_sq = (x[:, 1:] ** 2).sum(1)
        n = float(x.shape[1] - 1)
        s_min = torch.exp(torch.tensor(-self.clip, dtype=dt))
        s_max = torch.exp(torch.tensor(self.clip, dtype=dt))
        hi = v > self.clip
        lo = -self.clip > v
        s_eff = torch.where(lo, s_min, torch.where(hi, s_max, s))
        gv_in = v / (self.sigma ** 2) + 0.5 * (-sum_sq / s + n)
        gv_out = v / (self.sigma ** 2)
        gv = torch.where(hi | lo, gv_out, gv_in)
        g = x / s_eff[:, None]
        g = g.clone()
        g[:, 0] = gv
        return g


# --------------------------------------------------------------------------------------
# VAE posterior target and aux-conditioned nets (mnist_vae.py)
# --------------------------------------------------------------------------------------
def softplus_mlp(Ws, bs, x):
    """Sequential([Linear, softplus, ..., Linear]) (mnist_vae.py:104-111, 134-140): softplus between
    layers, none after the last.  tf.nn.softplus(x) = log(1 + exp(x)), evaluated overflow-free."""
    h = x
    for i, (W, b) in enumerate(zip(Ws, bs)):
        h = h @ W + b
        if i + 1 < len(Ws):
            h = torch.clamp(h, min=0) + torch.log1p(torch.exp(-torch.abs(h)))
    return h


class DecoderBernoulliEnergy(Energy):
    """energy(z, aux) of mnist_vae.py:122-126: sum_pix sigmoid_cross_entropy_with_logits(labels=aux,
    logits=decoder(z)) + 0.5 |z|^2, decoder = Linear/softplus/Linear/softplus/Linear (:104-111).
    TF's sigmoid_cross_entropy_with_logits is max(l, 0) - l*z + log(1 + exp(-|l|)).
    The per-chain aux rows are held by the object (Dynamics passes aux= through, utils/dynamics.py:209-212)."""

    def __init__(self, Ws, bs, aux, dtype=torch.float32):
        self.Ws = [torch.as_tensor(np.asarray(W)).to(dtype) for W in Ws]
        self.bs = [torch.as_tensor(np.asarray(b)).to(dtype) for b in bs]
        self.aux = torch.as_tensor(np.asarray(aux)).to(dtype)
        self.dtype = dtype

    def to(self, dtype):
        return DecoderBernoulliEnergy(self.Ws, self.bs, self.aux, dtype)

    def select(self, idx):
        e = DecoderBernoulliEnergy(self.Ws, self.bs, self.aux[idx], self.dtype)
        return e

    def energy(self, z):
        l = softplus_mlp(self.Ws, self.bs, z)
        bce = torch.clamp(l, min=0) - l * self.aux + torch.log1p(torch.exp(-torch.abs(l)))
        return bce.sum(1) + 0.5 * (z * z).sum(1)

    def grad(self, z):
        """Reverse mode written out: dU/dl = sigmoid(l) - aux; softplus' = sigmoid(pre)."""
        pres, hs = [], [z]
        h = z
        for i, (W, b) in enumerate(zip(self.Ws, self.bs)):
            pre = h @ W + b
            pres.append(pre)
            if i + 1 < len(self.Ws):
                h = torch.clamp(pre, min=0) + torch.log1p(torch.exp(-torch.abs(pre)))
                hs.append(h)
        d = torch.sigmoid(pres[-1]) - self.aux
        for i in range(len(self.Ws) - 1, -1, -1):
            d = d @ self.Ws[i].T
            if i > 0:
                d = d * torch.sigmoid(pres[i - 1])
        return d + z


from l2hmc_b200.synthetic import make_softplus_mlp  # noqa: E402,F401


# --------------------------------------------------------------------------------------
# Dynamics (utils/dynamics.py)
# --------------------------------------------------------------------------------------
@dataclass
class OracleDynamics:
    """State of a reference ``Dynamics`` object (utils/dynamics.py:35-81)."""
    x_dim: int
    T: int
    eps: float
    energy_obj: Energy
    mask: np.ndarray                       # [T, D] of {0,1}  (utils/dynamics.py:84-93)
    xnet: Optional[Dict[str, torch.Tensor]] = None
    vnet: Optional[Dict[str, torch.Tensor]] = None
    hmc: bool = False
    temperature: float = 1.0               # utils/dynamics.py:204-207 (1.0 unless use_temperature)
    dtype: torch.dtype = torch.float32
    _m: torch.Tensor = field(init=False, repr=False)

    def __post_init__(self):
        self._m = torch.as_tensor(np.asarray(self.mask, dtype=np.float32)).to(self.dtype)
        self.energy_obj = self.energy_obj.to(self.dtype)
        if self.xnet is not None:
            self.xnet = net_cast(self.xnet, self.dtype)
        if self.vnet is not None:
            self.vnet = net_cast(self.vnet, self.dtype)
        # eps = exp(alpha), alpha = log(eps) in fp32 (utils/dynamics.py:50-58)
        e32 = torch.exp(torch.log(torch.tensor(self.eps, dtype=torch.float32)))
        self._eps = e32.to(self.dtype)

    # -- pieces -------------------------------------------------------------------------
    def _XNet(self, a, b, tau, ae):
        return net_zero(None, a, b, tau) if self.hmc else net_apply(self.xnet, a, b, tau, ae)

    def _VNet(self, a, b, tau, ae):
        return net_zero(None, a, b, tau) if self.hmc else net_apply(self.vnet, a, b, tau, ae)

    def format_time(self, t: float, n: int):
        """_format_time (utils/dynamics.py:99-105)."""
        t_ = torch.tensor(float(t), dtype=self.dtype)
        arg = torch.tensor(TWO_PI, dtype=self.dtype) * t_ / torch.tensor(float(self.T), dtype=self.dtype)
        trig = torch.stack([torch.cos(arg), torch.sin(arg)])
        return trig[None, :].repeat(n, 1)

    def kinetic(self, v):
        return 0.5 * (v * v).sum(1)  # utils/dynamics.py:107-108

    def energy(self, x):
        return self.energy_obj.energy(x) / torch.tensor(np.float32(self.temperature)).to(self.dtype)  # :203-212

    def grad_energy(self, x):
        return self.energy_obj.grad(x) / torch.tensor(np.float32(self.temperature)).to(self.dtype)  # :217-218

    def hamiltonian(self, x, v):
        return self.energy(x) + self.kinetic(v)  # :214-215

    # -- steps --------------------------------------------------------------------------
    def forward_step(self, x, v, step: int, ae_x=None, ae_v=None):
        """_forward_step (utils/dynamics.py:115-157)."""
        eps = self._eps
        t = self.format_time(step, x.shape[0])
        grad1 = self.grad_energy(x)
        S1 = self._VNet(x, grad1, t, ae_v)
        sv1 = 0.5 * eps * S1[0]
        tv1 = S1[1]
        fv1 = eps * S1[2]
        v_h = v * torch.exp(sv1) + 0.5 * eps * (-(torch.exp(fv1) * grad1) + tv1)
        m = self._m[int(step)]
        mb = 1.0 - m
        X1 = self._XNet(v_h, m * x, t, ae_x)
        sx1 = eps * X1[0]
        tx1 = X1[1]
        fx1 = eps * X1[2]
        y = m * x + mb * (x * torch.exp(sx1) + eps * (torch.exp(fx1) * v_h + tx1))
        X2 = self._XNet(v_h, mb * y, t, ae_x)
        sx2 = eps * X2[0]
        tx2 = X2[1]
        fx2 = eps * X2[2]
        x_o = mb * y + m * (y * torch.exp(sx2) + eps * (torch.exp(fx2) * v_h + tx2))
        grad2 = self.grad_energy(x_o)
        S2 = self._VNet(x_o, grad2, t, ae_v)
        sv2 = 0.5 * eps * S2[0]
        tv2 = S2[1]
        fv2 = eps * S2[2]
        v_o = v_h * torch.exp(sv2) + 0.5 * eps * (-(torch.exp(fv2) * grad2) + tv2)
        log_jac = (sv1 + sv2 + mb * sx1 + m * sx2).sum(1)
        return x_o, v_o, log_jac

    def backward_step(self, x_o, v_o, step: int, ae_x=None, ae_v=None):
        """_backward_step (utils/dynamics.py:159-201)."""
        eps = self._eps
        t = self.format_time(step, x_o.shape[0])
        grad1 = self.grad_energy(x_o)
        S1 = self._VNet(x_o, grad1, t, ae_v)
        sv2 = -0.5 * eps * S1[0]
        tv2 = S1[1]
        fv2 = eps * S1[2]
        v_h = (v_o - 0.5 * eps * (-(torch.exp(fv2) * grad1) + tv2)) * torch.exp(sv2)
        m = self._m[int(step)]
        mb = 1.0 - m
        X1 = self._XNet(v_h, mb * x_o, t, ae_x)
        sx2 = -eps * X1[0]
        tx2 = X1[1]
        fx2 = eps * X1[2]
        y = mb * x_o + m * (torch.exp(sx2) * (x_o - eps * (torch.exp(fx2) * v_h + tx2)))
        X2 = self._XNet(v_h, m * y, t, ae_x)
        sx1 = -eps * X2[0]
        tx1 = X2[1]
        fx1 = eps * X2[2]
        x = m * y + mb * (torch.exp(sx1) * (y - eps * (torch.exp(fx1) * v_h + tx1)))
        grad2 = self.grad_energy(x)
        S2 = self._VNet(x, grad2, t, ae_v)
        sv1 = -0.5 * eps * S2[0]
        tv1 = S2[1]
        fv1 = eps * S2[2]
        v = torch.exp(sv1) * (v_h - 0.5 * eps * (-(torch.exp(fv1) * grad2) + tv1))
        return x, v, (sv1 + sv2 + mb * sx1 + m * sx2).sum(1)

    # -- T-step loops -------------------------------------------------------------------
    def forward(self, x, v, log_jac=False, ae_x=None, ae_v=None):
        """forward (utils/dynamics.py:246-272) with v injected (the reference draws it, :248)."""
        x = x.to(self.dtype)
        v = v.to(self.dtype)
        X, V = x, v
        j = torch.zeros(x.shape[0], dtype=self.dtype)
        for t in range(self.T):
            X, V, lj = self.forward_step(X, V, t, ae_x, ae_v)
            j = j + lj
        if log_jac:
            return X, V, j
        return X, V, self.p_accept(x, v, X, V, j)

    def backward(self, x, v, log_jac=False, ae_x=None, ae_v=None):
        """backward (utils/dynamics.py:274-300): steps T-1 ... 0 (:285)."""
        x = x.to(self.dtype)
        v = v.to(self.dtype)
        X, V = x, v
        j = torch.zeros(x.shape[0], dtype=self.dtype)
        for t in range(self.T):
            X, V, lj = self.backward_step(X, V, self.T - t - 1, ae_x, ae_v)
            j = j + lj
        if log_jac:
            return X, V, j
        return X, V, self.p_accept(x, v, X, V, j)

    def p_accept(self, x0, v0, x1, v1, log_jac):
        """p_accept (utils/dynamics.py:302-309)."""
        e_new = self.hamiltonian(x1, v1)
        e_old = self.hamiltonian(x0, v0)
        v = e_old - e_new + log_jac
        p = torch.exp(torch.minimum(v, torch.zeros_like(v)))
        return torch.where(torch.isfinite(p), p, torch.zeros_like(p))


# --------------------------------------------------------------------------------------
# Sampler (utils/sampler.py)
# --------------------------------------------------------------------------------------
def tf_accept(x, Lx, px, u):
    """tf_accept (utils/sampler.py:53-55) with the uniforms injected."""
    mask = (px - u) >= 0.0
    return torch.where(mask[:, None], Lx, x)


def propose(x, dyn: OracleDynamics, *, direction=None, v_f=None, v_b=None, u=None,
            init_v=None, do_mh_step=False, log_jac=False, ae_x=None, ae_v=None):
    """propose (utils/sampler.py:28-51) with injected randomness.

    Non-HMC: BOTH directions are run for every chain (each with its own fresh v) and blended by the
    direction bit, exactly as the reference does (:34-44).  HMC: forward only with init_v (:29-31).
    Returns (Lx, Lv, px, outputs) like the reference.
    """
    x = x.to(dyn.dtype)
    if dyn.hmc:
        v0 = init_v if init_v is not None else v_f
        Lx, Lv, px = dyn.forward(x, v0.to(dyn.dtype), ae_x=ae_x, ae_v=ae_v)
        return Lx, Lv, px, [tf_accept(x, Lx, px, u.to(dyn.dtype))]
    mask = direction.to(dyn.dtype)[:, None]
    Lx1, Lv1, px1 = dyn.forward(x, v_f, log_jac=log_jac, ae_x=ae_x, ae_v=ae_v)
    Lx2, Lv2, px2 = dyn.backward(x, v_b, log_jac=log_jac, ae_x=ae_x, ae_v=ae_v)
    Lx = mask * Lx1 + (1 - mask) * Lx2
    Lv = None
    if init_v is not None:
        Lv = mask * Lv1 + (1 - mask) * Lv2
    px = mask[:, 0] * px1 + (1 - mask)[:, 0] * px2
    outputs = []
    if do_mh_step:
        outputs.append(tf_accept(x, Lx, px, u.to(dyn.dtype)))
    return Lx, Lv, px, outputs


def propose_selected(x, dyn: OracleDynamics, *, direction, v, log_jac=False, ae_x=None, ae_v=None):
    """Same transition, but each chain runs only its selected direction with the single v it is given
    (what a fused kernel does).  Equal to propose(...) with v_f = v_b = v wherever the unselected
    direction stays finite.  Returns (Lx, Lv, px).  Energies that hold per-chain rows (aux) provide
    ``select(idx)``; ae_x / ae_v are the per-chain aux embeddings of the two nets."""
    import copy
    x = x.to(dyn.dtype)
    d = direction.to(torch.bool)
    Lx = torch.empty_like(x)
    Lv = torch.empty_like(x)
    px = torch.empty(x.shape[0], dtype=dyn.dtype)
    for sel, fn in ((d, "forward"), (~d, "backward")):
        if not sel.any():
            continue
        sub = dyn
        if hasattr(dyn.energy_obj, "select"):
            sub = copy.copy(dyn)
            sub.energy_obj = dyn.energy_obj.select(sel)
        a, b, c = getattr(sub, fn)(x[sel], v[sel].to(dyn.dtype), log_jac=log_jac,
                                   ae_x=None if ae_x is None else ae_x[sel],
                                   ae_v=None if ae_v is None else ae_v[sel])
        Lx[sel], Lv[sel], px[sel] = a, b, c
    return Lx, Lv, px


def chain_operator(init_x, dyn: OracleDynamics, nb_steps: int, *, init_v, directions, v_fs, v_bs, u=None,
                   do_mh_step=False):
    """chain_operator (utils/sampler.py:57-85): nb_steps proposals composed with accumulated log|J|, one
    MH at the end.  Note the reference's quirk: sub-proposals ignore the carried v (fresh v per
    direction, utils/sampler.py:35-36) but the final p_accept uses init_v and the last blended Lv."""
    x = init_x.to(dyn.dtype)
    v = init_v.to(dyn.dtype)
    lj = torch.zeros(x.shape[0], dtype=dyn.dtype)
    cx, cv = x, v
    for s in range(nb_steps):
        cx, cv, px, _ = propose(cx, dyn, direction=directions[s], v_f=v_fs[s].to(dyn.dtype),
                                v_b=v_bs[s].to(dyn.dtype), init_v=cv, log_jac=True)
        lj = lj + px
    p = dyn.p_accept(x, v, cx, cv, lj)
    outputs = []
    if do_mh_step:
        outputs.append(tf_accept(x, cx, p, u.to(dyn.dtype)))
    return cx, cv, p, outputs


# --------------------------------------------------------------------------------------
# Synthetic problem builders shared by tests / bench
# --------------------------------------------------------------------------------------
# (defined with the product's synthetic workloads so that bench.py's product arm needs no test infrastructure; the
# oracle sees the same arrays)
from l2hmc_b200.synthetic import trunc_normal, make_net, make_masks  # noqa: E402,F401


# --------------------------------------------------------------------------------------
# Diagnostics (utils/func_utils.py:45-54,114-120), numpy like the reference
# --------------------------------------------------------------------------------------
def autocovariance(X, tau=0):
    """utils/func_utils.py:45-54: mean over t of sum(X[t] * X[t+tau]) / N."""
    dT, dN, dX = np.shape(X)
    s = 0.
    for t in range(dT - tau):
        s += np.sum(X[t, :, :] * X[t + tau, :, :]) / dN
    return s / (dT - tau)


def acl_spectrum(X, scale):
    """utils/func_utils.py:114-116.  The samples are float32 (sess.run output) and the scale a float64 scalar; under the
    numpy of the reference's time a scalar does not upcast an array, so X / scale -- and every product and per-step sum
    in autocovariance -- is float32 (only the running total over t is a python float).  Written out explicitly here
    because NumPy >= 2 would promote to float64."""
    n = X.shape[0]
    Xs = (np.asarray(X, dtype=np.float32) / np.float32(scale)).astype(np.float32)
    return np.array([autocovariance(Xs, tau=t) for t in range(n - 1)])


def ESS(A):
    """utils/func_utils.py:118-120."""
    A = A * (A > 0.05)
    return 1. / (1. + 2 * np.sum(A[1:]))


# --------------------------------------------------------------------------------------
# Annealed importance sampling (utils/ais.py:30-82), HMC-mode Dynamics inside a scan over beta
# --------------------------------------------------------------------------------------
class MixedEnergy(Energy):
    """curr_energy of utils/ais.py:44-45: (1 - beta) * init_energy(z) + beta * final_energy(z)."""

    def __init__(self, e0: Energy, e1: Energy, beta: float):
        self.e0, self.e1, self.beta = e0, e1, float(beta)

    def to(self, dtype):
        return MixedEnergy(self.e0.to(dtype), self.e1.to(dtype), self.beta)

    def energy(self, x):
        return (1.0 - self.beta) * self.e0.energy(x) + self.beta * self.e1.energy(x)

    def grad(self, x):
        return (1.0 - self.beta) * self.e0.grad(x) + self.beta * self.e1.grad(x)


def ais_estimate(e0: Energy, e1: Energy, anneal_steps: int, initial_x, *, step_size, leapfrogs, v0, v_refresh, u,
                 num_splits=1, refresh=False, refreshment=0.1, dtype=torch.float64):
    """ais_estimate (utils/ais.py:30-82) with all randomness injected: v0 [N, D] (the scan's initial momentum, used
    only when refresh=True), v_refresh [steps, N, D] normals, u [steps, N] uniforms.
    Returns (log-mean-exp of the weights per split summed, mean accept probability, final x, final w)."""
    x = torch.as_tensor(np.asarray(initial_x)).to(dtype)
    n, D = x.shape
    w = torch.zeros((n,), dtype=dtype)
    v = torch.as_tensor(np.asarray(v0)).to(dtype)
    e0, e1 = e0.to(dtype), e1.to(dtype)
    # beta = linspace(0, 1, steps + 1)[1:], beta_diff = beta[1] - beta[0] in fp32 like tf.linspace (utils/ais.py:43-44)
    beta = np.linspace(0.0, 1.0, anneal_steps + 1, dtype=np.float32)[1:]
    beta_diff = float(beta[1] - beta[0]) if anneal_steps > 1 else float(beta[0])
    alphas = []
    for s in range(anneal_steps):
        z = torch.as_tensor(np.asarray(v_refresh[s])).to(dtype)
        rv = v * math.sqrt(1.0 - refreshment) + z * math.sqrt(refreshment) if refresh else z
        w = w + beta_diff * (-e1.energy(x) + e0.energy(x))
        dyn = OracleDynamics(D, leapfrogs, step_size, MixedEnergy(e0, e1, float(beta[s])), np.zeros((leapfrogs, D), np.float32),
                             hmc=True, dtype=dtype)
        Lx, Lv, px = dyn.forward(x, rv)
        mask = (px - torch.as_tensor(np.asarray(u[s])).to(dtype)) >= 0.0
        x = torch.where(mask[:, None], Lx, x)
        v = torch.where(mask[:, None], Lv, -Lv)
        alphas.append(px)
    lme = lambda zz: torch.logsumexp(zz, 0) - math.log(zz.shape[0])  # noqa: E731
    parts = torch.chunk(w, num_splits, dim=0)
    est = sum(lme(p) for p in parts)
    return est, torch.stack(alphas).mean(), x, w


# --------------------------------------------------------------------------------------
# Training objective (utils/losses.py:36-59, SCGExperiment.ipynb:159-181) -- oracle groundwork for SURVEY
# section 8(f)3: the losses as the reference writes them, differentiated by torch autograd through this
# restatement of the dynamics (the reference back-propagates through the unrolled tf.while_loop, including the
# tf.gradients of the energy, i.e. second-order terms; grad() of every Energy here is a differentiable expression).
# --------------------------------------------------------------------------------------
def c32(value):
    """A python float meeting an fp32 tensor becomes an fp32 constant in the reference's graph; the fp64 twin keeps that
    rounded value (the fp32 graph in wider arithmetic -- what the reference run on oracle/tf_shim computes)."""
    return float(np.float32(value))


EPS_V = c32(1e-4)


def loss_vec(x, X, p):
    """loss_vec (utils/losses.py:36-37): expected squared jump distance per chain, + 1e-4."""
    return ((X - x) ** 2).sum(1) * p + EPS_V


def loss_logsumexp(x, X, p):
    v = loss_vec(x, X, p)
    return torch.logsumexp(-v, 0) - math.log(v.shape[0])


def loss_inverse(x, X, p):
    v = loss_vec(x, X, p)
    return -1.0 / (1.0 / (v + EPS_V)).mean()


def loss_std(x, X, p):
    return -loss_vec(x, X, p).mean(0)


def loss_mixed(x, Lx, px, scale=1.0):
    """loss_mixed (utils/losses.py:53-59)."""
    v1 = loss_vec(x, Lx, px) / c32(scale)
    return (1.0 / v1).mean() - v1.mean()


def notebook_loss(x, z, dyn: OracleDynamics, rx: dict, rz: dict, scale=0.1):
    """The SCG notebook's objective (SCGExperiment.ipynb:159-181): proposals from data samples x and from noise z,
    loss = scale (E[1/v1] + E[1/v2]) - (E[v1] + E[v2]) / scale with v = |x - Lx|^2 p + 1e-4.
    rx / rz: injected randomness of the two propose calls {'direction', 'v_f', 'v_b'}."""
    x = x.to(dyn.dtype)
    z = z.to(dyn.dtype)
    Lx, _, px, _ = propose(x, dyn, direction=rx["direction"], v_f=rx["v_f"].to(dyn.dtype), v_b=rx["v_b"].to(dyn.dtype))
    Lz, _, pz, _ = propose(z, dyn, direction=rz["direction"], v_f=rz["v_f"].to(dyn.dtype), v_b=rz["v_b"].to(dyn.dtype))
    scale = c32(scale)
    v1 = ((x - Lx) ** 2).sum(1) * px + EPS_V
    v2 = ((z - Lz) ** 2).sum(1) * pz + EPS_V
    return scale * ((1.0 / v1).mean() + (1.0 / v2).mean()) + (-v1.mean() - v2.mean()) / scale


def trainable_parameters(dyn: OracleDynamics):
    """The tensors the reference trains: both nets' weights / biases / scales (utils/layers.py:31-34,83-84); alpha =
    log(eps) is trainable in the reference too (utils/dynamics.py:50-54) but is a constant of this restatement."""
    ps = []
    for net in (dyn.xnet, dyn.vnet):
        for k in sorted(net):
            net[k].requires_grad_(True)
            ps.append(net[k])
    return ps
