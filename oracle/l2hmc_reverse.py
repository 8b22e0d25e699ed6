"""Hand-derived reverse mode through the L2HMC transition and the notebook training objective.

TEST INFRASTRUCTURE, like everything under oracle/: imported only by tests/ (pinned: tests/test_reference_pin.py checks it against
tf.gradients of the reference's own loss cell run on oracle/tf_shim).  This file is the algorithm statement for SURVEY section 8(f)3, the training path:
the reference obtains its parameter gradients from TF1 autodiff through the unrolled ``tf.while_loop`` of
``Dynamics.forward / backward`` (utils/dynamics.py:246-300), through ``tf.gradients`` of the energy inside it (:217-218,
i.e. Hessian-vector products on the way back), ``p_accept`` (:302-309), ``propose`` (utils/sampler.py:28-51) and the
loss (utils/losses.py:36-59, SCGExperiment.ipynb:159-181).  Here the same gradient is written out by hand, sub-update
by sub-update, with NO autograd call (everything runs under ``torch.no_grad``) -- the form a CUDA backward kernel can
follow -- and tests/test_oracle.py checks it against torch.autograd applied to l2hmc_oracle.py.

Structure.  A leapfrog step is four sub-updates (utils/dynamics.py:115-157 forward, :159-201 backward):

    forward  step t:  V+(x, v) ; X+(keep = m_t) ; X+(keep = 1 - m_t) ; V+
    backward step t:  V-(x, v) ; X-(keep = 1 - m_t) ; X-(keep = m_t) ; V-

``V`` reads (x, grad U(x), t) through VNet and rescales / shifts v; ``X`` reads (v, keep * x, t) through XNet and
rescales / shifts the dimensions of x that are not kept.  The forward sweep records the state in front of every
sub-update; the reverse sweep walks the records backwards, recomputes the sub-update's intermediates from its record
(one net forward) and applies the vector-Jacobian product below.  Cotangents carried between sub-updates: gx, gv [N, D]
and glj [N] (the cotangent of the accumulated log|J|, the same for every step); accumulated on the side: the two
nets' parameter gradients and d/d(eps).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

import l2hmc_oracle as O


# --------------------------------------------------------------------------------------
# S/T/Q net: forward with saved activations, and its vector-Jacobian product
# (SCGExperiment.ipynb:51-77; utils/layers.py:29-37,81-86)
# --------------------------------------------------------------------------------------
def net_forward_saved(p, a, b, tau):
    z1 = (((0 + (a @ p["W1"] + p["b1"])) + (b @ p["W2"] + p["b2"])) + (tau @ p["W3"] + p["b3"]))
    h1 = torch.relu(z1)
    z2 = h1 @ p["W4"] + p["b4"]
    h2 = torch.relu(z2)
    ts = torch.tanh(h2 @ p["Ws"] + p["bs"])
    tq = torch.tanh(h2 @ p["Wq"] + p["bq"])
    es = torch.exp(p["ls"])
    eq = torch.exp(p["lq"])
    S = es * ts
    T = h2 @ p["Wt"] + p["bt"]
    Q = eq * tq
    return (S, T, Q), (a, b, tau, z1, h1, z2, h2, ts, tq, es, eq)


def net_vjp(p, saved, gS, gT, gQ, acc: Dict[str, torch.Tensor]):
    """Cotangents (gS, gT, gQ) [N, D] -> (ga, gb) [N, D]; parameter gradients are added into ``acc``."""
    a, b, tau, z1, h1, z2, h2, ts, tq, es, eq = saved
    g_us = gS * es * (1.0 - ts * ts)
    g_uq = gQ * eq * (1.0 - tq * tq)
    acc["ls"] += (gS * es * ts).sum(0).reshape(p["ls"].shape)   # d(e^l tanh u)/dl = e^l tanh u
    acc["lq"] += (gQ * eq * tq).sum(0).reshape(p["lq"].shape)
    acc["Ws"] += h2.T @ g_us
    acc["bs"] += g_us.sum(0).reshape(p["bs"].shape)
    acc["Wt"] += h2.T @ gT
    acc["bt"] += gT.sum(0).reshape(p["bt"].shape)
    acc["Wq"] += h2.T @ g_uq
    acc["bq"] += g_uq.sum(0).reshape(p["bq"].shape)
    g_h2 = g_us @ p["Ws"].T + gT @ p["Wt"].T + g_uq @ p["Wq"].T
    g_z2 = g_h2 * (z2 > 0).to(g_h2.dtype)
    acc["W4"] += h1.T @ g_z2
    acc["b4"] += g_z2.sum(0).reshape(p["b4"].shape)
    g_z1 = (g_z2 @ p["W4"].T) * (z1 > 0).to(g_h2.dtype)
    s1 = g_z1.sum(0)
    acc["W1"] += a.T @ g_z1
    acc["W2"] += b.T @ g_z1
    acc["W3"] += tau.T @ g_z1
    for k in ("b1", "b2", "b3"):
        acc[k] += s1.reshape(p[k].shape)
    return g_z1 @ p["W1"].T, g_z1 @ p["W2"].T


# --------------------------------------------------------------------------------------
# Hessian-vector products of the energies (what back-propagating through tf.gradients(energy, x) costs)
# --------------------------------------------------------------------------------------
def energy_hvp(e: O.Energy, x, w):
    """w -> w . d(grad U)/dx, closed form per energy kind (the Hessian is symmetric)."""
    if isinstance(e, O.GaussianEnergy):
        return 0.5 * (w @ e.S) + 0.5 * (w @ e.S.T)          # utils/distributions.py:50-57
    if isinstance(e, O.RoughWellEnergy):
        eps, den = e._scale(x)
        return w * (1.0 - eps * torch.cos(x / den) / (den * den))   # :90-97, diagonal Hessian
    if isinstance(e, O.GMMEnergy):
        # U = -logsumexp_i a_i, grad U = sum_i r_i g_i with r = softmax(a), g_i = A_i (x - mu_i), A_i = (S_i+S_i^T)/2
        # Hessian = sum_i r_i A_i - sum_i r_i g_i g_i^T + gbar gbar^T, gbar = grad U          (:125-134)
        r = torch.softmax(e._V(x), dim=1)
        out = torch.zeros_like(x)
        gbar = torch.zeros_like(x)
        for i, (m, S) in enumerate(zip(e.mus, e.Ss)):
            d = x - m
            gi = 0.5 * (d @ S.T) + 0.5 * (d @ S)
            ri = r[:, i:i + 1]
            out = out + ri * (0.5 * (w @ S) + 0.5 * (w @ S.T)) - ri * gi * (gi * w).sum(1, keepdim=True)
            gbar = gbar + ri * gi
        return out + gbar * (gbar * w).sum(1, keepdim=True)
    if isinstance(e, O.FunnelEnergy):
        # grad = (v / sigma^2 + (n - |y|^2 / s) / 2, y / s), v = x_0, y = x_1:, s = e^v inside the clip; outside it s is a
        # constant and the coupling to v drops out (utils/distributions.py:161-180)
        v, y = x[:, 0], x[:, 1:]
        out = (v > e.clip) | (-e.clip > v)
        inside = (~out).to(x.dtype)
        s = torch.where(out, torch.exp(torch.where(v > e.clip, torch.full_like(v, e.clip), torch.full_like(v, -e.clip))),
                        torch.exp(v))
        ss, wy = (y * y).sum(1), (w[:, 1:] * y).sum(1)
        h = torch.empty_like(x)
        h[:, 0] = w[:, 0] * (1.0 / e.sigma ** 2 + inside * 0.5 * ss / s) - inside * wy / s
        h[:, 1:] = (w[:, 1:] - (inside * w[:, 0])[:, None] * y) / s[:, None]
        return h
    # the decoder energy: differentiate its closed-form gradient expression instead of restating the Hessian
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        (h,) = torch.autograd.grad((e.grad(xr) * w).sum(), xr)
    return h


# --------------------------------------------------------------------------------------
# The two sub-updates, forward (with what the reverse needs) and reverse
# --------------------------------------------------------------------------------------
class _Acc:
    """Side accumulators of one reverse sweep."""

    def __init__(self, dyn: O.OracleDynamics):
        self.x = {k: torch.zeros_like(v) for k, v in dyn.xnet.items()}
        self.v = {k: torch.zeros_like(v) for k, v in dyn.vnet.items()}
        self.eps = torch.zeros((), dtype=dyn.dtype)


def _temp(dyn):
    return torch.tensor(np.float32(dyn.temperature)).to(dyn.dtype)


def v_update(dyn, x, v, tau, sign: int):
    """sign=+1: utils/dynamics.py:117-128 and :146-153; sign=-1: :161-170 and :190-199."""
    e = dyn._eps
    g = dyn.grad_energy(x)
    (S, T, Q), saved = net_forward_saved(dyn.vnet, x, g, tau)
    s = (0.5 * sign) * e * S
    f = e * Q
    shift = 0.5 * e * (-(torch.exp(f) * g) + T)
    v_o = v * torch.exp(s) + shift if sign > 0 else (v - shift) * torch.exp(s)
    return v_o, s.sum(1), (g, S, T, Q, s, f, saved)


def v_update_vjp(dyn, x, v, tau, sign: int, gx, gv_o, glj, acc: _Acc):
    e = dyn._eps
    v_o, _, (g, S, T, Q, s, f, saved) = v_update(dyn, x, v, tau, sign)
    es, ef = torch.exp(s), torch.exp(f)
    inner = -(ef * g) + T                       # shift = 0.5 e inner
    if sign > 0:
        gv = gv_o * es
        g_s = gv_o * v * es + glj[:, None]
        g_shift = gv_o
    else:
        gv = gv_o * es
        g_s = gv_o * v_o + glj[:, None]        # d((v - shift) e^s)/ds = v_o
        g_shift = -gv_o * es
    g_inner = g_shift * (0.5 * e)
    acc.eps += (g_shift * 0.5 * inner).sum()
    g_f = g_inner * (-(ef * g))
    g_g = g_inner * (-ef)
    gT = g_inner
    gS = g_s * ((0.5 * sign) * e)
    acc.eps += (g_s * (0.5 * sign) * S).sum()
    gQ = g_f * e
    acc.eps += (g_f * Q).sum()
    ga, gb = net_vjp(dyn.vnet, saved, gS, gT, gQ, acc.v)
    g_g = g_g + gb
    gx = gx + ga + energy_hvp(dyn.energy_obj, x, g_g) / _temp(dyn)
    return gx, gv


def x_update(dyn, x, v, tau, keep, sign: int):
    """sign=+1: utils/dynamics.py:131-144; sign=-1: :173-188.  ``keep`` [D] are the dimensions left unchanged."""
    e = dyn._eps
    upd = 1.0 - keep
    (S, T, Q), saved = net_forward_saved(dyn.xnet, v, keep * x, tau)
    s = sign * e * S
    f = e * Q
    shift = e * (torch.exp(f) * v + T)
    if sign > 0:
        x_o = keep * x + upd * (x * torch.exp(s) + shift)
    else:
        x_o = keep * x + upd * (torch.exp(s) * (x - shift))
    return x_o, (upd * s).sum(1), (S, T, Q, s, f, saved)


def x_update_vjp(dyn, x, v, tau, keep, sign: int, gx_o, gv, glj, acc: _Acc):
    e = dyn._eps
    upd = 1.0 - keep
    x_o, _, (S, T, Q, s, f, saved) = x_update(dyn, x, v, tau, keep, sign)
    es, ef = torch.exp(s), torch.exp(f)
    inner = ef * v + T                          # shift = e inner
    gu = gx_o * upd
    if sign > 0:
        gx = gx_o * keep + gu * es
        g_s = gu * x * es + glj[:, None] * upd
        g_shift = gu
    else:
        gx = gx_o * keep + gu * es
        g_s = gu * es * (x - e * inner) + glj[:, None] * upd
        g_shift = -gu * es
    g_inner = g_shift * e
    acc.eps += (g_shift * inner).sum()
    g_f = g_inner * ef * v
    gv = gv + g_inner * ef
    gT = g_inner
    gS = g_s * (sign * e)
    acc.eps += (g_s * sign * S).sum()
    gQ = g_f * e
    acc.eps += (g_f * Q).sum()
    ga, gb = net_vjp(dyn.xnet, saved, gS, gT, gQ, acc.x)
    return gx + keep * gb, gv + ga


def _schedule(dyn, forward: bool) -> List[Tuple[str, int, int, torch.Tensor]]:
    """The 4 T sub-updates of Dynamics.forward (utils/dynamics.py:260-267) or .backward (:285-293):
    (kind, step, sign, keep-mask)."""
    out = []
    for i in range(dyn.T):
        t = i if forward else dyn.T - i - 1
        m = dyn._m[t]
        mb = 1.0 - m
        if forward:
            out += [("v", t, +1, None), ("x", t, +1, m), ("x", t, +1, mb), ("v", t, +1, None)]
        else:
            out += [("v", t, -1, None), ("x", t, -1, mb), ("x", t, -1, m), ("v", t, -1, None)]
    return out


def transition(dyn, x0, v0, forward: bool):
    """Forward sweep of one direction, recording the state in front of each sub-update."""
    x, v = x0, v0
    lj = torch.zeros(x.shape[0], dtype=dyn.dtype)
    tape = []
    for kind, t, sign, keep in _schedule(dyn, forward):
        tau = dyn.format_time(t, x.shape[0])
        tape.append((x, v))
        if kind == "v":
            v, dl, _ = v_update(dyn, x, v, tau, sign)
        else:
            x, dl, _ = x_update(dyn, x, v, tau, keep, sign)
        lj = lj + dl
    return x, v, lj, tape


def transition_vjp(dyn, forward: bool, tape, gX, gV, glj, acc: _Acc):
    """Reverse sweep: cotangents of (X, V, log|J|) -> cotangents of (x0, v0)."""
    gx, gv = gX, gV
    for (kind, t, sign, keep), (x, v) in zip(reversed(_schedule(dyn, forward)), reversed(tape)):
        tau = dyn.format_time(t, x.shape[0])
        if kind == "v":
            gx, gv = v_update_vjp(dyn, x, v, tau, sign, gx, gv, glj, acc)
        else:
            gx, gv = x_update_vjp(dyn, x, v, tau, keep, sign, gx, gv, glj, acc)
    return gx, gv


# --------------------------------------------------------------------------------------
# p_accept, propose and the objective
# --------------------------------------------------------------------------------------
def accept_prob_vjp(dyn, x0, v0, X, V, lj, gp):
    """p = exp(min(H(x0,v0) - H(X,V) + log|J|, 0)), non-finite -> 0 (utils/dynamics.py:302-309).
    Returns p and the cotangents of (X, V, log|J|); x0 and v0 are data / noise here and get none."""
    arg = dyn.hamiltonian(x0, v0) - dyn.hamiltonian(X, V) + lj
    p = torch.exp(torch.minimum(arg, torch.zeros_like(arg)))
    ok = torch.isfinite(p)
    p = torch.where(ok, p, torch.zeros_like(p))
    g_arg = torch.where(ok & (arg < 0), gp * p, torch.zeros_like(p))
    gX = -g_arg[:, None] * dyn.grad_energy(X)
    gV = -g_arg[:, None] * V
    return p, gX, gV, g_arg


LOSS_KINDS = ("mixed", "standard", "inverse", "logsumexp")   # get_loss names, utils/losses.py:26-34


def loss_value_and_dv(v, kind: str, scale: float, count: float):
    """Loss of utils/losses.py:36-59 as a function of v = loss_vec(x, Lx, px) [n], and d loss / d v, by hand.
    ``count``: the number of chains the means run over (the batch, unless the caller splits one)."""
    scale = O.c32(scale)
    if kind == "mixed":      # scale mean(1/v) - mean(v)/scale   (:53-59; SCGExperiment.ipynb:171-181)
        return (scale / v).sum() / count - v.sum() / (count * scale), (-scale / (v * v) - 1.0 / scale) / count
    if kind == "standard":   # -mean(v)   (:49-51)
        return -v.sum() / count, torch.full_like(v, -1.0 / count)
    if kind == "inverse":    # -1 / mean(1 / (v + O.EPS_V))   (:44-47)
        m = (1.0 / (v + O.EPS_V)).sum() / count
        return -1.0 / m, -1.0 / (m * m * (v + O.EPS_V) ** 2 * count)
    if kind == "logsumexp":  # logsumexp(-v) - log n   (:39-42)
        mx = (-v).max()
        z = torch.exp(-v - mx).sum()
        return mx + torch.log(z) - math.log(count), -torch.exp(-v - mx) / z
    raise ValueError(kind)


def loss_and_grads(x, dyn: O.OracleDynamics, r: dict, scale: float, acc: _Acc, kind: str = "mixed", count=None):
    """One ``propose`` batch of a utils/losses.py objective on v = |x - Lx|^2 p + 1e-4 (kind 'mixed' with the scale is the
    notebook's, SCGExperiment.ipynb:171-181).  Each chain back-propagates through its selected direction only: the other
    one is multiplied by a zero mask (utils/sampler.py:38,44)."""
    x = x.to(dyn.dtype)
    n = x.shape[0]
    count = float(n if count is None else count)
    d = r["direction"].to(torch.bool)
    groups = []
    v_all = torch.zeros(n, dtype=dyn.dtype)
    zero = None
    for sel, fwd, vkey in ((d, True, "v_f"), (~d, False, "v_b")):
        if not sel.any():
            continue
        xs, vs = x[sel], r[vkey].to(dyn.dtype)[sel]
        X, V, lj, tape = transition(dyn, xs, vs, fwd)
        zero = torch.zeros(xs.shape[0], dtype=dyn.dtype)
        p, _, _, _ = accept_prob_vjp(dyn, xs, vs, X, V, lj, zero)
        sq = ((xs - X) ** 2).sum(1)
        v_all[sel] = sq * p + O.EPS_V
        groups.append((sel, fwd, xs, vs, X, V, lj, tape, p, sq))
    total, g_v_all = loss_value_and_dv(v_all, kind, scale, count)
    for sel, fwd, xs, vs, X, V, lj, tape, p, sq in groups:
        g_v = g_v_all[sel]
        gX = (g_v * p)[:, None] * 2.0 * (X - xs)
        _, gX2, gV, glj = accept_prob_vjp(dyn, xs, vs, X, V, lj, g_v * sq)
        transition_vjp(dyn, fwd, tape, gX + gX2, gV, glj, acc)
    return total


def notebook_loss_and_grads(x, z, dyn: O.OracleDynamics, rx: dict, rz: dict, scale=0.1):
    """Value and gradient of l2hmc_oracle.notebook_loss without autograd.
    Returns (loss, {'xnet': {...}, 'vnet': {...}, 'eps': d/d eps, 'alpha': d/d log eps})."""
    with torch.no_grad():
        acc = _Acc(dyn)
        loss = loss_and_grads(x, dyn, rx, scale, acc) + loss_and_grads(z, dyn, rz, scale, acc)
        return loss, {"xnet": acc.x, "vnet": acc.v, "eps": acc.eps, "alpha": acc.eps * dyn._eps}
