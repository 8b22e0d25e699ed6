"""Shared builders for the tests: synthetic problems expressed twice --
as an oracle ``OracleDynamics`` and as a product ``l2hmc_b200.Dynamics`` -- from the same arrays."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import l2hmc_oracle as O  # noqa: E402  (test infrastructure)


from l2hmc_b200 import synthetic as S  # noqa: E402  (product-side problem definitions)
from l2hmc_b200.synthetic import scg2_cov, scg_cov  # noqa: E402,F401


def oracle_energy(P):
    """The oracle's Energy for a synthetic problem: built from the SAME numbers the product's distribution holds."""
    g = P.dist
    if P.kind == "gaussian":
        return O.GaussianEnergy(np.asarray(g.mu, dtype=np.float32), g.i_sigma.astype(np.float32))
    if P.kind == "gmm":
        return O.GMMEnergy(g.mus, g.i_sigmas, g.constants)
    if P.kind == "roughwell":
        return O.RoughWellEnergy(g.eps, g.easy)
    if P.kind == "funnel":
        return O.FunnelEnergy(g.sigma, g.clip)
    raise ValueError(P.kind)


def target(kind, D, **kw):
    """Returns (product distribution object, oracle Energy, x0 sampler(n, rng))."""
    g, x0 = S.target(kind, D, **kw)
    P = type("T", (), {"kind": kind, "dist": g})
    return g, oracle_energy(P), x0


# ---- problems ---------------------------------------------------------------------------------------
class Problem(S.SyntheticProblem):
    """One synthetic configuration (l2hmc_b200/synthetic.py: weights, masks, target, product object) plus the oracle's
    view of the same arrays."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.energy = oracle_energy(self)

    def oracle(self, dtype=torch.float64, temperature=1.0):
        return O.OracleDynamics(self.D, self.T, self.eps, self.energy, self.mask, self.xnet, self.vnet,
                                hmc=self.hmc, temperature=temperature, dtype=dtype)


class VaeProblem(S.SyntheticVaeProblem):
    """BASELINE config 5 (l2hmc_b200/synthetic.py) plus the oracle's view of the same arrays."""

    def oracle_for(self, d, dtype=torch.float64, temperature=1.0):
        en = O.DecoderBernoulliEnergy(self.dec_W, self.dec_b, d["aux"], dtype)
        dyn = O.OracleDynamics(self.D, self.T, self.eps, en, self.mask, self.xnet, self.vnet, hmc=False,
                               temperature=temperature, dtype=dtype)
        ae = None
        if self.use_encoder:
            tt = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)  # noqa: E731
            ae = O.softplus_mlp([tt(W) for W in self.enc_W], [tt(b) for b in self.enc_b], tt(d["aux"]))
        return dyn, ae


def vae_weight_checksum(P):
    """The reference-pinned VAE fixture at mnist_vae.py's layer sizes (tests/golden/ref/c5_vae_full_n32.npz) does not
    store its 2.9 M weights: VaeProblem regenerates them from its seed; this checksum proves the regenerated weights
    are the ones the fixture was made with."""
    tot = 0.0
    for a in list(P.dec_W) + list(P.dec_b) + (list(P.enc_W) + list(P.enc_b) if P.use_encoder else []) + \
            [P.xnet[k] for k in sorted(P.xnet)] + [P.vnet[k] for k in sorted(P.vnet)]:
        a = np.asarray(a, dtype=np.float64)
        tot += float(np.abs(a).sum()) + 3.0 * float(a.reshape(-1)[:: max(1, a.size // 7)].sum())
    return tot


VAE_CONFIGS = S.VAE_CONFIGS


def t64(a):
    return torch.as_tensor(np.asarray(a)).to(torch.float64)


def max_rel(a, b):
    """max |a-b| / max(1, max|b|): the north star's 'relative fp32 tolerance' on a whole array."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


CONFIGS = S.CONFIGS  # BASELINE.json configs, reduced where noted by the tests


# ---- parity measurement (GPU) ------------------------------------------------------------------------
def run_oracle_propose(P, d, dtype, log_jac=False):
    """Reference-style propose (both directions computed, blended) on the CPU oracle."""
    dyn, ae = P.oracle_for(d, dtype) if hasattr(P, "oracle_for") else (P.oracle(dtype), None)
    tt = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    if P.hmc:
        Lx, Lv, px, outs = O.propose(tt(d["x"]), dyn, init_v=tt(d["v_f"]), u=tt(d["u"]), do_mh_step=True)
    else:
        Lx, Lv, px, outs = O.propose(tt(d["x"]), dyn, direction=torch.as_tensor(d["dir"].astype(np.float32)),
                                     v_f=tt(d["v_f"]), v_b=tt(d["v_b"]), u=tt(d["u"]), init_v=tt(d["v_f"]),
                                     do_mh_step=True, log_jac=log_jac, ae_x=ae, ae_v=ae)
    return {"Lx": Lx.numpy(), "Lv": Lv.numpy(), "px": px.numpy(), "x_next": outs[0].numpy()}


def run_kernel_propose(P, d, dyn=None, log_jac=False, device="cuda"):
    """Same transition through the product API (injected randomness)."""
    from l2hmc_b200 import propose
    dyn = dyn or P.product()
    g = lambda a: torch.as_tensor(np.asarray(a)).to(device)
    x = g(d["x"])
    aux = g(d["aux"]) if "aux" in d else None
    if P.hmc:
        Lx, Lv, px, outs = propose(x, dyn, init_v=g(d["v_f"]), aux=aux, do_mh_step=True, rng={"u": g(d["u"])})
    else:
        v_sel = np.where(d["dir"][:, None] != 0, d["v_f"], d["v_b"]).astype(np.float32)
        Lx, Lv, px, outs = propose(x, dyn, init_v=g(v_sel), aux=aux, do_mh_step=True, log_jac=log_jac,
                                   rng={"direction": g(d["dir"]), "v": g(v_sel), "u": g(d["u"])})
    torch.cuda.synchronize()
    return {"Lx": Lx.cpu().numpy(), "Lv": Lv.cpu().numpy(), "px": px.cpu().numpy(), "x_next": outs[0].cpu().numpy()}


def parity_report(P, n, seed=1, log_jac=False, dyn=None):
    """Errors of the kernel and of the fp32 oracle, both against the fp64 oracle, on the same inputs."""
    d = P.draws(n, seed)
    r64 = run_oracle_propose(P, d, torch.float64, log_jac)
    r32 = run_oracle_propose(P, d, torch.float32, log_jac)
    rk = run_kernel_propose(P, d, dyn=dyn, log_jac=log_jac)
    rep = {}
    for key in ("Lx", "Lv"):
        rep[key + "_kernel"] = max_rel(rk[key], r64[key])
        rep[key + "_o32"] = max_rel(r32[key], r64[key])
    ok = np.isfinite(r64["px"])
    rep["px_kernel"] = float(np.max(np.abs(rk["px"][ok] - r64["px"][ok])))
    rep["px_o32"] = float(np.max(np.abs(r32["px"][ok] - r64["px"][ok])))
    rep["px_mean_kernel"] = float(abs(rk["px"][ok].astype(np.float64).mean() - r64["px"][ok].mean()))
    rep["px_mean_o32"] = float(abs(r32["px"][ok].astype(np.float64).mean() - r64["px"][ok].mean()))
    rep["px_mean_ref"] = float(r64["px"][ok].mean())
    # accept decisions can legitimately flip only where |px - u| is inside the fp32 noise
    acc64 = (r64["px"] - d["u"]) >= 0
    acck = np.all(rk["x_next"] == rk["Lx"], axis=1)
    margin = np.abs(r64["px"] - d["u"])
    flips = (acc64 != acck) & (margin > 1e-4)
    rep["accept_flips_outside_noise"] = int(flips.sum())
    rep["n"] = n
    return rep, (d, r64, r32, rk)
