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


# ---- targets -------------------------------------------------------------------------------------
def scg2_cov():
    return np.array([[50.05, -49.95], [-49.95, 50.05]])  # SCGExperiment.ipynb:105


def scg_cov(D, seed=0):
    """Builder-defined D-dim strongly correlated Gaussian: spectrum logspace(2,-1), random rotation."""
    from scipy.stats import ortho_group
    if D == 2:
        return scg2_cov()
    R = ortho_group.rvs(D, random_state=seed)
    return R.T @ np.diag(np.logspace(2, -1, D)) @ R


def target(kind, D, **kw):
    """Returns (product distribution object, oracle Energy, x0 sampler(n, rng))."""
    from l2hmc_b200 import distributions as dist
    if kind == "gaussian":
        cov = scg_cov(D, kw.get("seed", 0))
        mu = np.asarray(kw.get("mu", np.zeros(D)), dtype=np.float64)
        g = dist.Gaussian(mu, cov)
        en = O.GaussianEnergy(mu.astype(np.float32), g.i_sigma.astype(np.float32))
        L = np.linalg.cholesky(cov)
        return g, en, (lambda n, rng: (rng.standard_normal((n, D)) @ L.T + mu).astype(np.float32))
    if kind == "gmm":
        var = kw.get("var", 0.1)
        mus = [np.array([-2.0, 0.0] + [0.0] * (D - 2)), np.array([2.0, 0.0] + [0.0] * (D - 2))]
        sig = [var * np.eye(D), var * np.eye(D)]
        g = dist.GMM(mus, sig, [0.5, 0.5])
        en = O.GMMEnergy(mus, g.i_sigmas, g.constants)

        def x0(n, rng):
            c = rng.integers(0, 2, n)
            return (np.stack(mus)[c] + np.sqrt(var) * rng.standard_normal((n, D))).astype(np.float32)
        return g, en, x0
    if kind == "roughwell":
        g = dist.RoughWell(D, kw.get("eps", 0.1), easy=kw.get("easy", False))
        en = O.RoughWellEnergy(g.eps, g.easy)
        return g, en, (lambda n, rng: rng.standard_normal((n, D)).astype(np.float32))
    if kind == "funnel":
        g = dist.GaussianFunnel(dim=D)
        en = O.FunnelEnergy(g.sigma, g.clip)

        def x0(n, rng):
            x = rng.standard_normal((n, D)).astype(np.float32)
            x[:, 0] *= 2.0
            return x
        return g, en, x0
    raise ValueError(kind)


# ---- problems ---------------------------------------------------------------------------------------
class Problem:
    """One synthetic configuration: weights, masks, target; builds the oracle and the product object."""

    def __init__(self, kind="gaussian", D=2, H=10, T=10, eps=0.1, regime="init", hmc=False, seed=0, **kw):
        self.kind, self.D, self.H, self.T, self.eps, self.hmc = kind, D, H, T, eps, hmc
        rng = np.random.default_rng(seed)
        self.dist, self.energy, self.x0 = target(kind, D, **kw)
        self.mask = O.make_masks(rng, T, D)
        self.xnet = None if hmc else O.make_net(rng, D, H, 2.0, regime)
        self.vnet = None if hmc else O.make_net(rng, D, H, 1.0, regime)
        self.rng = rng

    def oracle(self, dtype=torch.float64, temperature=1.0):
        return O.OracleDynamics(self.D, self.T, self.eps, self.energy, self.mask, self.xnet, self.vnet,
                                hmc=self.hmc, temperature=temperature, dtype=dtype)

    def net_factory(self):
        from l2hmc_b200.layers import Linear, Sequential, Zip, Parallel, ScaleTanh, relu, load_stq_net
        H = self.H
        params = {"XNet": self.xnet, "VNet": self.vnet}

        def network(x_dim, scope, factor):  # SCGExperiment.ipynb:51-77 with width H
            net = Sequential([
                Zip([
                    Linear(x_dim, H, scope='embed_1', factor=1.0 / 3),
                    Linear(x_dim, H, scope='embed_2', factor=factor * 1.0 / 3),
                    Linear(2, H, scope='embed_3', factor=1.0 / 3),
                    lambda _: 0.,
                ]),
                sum,
                relu,
                Linear(H, H, scope='linear_1'),
                relu,
                Parallel([
                    Sequential([Linear(H, x_dim, scope='linear_s', factor=0.001), ScaleTanh(x_dim, scope='scale_s')]),
                    Linear(H, x_dim, scope='linear_t', factor=0.001),
                    Sequential([Linear(H, x_dim, scope='linear_f', factor=0.001), ScaleTanh(x_dim, scope='scale_f')]),
                ])
            ])
            load_stq_net(net, params[scope])
            return net
        return network

    def product(self, **kw):
        from l2hmc_b200 import Dynamics
        d = Dynamics(self.D, self.dist.get_energy_function(), T=self.T, eps=self.eps, hmc=self.hmc,
                     net_factory=None if self.hmc else self.net_factory(), **kw)
        d.mask = self.mask
        return d

    def draws(self, n, seed=1):
        rng = np.random.default_rng(seed)
        return {
            "x": self.x0(n, rng),
            "v_f": rng.standard_normal((n, self.D)).astype(np.float32),
            "v_b": rng.standard_normal((n, self.D)).astype(np.float32),
            "dir": rng.integers(0, 2, n).astype(np.uint8),
            "u": rng.random(n).astype(np.float32),
        }


class VaeProblem:
    """BASELINE config 5 in miniature or at full layer sizes: the decoder-Bernoulli posterior target of
    mnist_vae.py:104-126 with S/T/Q nets that add a shared softplus-MLP encoding of aux to their first stage
    (mnist_vae.py:134-167).  Random weights, Bernoulli(0.5) aux rows (no dataset here)."""
    hmc = False

    def __init__(self, D=8, H=24, T=4, eps=0.1, dec=(64, 64), aux_dim=40, enc=(32, 32), regime="stress", seed=0,
                 use_encoder=True):
        self.kind, self.D, self.H, self.T, self.eps = "decoder", D, H, T, eps
        rng = np.random.default_rng(seed)
        self.aux_dim = aux_dim
        self.dec_w = [D] + list(dec) + [aux_dim]
        self.dec_W, self.dec_b = O.make_softplus_mlp(rng, self.dec_w, last_factor=0.01)
        self.use_encoder = use_encoder
        if use_encoder:
            self.enc_w = [aux_dim] + list(enc) + [H]
            self.enc_W, self.enc_b = O.make_softplus_mlp(rng, self.enc_w)
        self.mask = O.make_masks(rng, T, D)
        self.xnet = O.make_net(rng, D, H, 2.0, regime)
        self.vnet = O.make_net(rng, D, H, 1.0, regime)

    def oracle_for(self, d, dtype=torch.float64, temperature=1.0):
        en = O.DecoderBernoulliEnergy(self.dec_W, self.dec_b, d["aux"], dtype)
        dyn = O.OracleDynamics(self.D, self.T, self.eps, en, self.mask, self.xnet, self.vnet, hmc=False,
                               temperature=temperature, dtype=dtype)
        ae = None
        if self.use_encoder:
            tt = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)  # noqa: E731
            ae = O.softplus_mlp([tt(W) for W in self.enc_W], [tt(b) for b in self.enc_b], tt(d["aux"]))
        return dyn, ae

    @staticmethod
    def _mlp(widths, Ws, bs, scope):
        from l2hmc_b200.layers import Linear, Sequential, softplus
        layers = []
        for i in range(len(Ws)):
            l = Linear(widths[i], widths[i + 1], scope="%s_%d" % (scope, i + 1))
            l.W = torch.as_tensor(Ws[i]).clone()
            l.b = torch.as_tensor(bs[i]).clone()
            layers.append(l)
            if i + 1 < len(Ws):
                layers.append(softplus)
        return Sequential(layers)

    def net_factory(self):
        from l2hmc_b200.layers import Linear, Sequential, Zip, Parallel, ScaleTanh, relu, load_stq_net
        H = self.H
        params = {"XNet": self.xnet, "VNet": self.vnet}
        encoder_sampler = self._mlp(self.enc_w, self.enc_W, self.enc_b, "encoder") if self.use_encoder else (lambda _: 0.)

        def net_factory(x_dim, scope, factor):  # mnist_vae.py:142-167
            net = Sequential([
                Zip([
                    Linear(x_dim, H, scope='embed_1', factor=0.33),
                    Linear(x_dim, H, scope='embed_2', factor=factor * 0.33),
                    Linear(2, H, scope='embed_3', factor=0.33),
                    encoder_sampler,
                ]),
                sum,
                relu,
                Linear(H, H, scope='linear_1'),
                relu,
                Parallel([
                    Sequential([Linear(H, x_dim, scope='linear_s', factor=0.01), ScaleTanh(x_dim, scope='scale_s')]),
                    Linear(H, x_dim, scope='linear_t', factor=0.01),
                    Sequential([Linear(H, x_dim, scope='linear_f', factor=0.01), ScaleTanh(x_dim, scope='scale_f')]),
                ])
            ])
            load_stq_net(net, params[scope])
            return net
        return net_factory

    def product(self, **kw):
        from l2hmc_b200 import Dynamics
        from l2hmc_b200.vae import DecoderEnergy
        energy = DecoderEnergy(self._mlp(self.dec_w, self.dec_W, self.dec_b, "decoder"))
        d = Dynamics(self.D, energy, T=self.T, eps=self.eps, net_factory=self.net_factory(), **kw)
        d.mask = self.mask
        return d

    def draws(self, n, seed=1):
        rng = np.random.default_rng(seed)
        return {
            "x": rng.standard_normal((n, self.D)).astype(np.float32),  # latent prior, like init_x = latent_q
            "aux": (rng.random((n, self.aux_dim)) < 0.5).astype(np.float32),
            "v_f": rng.standard_normal((n, self.D)).astype(np.float32),
            "v_b": rng.standard_normal((n, self.D)).astype(np.float32),
            "dir": rng.integers(0, 2, n).astype(np.uint8),
            "u": rng.random(n).astype(np.float32),
        }


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


VAE_CONFIGS = {
    "c5_vae_mini": dict(D=8, H=24, T=4, dec=(64, 64), aux_dim=40, enc=(32, 32)),
    "c5_vae_ragged": dict(D=7, H=21, T=3, dec=(33,), aux_dim=19, enc=(10,)),   # nothing a multiple of 8
    "c5_vae_noenc": dict(D=8, H=24, T=4, dec=(64, 64), aux_dim=40, use_encoder=False),
    # the layer sizes of mnist_vae.py: latent 50, decoder 1024-1024-784, encoder 512-512-200, nets 200 wide, Lf=15
    "c5_vae_full": dict(D=50, H=200, T=15, dec=(1024, 1024), aux_dim=784, enc=(512, 512)),
}


def t64(a):
    return torch.as_tensor(np.asarray(a)).to(torch.float64)


def max_rel(a, b):
    """max |a-b| / max(1, max|b|): the north star's 'relative fp32 tolerance' on a whole array."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


CONFIGS = {
    # name: Problem kwargs  (BASELINE.json configs, reduced where noted by the tests)
    "c1_scg2": dict(kind="gaussian", D=2, H=10, T=10, eps=0.1),
    "c2_scg50": dict(kind="gaussian", D=50, H=100, T=10, eps=0.1),
    "c3_mog2": dict(kind="gmm", D=2, H=10, T=25, eps=0.1),
    "c4_rw32": dict(kind="roughwell", D=32, H=100, T=10, eps=0.1, easy=True),
    # easy=False has curvature 1/eps_rw^3 = 1000: leapfrog is only stable below ~0.06, and at step 0.1
    # trajectories are chaotic (fp32 and fp64 oracles differ by O(1), accept prob 0), so the hard
    # variant is exercised at step 0.01 where parity is meaningful.
    "c4_rw32_hard": dict(kind="roughwell", D=32, H=100, T=10, eps=0.01, easy=False),
    "funnel3": dict(kind="funnel", D=3, H=10, T=10, eps=0.1),
}


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
