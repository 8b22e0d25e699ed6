"""Synthetic workloads of the L2HMC sampling path: the BASELINE.json configurations as
product-side objects -- target distribution, S/T/Q net weights, masks, start points -- built from seeds.

The reference ships one concrete problem (the 2-d strongly correlated Gaussian of SCGExperiment.ipynb:86-108) and no
weights; the others are defined here: `bench.py` and `__graft_entry__.smoke()` build their workloads from this module,
and the tests build the oracle's view of the *same arrays* from it (tests/util.py), so nothing under l2hmc_b200/ or in
the benchmark's product arm needs test infrastructure to exist.

    P = SyntheticProblem(regime="stress", **CONFIGS["c2_scg50"])
    dyn = P.product()                       # l2hmc_b200.Dynamics carrying P's nets, masks and target
    x0 = P.x0(n, np.random.default_rng(0))  # start points (exact target samples where the target allows)
"""
from __future__ import annotations

import math

import numpy as np
import torch


# ---- targets -----------------------------------------------------------------------------------------
def scg2_cov():
    return np.array([[50.05, -49.95], [-49.95, 50.05]])  # SCGExperiment.ipynb:105


def scg_cov(D, seed=0):
    """D-dim strongly correlated Gaussian (BASELINE config 2 has no constructor in the reference): spectrum
    logspace(2, -1) -- the end points [100, 0.1] of the notebook's 2-d SCG -- under a random rotation."""
    from scipy.stats import ortho_group
    if D == 2:
        return scg2_cov()
    R = ortho_group.rvs(D, random_state=seed)
    return R.T @ np.diag(np.logspace(2, -1, D)) @ R


def target(kind, D, **kw):
    """Returns (product distribution object, x0 sampler(n, rng) -> float32 [n, D])."""
    from . import distributions as dist
    if kind == "gaussian":
        cov = scg_cov(D, kw.get("seed", 0))
        mu = np.asarray(kw.get("mu", np.zeros(D)), dtype=np.float64)
        g = dist.Gaussian(mu, cov)
        L = np.linalg.cholesky(cov)
        return g, (lambda n, rng: (rng.standard_normal((n, D)) @ L.T + mu).astype(np.float32))
    if kind == "gmm":
        var = kw.get("var", 0.1)
        mus = [np.array([-2.0, 0.0] + [0.0] * (D - 2)), np.array([2.0, 0.0] + [0.0] * (D - 2))]
        sig = [var * np.eye(D), var * np.eye(D)]
        g = dist.GMM(mus, sig, [0.5, 0.5])

        def x0(n, rng):
            c = rng.integers(0, 2, n)
            return (np.stack(mus)[c] + np.sqrt(var) * rng.standard_normal((n, D))).astype(np.float32)
        return g, x0
    if kind == "roughwell":
        g = dist.RoughWell(D, kw.get("eps", 0.1), easy=kw.get("easy", False))
        return g, (lambda n, rng: rng.standard_normal((n, D)).astype(np.float32))
    if kind == "funnel":
        g = dist.GaussianFunnel(dim=D)

        def x0(n, rng):
            x = rng.standard_normal((n, D)).astype(np.float32)
            x[:, 0] *= 2.0
            return x
        return g, x0
    raise ValueError(kind)


# ---- weights and masks ----------------------------------------------------------------------------------
def trunc_normal(rng: np.random.Generator, shape, std):
    """variance_scaling_initializer(uniform=False) draws a truncated normal (+-2 sigma), utils/layers.py:32."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


def make_net(rng, D, H, factor, regime="init"):
    """Weights of one S/T/Q net.  regime 'init' follows SCGExperiment.ipynb:55-71 (embed factors 1/3,
    factor/3, 1/3; hidden 1.0; heads 0.001; biases 0; log-scales 0).  'stress' uses head factor 0.003, N(0, 0.05^2)
    biases and log-scales ~ U(-0.5, 0.5): S, Q are O(0.1), log|J| is O(1) and accept probabilities spread over (0, 1)
    instead of sitting near the untrained value -- every term of the update is exercised (parity tests)."""
    def lin(i, o, f):
        return trunc_normal(rng, (i, o), math.sqrt(1.3 * 2.0 * f / i))
    hf = 0.001 if regime == "init" else 0.003
    p = {
        "W1": lin(D, H, 1.0 / 3), "W2": lin(D, H, factor / 3.0), "W3": lin(2, H, 1.0 / 3),
        "W4": lin(H, H, 1.0), "Ws": lin(H, D, hf), "Wt": lin(H, D, hf), "Wq": lin(H, D, hf),
    }
    for k, n in (("b1", H), ("b2", H), ("b3", H), ("b4", H), ("bs", D), ("bt", D), ("bq", D)):
        p[k] = (np.zeros(n, np.float32) if regime == "init"
                else (0.05 * rng.standard_normal(n)).astype(np.float32))
    for k in ("ls", "lq"):
        p[k] = (np.zeros(D, np.float32) if regime == "init"
                else rng.uniform(-0.5, 0.5, D).astype(np.float32))
    return p


def make_masks(rng, T, D):
    """_init_mask (utils/dynamics.py:84-93): floor(D/2) ones at a random permutation's head, per step."""
    m = np.zeros((T, D), np.float32)
    for t in range(T):
        m[t, rng.permutation(D)[: int(D / 2)]] = 1.0
    return m


def make_softplus_mlp(rng, widths, last_factor=1.0):
    """Weights of a Linear/softplus stack with the reference initialiser (utils/layers.py:29-37,
    factor 1.0; the decoder's last layer uses factor 0.01, mnist_vae.py:110); small biases so the
    synthetic problem exercises them."""
    Ws, bs = [], []
    for i in range(len(widths) - 1):
        f = last_factor if i == len(widths) - 2 else 1.0
        Ws.append(trunc_normal(rng, (widths[i], widths[i + 1]), math.sqrt(1.3 * 2.0 * f / widths[i])))
        bs.append((0.05 * rng.standard_normal(widths[i + 1])).astype(np.float32))
    return Ws, bs


# ---- problems ---------------------------------------------------------------------------------------------
class SyntheticProblem:
    """One closed-form-target configuration: weights, masks, target; builds the product ``Dynamics``."""

    def __init__(self, kind="gaussian", D=2, H=10, T=10, eps=0.1, regime="init", hmc=False, seed=0, **kw):
        self.kind, self.D, self.H, self.T, self.eps, self.hmc = kind, D, H, T, eps, hmc
        self.regime = regime
        rng = np.random.default_rng(seed)
        self.dist, self.x0 = target(kind, D, **kw)
        self.mask = make_masks(rng, T, D)
        self.xnet = None if hmc else make_net(rng, D, H, 2.0, regime)
        self.vnet = None if hmc else make_net(rng, D, H, 1.0, regime)
        self.rng = rng

    def net_factory(self):
        from .layers import Linear, Sequential, Zip, Parallel, ScaleTanh, relu, load_stq_net
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
        from .dynamics import Dynamics
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


class SyntheticVaeProblem:
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
        self.dec_W, self.dec_b = make_softplus_mlp(rng, self.dec_w, last_factor=0.01)
        self.use_encoder = use_encoder
        if use_encoder:
            self.enc_w = [aux_dim] + list(enc) + [H]
            self.enc_W, self.enc_b = make_softplus_mlp(rng, self.enc_w)
        self.mask = make_masks(rng, T, D)
        self.xnet = make_net(rng, D, H, 2.0, regime)
        self.vnet = make_net(rng, D, H, 1.0, regime)

    @staticmethod
    def _mlp(widths, Ws, bs, scope):
        from .layers import Linear, Sequential, softplus
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
        from .layers import Linear, Sequential, Zip, Parallel, ScaleTanh, relu, load_stq_net
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
        from .dynamics import Dynamics
        from .vae import DecoderEnergy
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


CONFIGS = {
    # name: SyntheticProblem kwargs  (BASELINE.json configs; chain counts are the caller's)
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

VAE_CONFIGS = {
    "c5_vae_mini": dict(D=8, H=24, T=4, dec=(64, 64), aux_dim=40, enc=(32, 32)),
    "c5_vae_ragged": dict(D=7, H=21, T=3, dec=(33,), aux_dim=19, enc=(10,)),   # nothing a multiple of 8
    "c5_vae_noenc": dict(D=8, H=24, T=4, dec=(64, 64), aux_dim=40, use_encoder=False),
    # the layer sizes of mnist_vae.py: latent 50, decoder 1024-1024-784, encoder 512-512-200, nets 200 wide, Lf=15
    "c5_vae_full": dict(D=50, H=200, T=15, dec=(1024, 1024), aux_dim=784, enc=(512, 512)),
}
