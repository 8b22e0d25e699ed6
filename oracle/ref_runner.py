"""Run the UNMODIFIED reference (utils/dynamics.py, sampler.py, layers.py, distributions.py, losses.py,
ais.py, func_utils.py, loaded by oracle/ref_loader.py on the eager TensorFlow stand-in) on one synthetic
problem with injected parameters and randomness.  TEST INFRASTRUCTURE ONLY.

Used by tests/golden/make_ref_golden.py to write tests/golden/ref_*.npz, and by tests/test_reference_pin.py
(when /root/reference is present) to check the fixtures are reproducible.  Inputs are the plain arrays of
a tests/util.py ``Problem`` (weights, masks, target parameters, draws); everything between those arrays
and the returned outputs is reference code.
"""
from __future__ import annotations

import numpy as np

import ref_loader

# oracle key -> (reference variable scope under <Net>/, variable name)   SCGExperiment.ipynb:52-74, utils/layers.py:31-34,83-84
NET_VARS = {
    "W1": "embed_1/W", "b1": "embed_1/b", "W2": "embed_2/W", "b2": "embed_2/b", "W3": "embed_3/W", "b3": "embed_3/b",
    "W4": "linear_1/W", "b4": "linear_1/b", "Ws": "linear_s/W", "bs": "linear_s/b", "Wt": "linear_t/W", "bt": "linear_t/b",
    "Wq": "linear_f/W", "bq": "linear_f/b", "ls": "scale_s/scale", "lq": "scale_f/scale",
}


def eps_fp32(eps):
    """What the reference's fp32 graph evaluates self.eps to: exp(log(fp32(eps))) in fp32 (utils/dynamics.py:50-58)."""
    return np.exp(np.log(np.float32(eps), dtype=np.float32), dtype=np.float32)


def preload_nets(tf, xnet, vnet, prefix=""):
    vals = {}
    for scope, p in (("XNet", xnet), ("VNet", vnet)):
        for k, name in NET_VARS.items():
            a = np.asarray(p[k], dtype=np.float32)
            if k in ("ls", "lq"):
                a = a.reshape(1, -1)  # ScaleTanh's variable is [1, in_] (utils/layers.py:84)
            vals["%s%s/%s" % (prefix, scope, name)] = a
    tf.shim.preload(vals)


def width_network(ref, H):
    """SCGExperiment.ipynb's `network` with the hard-coded width 10 replaced by H (BASELINE configs 2 and 4
    use width 100; the cell itself is executed verbatim as ref.notebook_network for H = 10)."""
    tf, L = ref.tf, ref.layers

    def network(x_dim, scope, factor):
        with tf.variable_scope(scope):
            net = L.Sequential([
                L.Zip([
                    L.Linear(x_dim, H, scope='embed_1', factor=1.0 / 3),
                    L.Linear(x_dim, H, scope='embed_2', factor=factor * 1.0 / 3),
                    L.Linear(2, H, scope='embed_3', factor=1.0 / 3),
                    lambda _: 0.,
                ]),
                sum,
                tf.nn.relu,
                L.Linear(H, H, scope='linear_1'),
                tf.nn.relu,
                L.Parallel([
                    L.Sequential([L.Linear(H, x_dim, scope='linear_s', factor=0.001), L.ScaleTanh(x_dim, scope='scale_s')]),
                    L.Linear(H, x_dim, scope='linear_t', factor=0.001),
                    L.Sequential([L.Linear(H, x_dim, scope='linear_f', factor=0.001), L.ScaleTanh(x_dim, scope='scale_f')]),
                ])
            ])
        return net
    return network


def reference_distribution(ref, P):
    """The reference's own distribution object for a tests/util.py Problem (constructed from the same
    numbers the product object gets: utils/distributions.py:41-48, 84-88, 104-123, 155-159)."""
    D = ref.distributions
    g = P.dist
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # Gaussian.__init__ prints det(sigma) (utils/distributions.py:46)
        if P.kind == "gaussian":
            return D.Gaussian(np.asarray(g.mu), np.asarray(g.sigma))
        if P.kind == "gmm":
            return D.GMM([np.asarray(m) for m in g.mus], [np.asarray(s) for s in g.sigmas], list(g.pis))
        if P.kind == "roughwell":
            return D.RoughWell(g.dim, g.eps, easy=g.easy)
        if P.kind == "funnel":
            return D.GaussianFunnel(dim=g.dim)
    raise ValueError(P.kind)


def build_dynamics(P, real="float64", temperature=None, energy_function=None, net_factory=None, verbatim_cell=True):
    """A fresh reference ``Dynamics`` carrying P's parameters.  Returns (ref, dynamics)."""
    ref = ref_loader.load()
    tf = ref.tf
    tf.shim.reset()
    tf.shim.set_real(real)
    import contextlib
    import io
    if energy_function is None:
        dist = reference_distribution(ref, P)
        with contextlib.redirect_stdout(io.StringIO()):  # GaussianFunnel.get_energy_function prints (:162)
            energy_function = dist.get_energy_function()
    e32 = eps_fp32(P.eps)
    if not P.hmc:
        preload_nets(tf, P.xnet, P.vnet)
        # alpha is the trained variable; the fp32 graph holds log(fp32 eps).  For the fp64 run alpha is chosen so
        # that exp(alpha) is the fp32 graph's eps: same parameters, wider arithmetic.
        tf.shim.preload({"alpha": np.log(np.float32(P.eps), dtype=np.float32) if real == "float32"
                         else np.log(np.float64(e32))})
        if net_factory is None:
            net_factory = ref.notebook_network if (P.H == 10 and verbatim_cell) else width_network(ref, P.H)
    # _init_mask draws from numpy's global RNG (utils/dynamics.py:88); the mask is replaced right after, as
    # eval_sampler.py:156 does, so the draw only has to be harmless
    st = np.random.get_state()
    dyn = ref.dynamics.Dynamics(P.D, energy_function, T=P.T, eps=float(e32) if P.hmc else P.eps, hmc=P.hmc,
                                net_factory=net_factory, use_temperature=temperature is not None)
    np.random.set_state(st)
    dyn.mask = tf.constant(np.asarray(P.mask), dtype=tf.float32)
    if temperature is not None:
        tf.shim.feed(dyn.temperature, np.float32(temperature))
    return ref, dyn


def _np(t):
    return None if t is None else t.numpy()


def run_propose(P, d, real="float64", log_jac=False, temperature=None, aux=None, energy_function=None,
                net_factory=None):
    """utils/sampler.py:28-51 `propose(x, dynamics, init_v, aux, do_mh_step=True, log_jac)` on injected draws.
    Draw order of the reference: direction bits (:34), v of forward (dynamics.py:248), v of backward (:276),
    accept uniforms (sampler.py:54)."""
    ref, dyn = build_dynamics(P, real, temperature, energy_function, net_factory)
    tf = ref.tf
    x = tf.shim.input(d["x"])
    if P.hmc:
        tf.shim.feed_random(uniform=[d["u"]])
        Lx, Lv, px, outs = ref.sampler.propose(x, dyn, init_v=tf.constant(d["v_f"], dtype=tf.float32), aux=aux,
                                               do_mh_step=True)
    else:
        tf.shim.feed_random(normal=[d["v_f"], d["v_b"]],
                            uniform=[d["dir"].astype(np.int32).reshape(-1, 1), d["u"]])
        # init_v only decides whether Lv is returned (sampler.py:40-42); forward/backward draw their own v
        Lx, Lv, px, outs = ref.sampler.propose(x, dyn, init_v=tf.constant(d["v_f"], dtype=tf.float32), aux=aux,
                                               do_mh_step=True, log_jac=log_jac)
    assert tf.shim.pending_random() == (0, 0), "the reference drew fewer random arrays than injected"
    return {"Lx": _np(Lx), "Lv": _np(Lv), "px": _np(px), "x_next": _np(outs[0])}


def run_methods(P, d, real="float64", temperature=None):
    """The Dynamics methods one by one (utils/dynamics.py:107-108, 203-218, 246-309) on the same inputs."""
    ref, dyn = build_dynamics(P, real, temperature)
    tf = ref.tf
    x = tf.shim.input(d["x"])
    v = tf.constant(d["v_f"], dtype=tf.float32)
    out = {"energy": _np(dyn.energy(x)), "grad_energy": _np(dyn.grad_energy(x)), "kinetic": _np(dyn.kinetic(v)),
           "hamiltonian": _np(dyn.hamiltonian(x, v))}
    fx, fv, fp = dyn.forward(x, init_v=v)
    _, _, fj = dyn.forward(x, init_v=v, log_jac=True)
    bx, bv, bp = dyn.backward(x, init_v=v)
    _, _, bj = dyn.backward(x, init_v=v, log_jac=True)
    out.update(fwd_x=_np(fx), fwd_v=_np(fv), fwd_p=_np(fp), fwd_logjac=_np(fj),
               bwd_x=_np(bx), bwd_v=_np(bv), bwd_p=_np(bp), bwd_logjac=_np(bj))
    out["p_accept"] = _np(dyn.p_accept(x, v, fx, fv, fj))
    if not P.hmc:
        # one raw net call each (utils/dynamics.py:119,131): [S, T, Q] of VNet([x, grad, t]) and XNet([v, m*x, t])
        t = dyn._format_time(tf.constant(1., dtype=tf.float32), tile=tf.shape(x)[0])
        m, mb = dyn._get_mask(tf.constant(1., dtype=tf.float32))
        S = dyn.VNet([x, dyn.grad_energy(x), t, None])
        X = dyn.XNet([v, m * x, t, None])
        for k, val in zip(("vnet_S", "vnet_T", "vnet_Q"), S):
            out[k] = _np(val)
        for k, val in zip(("xnet_S", "xnet_T", "xnet_Q"), X):
            out[k] = _np(val)
    return out


def run_chain_operator(P, x, nb_steps, init_v, directions, v_fs, v_bs, u, real="float64"):
    """utils/sampler.py:57-85.  The reference draws init_v itself (`if not init_v`, :58-59: a tensor there would
    raise in TF1), then per sub-proposal direction bits, forward v, backward v; the final accept uniforms last."""
    ref, dyn = build_dynamics(P, real)
    tf = ref.tf
    normal, uniform = [init_v], []
    for s in range(nb_steps):
        uniform.append(np.asarray(directions[s]).astype(np.int32).reshape(-1, 1))
        normal += [v_fs[s], v_bs[s]]
    uniform.append(u)
    tf.shim.feed_random(normal=normal, uniform=uniform)
    fx, fv, p, outs = ref.sampler.chain_operator(tf.shim.input(x), dyn, nb_steps, do_mh_step=True)
    assert tf.shim.pending_random() == (0, 0)
    return {"final_x": _np(fx), "final_v": _np(fv), "p_accept": _np(p), "x_next": _np(outs[0])}


def run_notebook_loss(P, x, z, rx, rz, scale=0.1, real="float64"):
    """The training objective of SCGExperiment.ipynb (the cell building `loss`, lines 159-169 of the .ipynb) and
    its gradient with respect to every trainable variable (what optimizer.minimize differentiates, :186-188).
    rx / rz: draws {dir, v_f, v_b} for the x batch and the z batch; z itself replaces tf.random_normal (:157)."""
    ref, dyn = build_dynamics(P, real)
    tf = ref.tf
    xt = tf.shim.input(x)
    # draw order of the cell: z (random_normal), then propose(x): dir, v_f, v_b, u (do_mh_step=True); propose(z): dir, v_f, v_b
    tf.shim.feed_random(normal=[z, rx["v_f"], rx["v_b"], rz["v_f"], rz["v_b"]],
                        uniform=[rx["dir"].astype(np.int32).reshape(-1, 1), rx["u"],
                                 rz["dir"].astype(np.int32).reshape(-1, 1)])
    zt = tf.random_normal(tf.shape(xt))
    Lx, _, px, output = ref.sampler.propose(xt, dyn, do_mh_step=True)
    Lz, _, pz, _ = ref.sampler.propose(zt, dyn, do_mh_step=False)
    loss = 0.
    v1 = (tf.reduce_sum(tf.square(xt - Lx), axis=1) * px) + 1e-4
    v2 = (tf.reduce_sum(tf.square(zt - Lz), axis=1) * pz) + 1e-4
    loss += scale * (tf.reduce_mean(1.0 / v1) + tf.reduce_mean(1.0 / v2))
    loss += (- tf.reduce_mean(v1) - tf.reduce_mean(v2)) / scale
    names = list(tf.shim.variables.keys())
    vs = [tf.shim.variables[k] for k in names]
    grads = tf.gradients(loss, vs)
    assert tf.shim.pending_random() == (0, 0)
    out = {"loss": _np(loss), "x_next": _np(output[0]), "px": _np(px), "pz": _np(pz)}
    for k, g in zip(names, grads):
        out["grad/" + k] = np.zeros(tf.shim.variables[k].shape) if g is None else _np(g)
    return out


def run_losses(x, X, p, real="float64"):
    """utils/losses.py:26-59 on plain arrays."""
    ref = ref_loader.load()
    tf = ref.tf
    tf.shim.set_real(real)
    c = lambda a: tf.constant(np.asarray(a), dtype=tf.float32)  # noqa: E731
    return {name: _np(ref.losses.get_loss(name)(c(x), c(X), c(p))) for name in ("mixed", "standard", "inverse", "logsumexp")}


# ---- BASELINE config 5: the VAE posterior target (mnist_vae.py) ---------------------------------------------
def _mnist_vae_text(first, last):
    """Lines first..last (1-based, inclusive) of mnist_vae.py, dedented: the file itself cannot be imported (Python-2
    print statements, flags, MNIST download at import), and the decoder / energy / sampler nets are locals of main()."""
    import os
    import textwrap
    with open(os.path.join(ref_loader.REF_ROOT, "mnist_vae.py")) as fh:
        lines = fh.read().split("\n")
    return textwrap.dedent("\n".join(lines[first - 1:last]))


def _vae_preload(tf, P, sampler_prefix="sampler/"):
    vals = {}
    for i, (W, b) in enumerate(zip(P.dec_W, P.dec_b)):
        vals["decoder/decoder_%d/W" % (i + 1)] = W
        vals["decoder/decoder_%d/b" % (i + 1)] = b
    if P.use_encoder:
        for i, (W, b) in enumerate(zip(P.enc_W, P.enc_b)):
            vals["%sencoder_%d/W" % (sampler_prefix, i + 1)] = W
            vals["%sencoder_%d/b" % (sampler_prefix, i + 1)] = b
    tf.shim.preload(vals)
    preload_nets(tf, P.xnet, P.vnet, prefix=sampler_prefix)


def build_vae_dynamics(P, real="float64", verbatim=None):
    """Reference Dynamics on the decoder-Bernoulli posterior.  verbatim=True executes mnist_vae.py's own text for the
    decoder (:104-111), energy(z, aux) (:122-126) and the sampler scope -- encoder_sampler, net_factory, the Dynamics
    constructor call (:130-178) -- which hard-codes the layer sizes 1024/784/512/200, so it needs P at those sizes
    (tests/util.py VAE_CONFIGS['c5_vae_full']).  Other sizes use the same constructs, restated with the reference's
    layer classes."""
    import types
    ref = ref_loader.load()
    tf, L = ref.tf, ref.layers
    full = (P.dec_w == [50, 1024, 1024, 784] and P.use_encoder and P.enc_w == [784, 512, 512, 200] and P.H == 200)
    verbatim = full if verbatim is None else verbatim
    tf.shim.reset()
    tf.shim.set_real(real)
    e32 = eps_fp32(P.eps)
    alpha = np.log(np.float32(P.eps), dtype=np.float32) if real == "float32" else np.log(np.float64(e32))
    st = np.random.get_state()
    if verbatim:
        assert full, "mnist_vae.py hard-codes its layer sizes"
        _vae_preload(tf, P, "sampler/")
        tf.shim.preload({"sampler/alpha": alpha})
        ns = {"tf": tf, "Dynamics": ref.dynamics.Dynamics,
              "hps": types.SimpleNamespace(latent_dim=P.D, leapfrogs=P.T, eps=P.eps, hmc=False)}
        for k in ("Linear", "Sequential", "Zip", "Parallel", "ScaleTanh"):
            ns[k] = getattr(L, k)
        exec(_mnist_vae_text(104, 111), ns)   # decoder
        exec(_mnist_vae_text(122, 126), ns)   # energy(z, aux)
        exec(_mnist_vae_text(130, 178), ns)   # with tf.variable_scope('sampler'): encoder_sampler, net_factory, dynamics
        dyn, energy = ns["dynamics"], ns["energy"]
    else:
        _vae_preload(tf, P, "")
        tf.shim.preload({"alpha": alpha})
        seq = []
        with tf.variable_scope('decoder'):
            for i in range(len(P.dec_W)):
                seq.append(L.Linear(P.dec_w[i], P.dec_w[i + 1], scope='decoder_%d' % (i + 1)))
                if i + 1 < len(P.dec_W):
                    seq.append(tf.nn.softplus)
            decoder = L.Sequential(seq)

        def energy(z, aux=None):  # mnist_vae.py:122-126
            logits = decoder(z)
            log_posterior = -tf.reduce_sum(tf.nn.sigmoid_cross_entropy_with_logits(labels=aux, logits=logits), axis=1)
            log_prior = -0.5 * tf.reduce_sum(tf.square(z), axis=1)
            return (-log_posterior - log_prior)

        if P.use_encoder:
            seq = []
            for i in range(len(P.enc_W)):
                seq.append(L.Linear(P.enc_w[i], P.enc_w[i + 1], scope='encoder_%d' % (i + 1)))
                if i + 1 < len(P.enc_W):
                    seq.append(tf.nn.softplus)
            encoder_sampler = L.Sequential(seq)
        else:
            encoder_sampler = lambda _: 0.  # noqa: E731
        H = P.H

        def net_factory(x_dim, scope, factor):  # mnist_vae.py:142-167
            with tf.variable_scope(scope):
                net = L.Sequential([
                    L.Zip([
                        L.Linear(x_dim, H, scope='embed_1', factor=0.33),
                        L.Linear(x_dim, H, scope='embed_2', factor=factor * 0.33),
                        L.Linear(2, H, scope='embed_3', factor=0.33),
                        encoder_sampler,
                    ]),
                    sum,
                    tf.nn.relu,
                    L.Linear(H, H, scope='linear_1'),
                    tf.nn.relu,
                    L.Parallel([
                        L.Sequential([L.Linear(H, x_dim, scope='linear_s', factor=0.01), L.ScaleTanh(x_dim, scope='scale_s')]),
                        L.Linear(H, x_dim, scope='linear_t', factor=0.01),
                        L.Sequential([L.Linear(H, x_dim, scope='linear_f', factor=0.01), L.ScaleTanh(x_dim, scope='scale_f')]),
                    ])
                ])
            return net
        dyn = ref.dynamics.Dynamics(P.D, energy, T=P.T, eps=P.eps, hmc=False, net_factory=net_factory,
                                    eps_trainable=True, use_temperature=False)
    np.random.set_state(st)
    dyn.mask = tf.constant(np.asarray(P.mask), dtype=tf.float32)
    return ref, dyn, energy


def run_vae_propose(P, d, real="float64"):
    """propose(init_x, dynamics, aux=inp, do_mh_step=True) as mnist_vae.py:204 calls it."""
    ref, dyn, energy = build_vae_dynamics(P, real)
    tf = ref.tf
    x = tf.shim.input(d["x"])
    aux = tf.constant(d["aux"], dtype=tf.float32)
    tf.shim.feed_random(normal=[d["v_f"], d["v_b"]], uniform=[d["dir"].astype(np.int32).reshape(-1, 1), d["u"]])
    Lx, Lv, px, outs = ref.sampler.propose(x, dyn, init_v=tf.constant(d["v_f"], dtype=tf.float32), aux=aux, do_mh_step=True)
    assert tf.shim.pending_random() == (0, 0)
    return {"Lx": _np(Lx), "Lv": _np(Lv), "px": _np(px), "x_next": _np(outs[0]),
            "energy": _np(energy(x, aux=aux)), "grad_energy": _np(dyn.grad_energy(x, aux=aux))}


def run_ais(e0_dist, e1_dist, anneal_steps, initial_x, v0, v_refresh, u, step_size, leapfrogs, real="float64"):
    """utils/ais.py:30-82 between two of the reference's distribution objects.  Draw order: the scan initialiser's
    random_normal first (:68), then per beta the refreshed momentum (:54) and the accept uniforms (:60)."""
    ref = ref_loader.load()
    tf = ref.tf
    tf.shim.reset()
    tf.shim.set_real(real)
    normal, uniform = [v0], []
    for s in range(anneal_steps):
        normal.append(v_refresh[s])
        uniform.append(u[s])
    tf.shim.feed_random(normal=normal, uniform=uniform)
    st = np.random.get_state()
    est, alpha = ref.ais.ais_estimate(e0_dist.get_energy_function(), e1_dist.get_energy_function(), anneal_steps,
                                      tf.shim.input(initial_x), step_size=step_size, leapfrogs=leapfrogs,
                                      x_dim=initial_x.shape[1])
    np.random.set_state(st)
    assert tf.shim.pending_random() == (0, 0)
    return {"estimate": _np(est), "mean_accept": _np(alpha)}
