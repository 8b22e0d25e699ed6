"""The oracle pinned to the reference's OWN code.

tests/golden/ref_*.npz and tests/golden/ref/*.npz are outputs of the unmodified /root/reference sources executed on
the eager TensorFlow stand-in (oracle/tf_shim, oracle/ref_loader.py, oracle/ref_runner.py; generator:
tests/golden/make_ref_golden.py).  Here:
  * both restatements (torch oracle, plain-C oracle) must reproduce those vectors -- fp64 against the fp64 reference
    run to rounding, so the oracle every other test compares with IS the reference's arithmetic;
  * where /root/reference is present (the build container, not the GPU box) the fixtures are regenerated and must be
    reproducible, and the stand-in's TF-1 semantics are unit-tested.
GPU counterpart: tests/test_gpu_reference_pin.py.
"""
import glob
import json
import os
import sys

import numpy as np
import pytest
import torch

import util as U

O = U.O
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ref_loader  # noqa: E402

have_reference = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")

PROPOSE_FILES = sorted(glob.glob(os.path.join(GOLD, "ref_*.npz")))


def _meta(z):
    return json.loads(bytes(z["meta"]).decode())


def _tol(name):
    # the hard rough well has curvature 1/eps^3 = 1000: two fp64 orderings of the same expression already differ by 1e-7
    return 2e-6 if "rw32_hard" in name else 1e-9


def test_fixture_inventory():
    names = {os.path.basename(f) for f in PROPOSE_FILES}
    for cfg in ("c1_scg2", "c2_scg50", "c3_mog2", "c4_rw32", "funnel3", "hmc_scg2"):
        assert any(cfg in n for n in names), cfg
    for extra in ("chain_operator_c1_n64", "chain_operator_c2_n32", "notebook_loss_c1_n200", "notebook_loss_c3_n64",
                  "losses_diagnostics", "ais_gauss3", "ais_roughwell4", "c5_vae_mini_n96", "c5_vae_full_n32"):
        assert os.path.exists(os.path.join(GOLD, "ref", extra + ".npz")), extra
    for f in PROPOSE_FILES:
        assert "unmodified /root/reference" in _meta(np.load(f))["source"]


@pytest.mark.parametrize("path", PROPOSE_FILES, ids=[os.path.basename(f)[:-4] for f in PROPOSE_FILES])
def test_oracle_reproduces_reference_propose(path):
    """utils/sampler.py:28-55 + utils/dynamics.py:115-309 as run by the reference itself == the torch oracle (fp64)."""
    import golden_io
    P, d, ref = golden_io.load(path)
    got = U.run_oracle_propose(P, d, torch.float64, P.meta["log_jac"])
    tol = _tol(path)
    for k in ("Lx", "Lv", "px", "x_next"):
        assert np.max(np.abs(got[k] - ref[k])) <= tol * max(1.0, np.abs(ref[k]).max()), (path, k)
    # the fp32 twin of the oracle against the reference's own fp32 run: same rounding noise floor (mostly bit-equal)
    z = np.load(path)
    got32 = U.run_oracle_propose(P, d, torch.float32, P.meta["log_jac"])
    for k in ("Lx", "Lv"):
        ref_noise = U.max_rel(z["out32_" + k], ref[k])
        assert U.max_rel(got32[k], ref[k]) <= 4 * ref_noise + 1e-6, (path, k)


@pytest.mark.parametrize("path", [f for f in PROPOSE_FILES if "logjac" not in f],
                         ids=[os.path.basename(f)[:-4] for f in PROPOSE_FILES if "logjac" not in f])
def test_oracle_reproduces_reference_methods(path):
    """Dynamics.energy / grad_energy / kinetic / hamiltonian / forward / backward / p_accept and one raw call of each
    net (utils/dynamics.py:107-108, 203-218, 246-309; utils/layers.py) against the oracle's methods."""
    import golden_io
    P, d, _ = golden_io.load(path)
    z = np.load(path)
    m = {k[2:]: z[k] for k in z.files if k.startswith("m_")}
    dyn = P.oracle(torch.float64)
    x, v = U.t64(d["x"]), U.t64(d["v_f"])
    tol = _tol(path)

    def close(a, b, what):
        assert np.max(np.abs(np.asarray(a) - b)) <= tol * max(1.0, np.abs(b).max()), (path, what)
    close(dyn.energy(x).numpy(), m["energy"], "energy")
    close(dyn.grad_energy(x).numpy(), m["grad_energy"], "grad_energy")   # analytic == tf.gradients of the closure
    close(dyn.kinetic(v).numpy(), m["kinetic"], "kinetic")
    close(dyn.hamiltonian(x, v).numpy(), m["hamiltonian"], "hamiltonian")
    fx, fv, fj = dyn.forward(x, v, log_jac=True)
    close(fx.numpy(), m["fwd_x"], "fwd_x"); close(fv.numpy(), m["fwd_v"], "fwd_v"); close(fj.numpy(), m["fwd_logjac"], "fwd_logjac")
    close(dyn.forward(x, v)[2].numpy(), m["fwd_p"], "fwd_p")
    bx, bv, bj = dyn.backward(x, v, log_jac=True)
    close(bx.numpy(), m["bwd_x"], "bwd_x"); close(bv.numpy(), m["bwd_v"], "bwd_v"); close(bj.numpy(), m["bwd_logjac"], "bwd_logjac")
    close(dyn.backward(x, v)[2].numpy(), m["bwd_p"], "bwd_p")
    close(dyn.p_accept(x, v, fx, fv, fj).numpy(), m["p_accept"], "p_accept")
    if not P.hmc:
        tau = dyn.format_time(1.0, x.shape[0])
        mk = torch.as_tensor(P.mask[1]).double()
        for key, val in zip(("vnet_S", "vnet_T", "vnet_Q"), O.net_apply(dyn.vnet, x, dyn.grad_energy(x), tau)):
            close(val.numpy(), m[key], key)
        for key, val in zip(("xnet_S", "xnet_T", "xnet_Q"), O.net_apply(dyn.xnet, v, mk * x, tau)):
            close(val.numpy(), m[key], key)


@pytest.mark.parametrize("path", [f for f in PROPOSE_FILES if "hmc" not in f],
                         ids=[os.path.basename(f)[:-4] for f in PROPOSE_FILES if "hmc" not in f])
def test_c_restatement_reproduces_reference(path):
    """The plain-C twin against the same reference vectors."""
    import c_oracle
    import golden_io
    P, d, ref = golden_io.load(path)
    co = c_oracle.COracle(P)
    got = co.propose(d, np.float64, log_jac=P.meta["log_jac"])
    tol = max(_tol(path), 1e-8)
    for k in ("Lx", "Lv", "px"):
        assert np.max(np.abs(got[k] - ref[k])) <= tol * max(1.0, np.abs(ref[k]).max()), (path, k)


@pytest.mark.parametrize("name", ["chain_operator_c1_n64", "chain_operator_c2_n32"])
def test_oracle_reproduces_reference_chain_operator(name):
    """utils/sampler.py:57-85 run by the reference (including its quirk: the carried v is ignored by the sub-proposals,
    the final p_accept uses init_v and the last blended Lv)."""
    z = np.load(os.path.join(GOLD, "ref", name + ".npz"))
    meta = _meta(z)
    P = U.Problem(regime=meta["regime"], **meta["kw"])
    P.mask = z["mask"]
    P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
    P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
    steps = meta["nb_steps"]
    fx, fv, p, outs = O.chain_operator(
        U.t64(z["in_x"]), P.oracle(torch.float64), steps, init_v=U.t64(z["in_init_v"]),
        directions=[torch.as_tensor(z["in_dirs"][s].astype(np.float64)) for s in range(steps)],
        v_fs=[U.t64(z["in_v_fs"][s]) for s in range(steps)], v_bs=[U.t64(z["in_v_bs"][s]) for s in range(steps)],
        u=U.t64(z["in_u"]), do_mh_step=True)
    for got, key in ((fx, "final_x"), (fv, "final_v"), (p, "p_accept"), (outs[0], "x_next")):
        assert np.max(np.abs(got.numpy() - z["out_" + key])) <= 1e-9 * max(1.0, np.abs(z["out_" + key]).max()), key


@pytest.mark.parametrize("name", ["notebook_loss_c1_n200", "notebook_loss_c3_n64"])
def test_oracle_reproduces_reference_training_objective_and_gradients(name):
    """SCGExperiment.ipynb's loss cell and tf.gradients of it w.r.t. every trainable variable (both nets + alpha), as
    computed by the reference code through the stand-in, against (a) autograd through the oracle and (b) the
    hand-written reverse sweep (oracle/l2hmc_reverse.py) the CUDA training kernels are written after."""
    import ref_runner as R
    z = np.load(os.path.join(GOLD, "ref", name + ".npz"))
    meta = _meta(z)
    P = U.Problem(regime=meta["regime"], **meta["kw"])
    P.mask = z["mask"]
    P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
    P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
    dyn = P.oracle(torch.float64)
    params = O.trainable_parameters(dyn)
    r = lambda pre: {"direction": torch.as_tensor(z[pre + "dir"].astype(np.float64)), "v_f": U.t64(z[pre + "v_f"]),  # noqa: E731
                     "v_b": U.t64(z[pre + "v_b"])}
    loss = O.notebook_loss(U.t64(z["in_x"]), U.t64(z["in_z"]), dyn, r("in_rx_"), r("in_rz_"), scale=meta["scale"])
    ref_loss = float(z["out_loss"])
    assert abs(float(loss.detach()) - ref_loss) <= 1e-9 * max(1.0, abs(ref_loss))
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    i = 0
    worst = 0.0
    for scope, net in (("XNet", dyn.xnet), ("VNet", dyn.vnet)):
        for k in sorted(net):
            g = grads[i]
            i += 1
            ref_g = z["out_grad__%s__%s" % (scope, R.NET_VARS[k].replace("/", "__"))].reshape(net[k].shape)
            got = np.zeros_like(ref_g) if g is None else g.numpy()
            err = np.max(np.abs(got - ref_g)) / max(1e-12, np.abs(ref_g).max())
            worst = max(worst, err)
            assert err <= 1e-7, (scope, k, err)
    assert np.abs(z["out_grad__alpha"]).max() > 0  # alpha = log(eps) is trained by the reference (utils/dynamics.py:50-54)
    # (b) the hand-written reverse sweep, alpha included
    import l2hmc_reverse as Rv
    dyn2 = P.oracle(torch.float64)
    loss_h, g_h = Rv.notebook_loss_and_grads(U.t64(z["in_x"]), U.t64(z["in_z"]), dyn2, r("in_rx_"), r("in_rz_"), scale=meta["scale"])
    assert abs(float(loss_h) - ref_loss) <= 1e-9 * max(1.0, abs(ref_loss))
    ga = float(np.asarray(z["out_grad__alpha"]).reshape(-1)[0])
    assert abs(float(g_h["alpha"]) - ga) <= 1e-7 * max(1.0, abs(ga))  # d loss / d alpha = eps * d loss / d eps
    for scope, key in (("XNet", "xnet"), ("VNet", "vnet")):
        for k, g in g_h[key].items():
            ref_g = z["out_grad__%s__%s" % (scope, R.NET_VARS[k].replace("/", "__"))].reshape(g.shape)
            assert np.max(np.abs(g.numpy() - ref_g)) <= 1e-7 * max(1e-12, np.abs(ref_g).max()), (scope, k)


def test_oracle_reproduces_reference_losses_and_diagnostics():
    """utils/losses.py:26-59 (all four of get_loss) and utils/func_utils.py:45-54,114-120."""
    z = np.load(os.path.join(GOLD, "ref", "losses_diagnostics.npz"))
    x, X, p = U.t64(z["in_x"]), U.t64(z["in_X"]), U.t64(z["in_p"])
    for name, fn in (("mixed", O.loss_mixed), ("standard", O.loss_std), ("inverse", O.loss_inverse),
                     ("logsumexp", O.loss_logsumexp)):
        assert abs(float(fn(x, X, p)) - float(z["loss_" + name])) <= 1e-9 * max(1.0, abs(float(z["loss_" + name]))), name
    trace, scale = z["in_trace"], float(z["in_scale"])
    # the reference's numpy of the time kept float32 arrays float32 under a python-float scale; NumPy 2 (which ran the
    # reference here) promotes to float64 -- the oracle documents that it follows the old rule, so compare at fp32 level
    spec = O.acl_spectrum(trace, scale)
    assert np.max(np.abs(spec - z["acl_spectrum"])) <= 2e-6 * np.abs(z["acl_spectrum"]).max()
    assert abs(O.ESS(z["acl_spectrum"]) - float(z["ess"])) <= 1e-12
    assert abs(O.autocovariance(trace, 3) - float(z["autocov_3"])) <= 2e-6 * abs(float(z["autocov_3"]))  # fp32 products


def test_oracle_reproduces_reference_ais():
    """utils/ais.py:30-82 (HMC-mode Dynamics per beta inside tf.scan) between two reference Gaussians."""
    z = np.load(os.path.join(GOLD, "ref", "ais_gauss3.npz"))
    meta = _meta(z)
    D = meta["D"]
    e0 = O.GaussianEnergy(np.zeros(D), np.eye(D))
    e1 = O.GaussianEnergy(z["mu1"].astype(np.float32), np.linalg.inv(z["cov1"]).astype(np.float32))
    est, alpha, _, _ = O.ais_estimate(e0, e1, meta["anneal_steps"], z["in_x"], step_size=meta["step_size"],
                                      leapfrogs=meta["leapfrogs"], v0=z["in_v0"], v_refresh=z["in_v_refresh"], u=z["in_u"])
    assert abs(float(est) - float(z["out_estimate"])) <= 1e-6 * max(1.0, abs(float(z["out_estimate"])))
    assert abs(float(alpha) - float(z["out_mean_accept"])) <= 1e-6


def test_oracle_reproduces_reference_ais_between_unlike_energies():
    """utils/ais.py:44-45 with a pair whose mixture is not one of the closed forms: Gaussian -> rough well."""
    z = np.load(os.path.join(GOLD, "ref", "ais_roughwell4.npz"))
    meta = _meta(z)
    D = meta["D"]
    e0 = O.GaussianEnergy(np.zeros(D), np.linalg.inv(z["cov0"]).astype(np.float32))
    e1 = O.RoughWellEnergy(meta["rw_eps"], meta["easy"])
    est, alpha, _, _ = O.ais_estimate(e0, e1, meta["anneal_steps"], z["in_x"], step_size=meta["step_size"],
                                      leapfrogs=meta["leapfrogs"], v0=z["in_v0"], v_refresh=z["in_v_refresh"], u=z["in_u"])
    assert abs(float(est) - float(z["out_estimate"])) <= 1e-6 * max(1.0, abs(float(z["out_estimate"])))
    assert abs(float(alpha) - float(z["out_mean_accept"])) <= 1e-6


def _load_vae(name):
    weight_checksum = U.vae_weight_checksum
    z = np.load(os.path.join(GOLD, "ref", name + ".npz"))
    meta = _meta(z)
    P = U.VaeProblem(**meta["kw"])
    if meta["weights_stored"]:
        P.mask = z["mask"]
        P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
        P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
        P.dec_W = [z["decW_%d" % i] for i in range(len(P.dec_W))]
        P.dec_b = [z["decb_%d" % i] for i in range(len(P.dec_b))]
        if P.use_encoder:
            P.enc_W = [z["encW_%d" % i] for i in range(len(P.enc_W))]
            P.enc_b = [z["encb_%d" % i] for i in range(len(P.enc_b))]
    else:
        assert np.array_equal(P.mask, z["mask"])
    assert abs(weight_checksum(P) - meta["weight_checksum"]) <= 1e-9 * abs(meta["weight_checksum"]), \
        "regenerated VAE weights differ from the ones the fixture was made with"
    d = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    return P, d, z


@pytest.mark.parametrize("name", ["c5_vae_mini_n96", "c5_vae_full_n32"])
def test_oracle_reproduces_reference_vae_target(name):
    """BASELINE config 5: decoder-Bernoulli posterior + aux-conditioned nets.  c5_vae_full is mnist_vae.py's own text
    (:104-111, 122-126, 130-178) at its own layer sizes; the energy there is O(500), so fp64 agreement is ~1e-12."""
    P, d, z = _load_vae(name)
    got = U.run_oracle_propose(P, d, torch.float64)
    for k in ("Lx", "Lv", "px", "x_next"):
        assert np.max(np.abs(got[k] - z["out_" + k])) <= 1e-9 * max(1.0, np.abs(z["out_" + k]).max()), (name, k)
    dyn, _ = P.oracle_for(d, torch.float64)
    assert np.max(np.abs(dyn.energy(U.t64(d["x"])).numpy() - z["out_energy"])) <= 1e-9 * np.abs(z["out_energy"]).max()
    assert np.max(np.abs(dyn.grad_energy(U.t64(d["x"])).numpy() - z["out_grad_energy"])) <= 1e-9 * np.abs(z["out_grad_energy"]).max()


# ---- with the reference present: reproducibility and the stand-in's semantics ------------------------------
@have_reference
@pytest.mark.parametrize("path", [f for f in PROPOSE_FILES if "c1_scg2_n200_stress" in f or "c3_mog2" in f],
                         ids=lambda f: os.path.basename(f)[:-4])
def test_fixtures_regenerate_from_the_reference(path):
    import golden_io
    import ref_runner as R
    P, d, ref = golden_io.load(path)
    again = R.run_propose(P, d, "float64", P.meta["log_jac"])
    for k in ("Lx", "Lv", "px", "x_next"):
        assert np.array_equal(again[k], ref[k]), k


@have_reference
def test_reference_sources_are_loaded_unmodified_from_their_path():
    ref = ref_loader.load()
    assert ref.dynamics.__file__ == os.path.join(ref_loader.REF_ROOT, "utils", "dynamics.py")
    assert ref.sampler.__file__.endswith("utils/sampler.py")
    assert ref.dynamics.Dynamics.__module__ == "dynamics"
    assert hasattr(ref.tf, "shim")  # the stand-in, not a real TensorFlow
    # the notebook's network cell is executed verbatim: width 10, three heads
    import ref_runner as R
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    _, dyn = R.build_dynamics(P, "float32")
    assert sorted(ref.tf.shim.variables)[:3] == ["VNet/embed_1/W", "VNet/embed_1/b", "VNet/embed_2/W"]
    assert "alpha" in ref.tf.shim.variables and len(ref.tf.shim.variables) == 33
    assert dyn.XNet.layers[0].layers[0].W.shape == (2, 10)


@have_reference
def test_stand_in_follows_tf1_semantics():
    tf = ref_loader.load().tf
    tf.shim.reset()
    tf.shim.set_real("float32")
    x = tf.constant(np.arange(6, dtype=np.float32).reshape(3, 2))
    # rank-1 condition selects rows (utils/sampler.py:55)
    w = tf.where(tf.constant(np.array([True, False, True])), x, tf.zeros_like(x)).numpy()
    assert np.array_equal(w, [[0, 1], [0, 0], [4, 5]])
    # python / numpy operands take the tensor dtype (utils/distributions.py:127: fp32 x - float64 mus)
    assert (x - np.array([1.0, 2.0])).dtype == torch.float32 and (2 * np.pi * x).dtype == torch.float32
    # diag_part(matmul) of quadratic_gaussian (utils/distributions.py:31-32)
    S = np.array([[2.0, 0.5], [0.5, 1.0]], np.float32)
    q = tf.diag_part(0.5 * tf.matmul(tf.matmul(x, S), tf.transpose(x))).numpy()
    assert np.allclose(q, 0.5 * np.einsum("ni,ij,nj->n", x.numpy(), S, x.numpy()))
    # tf.gradients sums the outputs and is per-row for row-wise energies (utils/dynamics.py:217-218)
    xi = tf.shim.input(x.numpy())
    g = tf.gradients(tf.reduce_sum(tf.square(xi), 1), xi)[0].numpy()
    assert np.allclose(g, 2 * x.numpy())
    # while_loop with a float counter (utils/dynamics.py:253-267)
    out = tf.while_loop(lambda a, t: tf.less(t, 3), lambda a, t: (a + t, t + 1), [tf.constant(0.), tf.constant(0.)])
    assert float(out[0]) == 3.0 and float(out[1]) == 3.0
    # injected randomness is consumed in call order and shape-checked
    tf.shim.feed_random(normal=[np.ones((2, 2))], uniform=[np.array([[1], [0]])])
    assert np.array_equal(tf.random_uniform((2, 1), maxval=2, dtype=tf.int32).numpy(), [[1], [0]])
    with pytest.raises(ValueError):
        tf.random_normal((3, 2))
    tf.shim.reset()
    with pytest.raises(RuntimeError):
        tf.random_normal((2, 2))
    # variables: scoped names, preloaded values, duplicate names refused like TF1 without reuse
    tf.shim.preload({"a/W": np.full((2, 3), 7.0)})
    with tf.variable_scope("a"):
        W = tf.get_variable("W", shape=(2, 3), initializer=tf.constant_initializer(0.))
        b = tf.get_variable("b", shape=(3,), initializer=tf.constant_initializer(0.))
        with pytest.raises(ValueError):
            tf.get_variable("W", shape=(2, 3), initializer=tf.constant_initializer(0.))
    assert W.name == "a/W:0" and float(W.numpy()[0, 0]) == 7.0 and float(b.numpy().sum()) == 0.0
    # the fp64 mode keeps python constants fp32-rounded (the fp32 graph in wider arithmetic)
    tf.shim.set_real("float64")
    y = tf.constant(np.ones(2, np.float32), dtype=tf.float32) * 0.1
    assert y.dtype == torch.float64 and float(y.numpy()[0]) == float(np.float32(0.1))
    tf.shim.set_real("float32")


@have_reference
def test_linear_initialiser_matches_the_product_layers():
    """utils/layers.py:29-37 on the stand-in's restatement of tf.contrib's variance_scaling_initializer: the truncated
    normal's scale sqrt(1.3 * 2 * factor / fan_in) is what l2hmc_b200.layers.Linear draws from."""
    ref = ref_loader.load()
    tf = ref.tf
    tf.shim.reset(seed=3)
    tf.shim.set_real("float32")
    lin = ref.layers.Linear(400, 300, scope="probe", factor=0.5)
    W = lin.W.numpy()
    std = np.sqrt(1.3 * 2 * 0.5 / 400)
    assert np.abs(W).max() <= 2 * std * (1 + 1e-6) and abs(W.std() / (std * 0.8796) - 1) < 0.02  # truncation shrinks std by 0.8796
    assert float(np.abs(lin.b.numpy()).max()) == 0.0
    from l2hmc_b200.layers import Linear
    torch.manual_seed(0)
    Wp = np.asarray(Linear(400, 300, scope="probe", factor=0.5).W)
    assert np.abs(Wp).max() <= 2 * std * (1 + 1e-6) and abs(Wp.std() / W.std() - 1) < 0.03
