"""GPU tests of the training path (l2hmc_loss_grad; l2hmc_b200/training.py): gradient parity with the hand-written reverse
sweep of the oracle and with tf.gradients of the reference's own loss cell (tests/golden/ref/notebook_loss_*.npz), every
loss of utils/losses.py, the optimiser loop, data-parallel sharding of the internal draws."""
import json
import os
import numpy as np
import pytest
import torch

import util as U
import l2hmc_reverse as R
from l2hmc_b200 import propose, training

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ORACLE_NETS = (("XNet", "x"), ("VNet", "v"))


def _setup(name, n, seed=4, **over):
    kw = dict(U.CONFIGS[name])
    kw.update(over)
    P = U.Problem(regime="stress", **kw)
    rng = np.random.default_rng(seed)
    x = P.x0(n, rng)
    d = rng.integers(0, 2, n).astype(np.uint8)
    v = rng.standard_normal((n, P.D)).astype(np.float32)
    return P, x, d, v


def _oracle(P, x, d, v, scale, temperature=1.0):
    dyn = P.oracle(torch.float64, temperature=temperature)
    r = {"direction": torch.as_tensor(d.astype(np.float64)), "v_f": torch.as_tensor(v).double(), "v_b": torch.as_tensor(v).double()}
    with torch.no_grad():
        acc = R._Acc(dyn)
        loss = R.loss_and_grads(torch.as_tensor(x).double(), dyn, r, scale, acc)
    return float(loss), acc


def _worst(grads, acc):
    worst = 0.0
    for key, attr in ORACLE_NETS:
        ref = getattr(acc, attr)
        for k in training.NAMES:
            a = grads[key][k].double().cpu().numpy()
            b = ref[k].numpy().reshape(a.shape)
            assert np.isfinite(a).all(), (key, k)
            worst = max(worst, float(np.abs(a - b).max() / max(1e-12, np.abs(b).max())))
    return worst


# tolerances: the fp32 noise floor of this gradient at exactly these settings is 0.4 .. 2.1e-5 (fp32 vs fp64 run of the
# oracle sweep)
@pytest.mark.parametrize("name,n,tol", [("c1_scg2", 200, 5e-4), ("c2_scg50", 256, 1e-3), ("c3_mog2", 200, 1e-3),
                                        ("c4_rw32", 128, 1e-3), ("funnel3", 128, 1e-3)])
def test_loss_grad_matches_the_hand_written_reverse_pass(name, n, tol):
    P, x, d, v = _setup(name, n)
    dyn = P.product()
    rng = {"direction": torch.as_tensor(d, device=DEV), "v": torch.as_tensor(v, device=DEV)}
    loss, grads, Lx, px = training.loss_and_grads(dyn, torch.as_tensor(x, device=DEV), rng=rng, scale=0.1)
    loss_o, acc = _oracle(P, x, d, v, 0.1)
    assert float(loss[0]) == pytest.approx(loss_o, rel=1e-3)
    assert _worst(grads, acc) < tol
    assert float(grads["eps"][0]) == pytest.approx(float(acc.eps), rel=5e-3, abs=1e-3 * abs(loss_o))
    assert float(grads["alpha"][0]) == pytest.approx(float(acc.eps) * P.eps, rel=5e-3, abs=1e-4 * abs(loss_o))
    # the forward sweep is the sampling transition: same Lx / px as propose with the same direction and momentum
    Lx2, _, px2, _ = propose(torch.as_tensor(x, device=DEV), dyn, rng=rng)
    assert U.max_rel(Lx.cpu().numpy(), Lx2.cpu().numpy()) < 2e-5
    assert float((px - px2).abs().max()) < 2e-4   # two fp32 evaluations of H0 - H1 + log|J| (O(100) at 50 dimensions)


def test_temperature_enters_the_gradient():
    P, x, d, v = _setup("c1_scg2", 64)
    dyn = P.product(use_temperature=True)
    dyn.temperature = 1.7
    rng = {"direction": torch.as_tensor(d, device=DEV), "v": torch.as_tensor(v, device=DEV)}
    loss, grads, _, _ = training.loss_and_grads(dyn, torch.as_tensor(x, device=DEV), rng=rng)
    loss_o, acc = _oracle(P, x, d, v, 0.1, temperature=1.7)
    assert float(loss[0]) == pytest.approx(loss_o, rel=1e-3)
    assert _worst(grads, acc) < 5e-4


def test_the_two_batches_of_the_notebook_objective_add_up():
    P, x, d, v = _setup("c1_scg2", 96)
    rng2 = np.random.default_rng(8)
    z = rng2.standard_normal((96, P.D)).astype(np.float32)
    d2, v2 = rng2.integers(0, 2, 96).astype(np.uint8), rng2.standard_normal((96, P.D)).astype(np.float32)
    dyn = P.product()
    t = lambda a: torch.as_tensor(a, device=DEV)  # noqa: E731
    loss, grads, _, _ = training.notebook_loss_and_grads(dyn, t(x), t(z), rng_x={"direction": t(d), "v": t(v)},
                                                         rng_z={"direction": t(d2), "v": t(v2)})
    odyn = P.oracle(torch.float64)
    rx = {"direction": torch.as_tensor(d.astype(np.float64)), "v_f": torch.as_tensor(v).double(), "v_b": torch.as_tensor(v).double()}
    rz = {"direction": torch.as_tensor(d2.astype(np.float64)), "v_f": torch.as_tensor(v2).double(), "v_b": torch.as_tensor(v2).double()}
    loss_o, g = R.notebook_loss_and_grads(torch.as_tensor(x), torch.as_tensor(z), odyn, rx, rz)
    assert float(loss[0]) == pytest.approx(float(loss_o), rel=1e-3)

    class A:  # the oracle returns plain dicts
        x, v, eps = g["xnet"], g["vnet"], g["eps"]
    assert _worst(grads, A) < 5e-4


def test_unsupported_targets_and_modes_raise():
    from l2hmc_b200 import _lib
    H = U.Problem(hmc=True, **U.CONFIGS["c1_scg2"]).product()
    x = torch.as_tensor(np.random.default_rng(0).standard_normal((8, 2)).astype(np.float32), device=DEV)
    with pytest.raises(ValueError):
        training.loss_and_grads(H, x)
    V = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])          # decoder energy / aux-conditioned nets: not covered yet
    d = V.draws(8)
    dyn = V.product()
    with pytest.raises(_lib.L2HMCError):
        training.loss_and_grads(dyn, torch.as_tensor(d["x"], device=DEV))


def test_training_loop_lowers_the_loss_and_moves_every_parameter():
    """SCGExperiment.ipynb:254-270 in miniature: Adam with the notebook's schedule on both nets and alpha."""
    P, x, d, v = _setup("c1_scg2", 200)
    dyn = P.product()
    before = [{k: np.array(p[k]) for k in training.NAMES} for p in dyn._net_params]
    eps0 = dyn.eps
    t = lambda a: torch.as_tensor(a, device=DEV)  # noqa: E731
    fixed = dict(rng_x={"direction": t(d), "v": t(v)}, rng_z={"direction": t(d), "v": t(v)})
    z = t(np.random.default_rng(3).standard_normal((200, P.D)).astype(np.float32))
    first, _, _, _ = training.notebook_loss_and_grads(dyn, t(x), z, **fixed)
    opt = training.Adam(dyn)
    u = t(np.random.default_rng(5).random(200).astype(np.float32))
    for it in range(12):   # fixed batch and fixed randomness: plain Adam descent on one objective
        out = training.train_step(dyn, opt, t(x), z=z, u=u, **fixed)
        assert np.isfinite(out["loss"])
    last, _, _, _ = training.notebook_loss_and_grads(dyn, t(x), z, **fixed)
    assert float(last[0]) < float(first[0])
    samples = out["samples"]
    for it in range(3):    # the notebook's loop proper: fresh noise, Metropolis output fed back (:254-270)
        out = training.train_step(dyn, opt, samples)
        samples = out["samples"]
        assert np.isfinite(out["loss"]) and samples.shape == (200, P.D)
    assert dyn.eps != eps0
    for p0, p1 in zip(before, dyn._net_params):
        for k in training.NAMES:
            assert not np.array_equal(p0[k], np.asarray(p1[k])), k


@pytest.mark.parametrize("kind", ["standard", "inverse", "logsumexp"])
def test_library_losses_match_the_hand_written_reverse_pass(kind):
    """get_loss(name) of utils/losses.py:26-59."""
    P, x, d, v = _setup("c1_scg2", 600)
    dyn = P.product()
    rng = {"direction": torch.as_tensor(d, device=DEV), "v": torch.as_tensor(v, device=DEV)}
    loss, grads, _, _ = training.loss_and_grads(dyn, torch.as_tensor(x, device=DEV), rng=rng, loss=kind)
    odyn = P.oracle(torch.float64)
    r = {"direction": torch.as_tensor(d.astype(np.float64)), "v_f": torch.as_tensor(v).double(), "v_b": torch.as_tensor(v).double()}
    with torch.no_grad():
        acc = R._Acc(odyn)
        loss_o = R.loss_and_grads(torch.as_tensor(x).double(), odyn, r, 0.1, acc, kind=kind)
    assert float(loss[0]) == pytest.approx(float(loss_o), rel=1e-3)
    assert _worst(grads, acc) < 1e-3


def test_loss_value_agrees_with_the_losses_module():
    """The value l2hmc_loss_grad returns is utils/losses.py's on the proposals it returns."""
    from l2hmc_b200 import losses
    P, x, d, v = _setup("c1_scg2", 300)
    dyn = P.product()
    xt = torch.as_tensor(x, device=DEV)
    rng = {"direction": torch.as_tensor(d, device=DEV), "v": torch.as_tensor(v, device=DEV)}
    for name in ("mixed", "standard", "inverse", "logsumexp"):
        loss, _, Lx, px = training.loss_and_grads(dyn, xt, rng=rng, scale=0.1, loss=name)
        ref = losses.loss_mixed(xt, Lx, px, scale=0.1) if name == "mixed" else losses.get_loss(name)(xt, Lx, px)
        assert float(loss[0]) == pytest.approx(float(ref), rel=1e-4), name


@pytest.mark.parametrize("name", ["notebook_loss_c1_n200", "notebook_loss_c3_n64"])
def test_training_gradients_match_the_reference(name):
    """The notebook objective (SCGExperiment.ipynb:159-181) and tf.gradients of it w.r.t. every variable of both nets and
    alpha, as the UNMODIFIED reference computes them on oracle/tf_shim (tests/golden/make_ref_golden.py), against
    l2hmc_loss_grad on the same parameters and draws."""
    import ref_runner  # variable-name table only (oracle/: test infrastructure)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref", name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    P = U.Problem(regime=meta["regime"], **meta["kw"])
    P.mask = z["mask"]
    P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
    P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
    dyn = P.product()
    t = lambda a: torch.as_tensor(np.asarray(a), device=DEV)  # noqa: E731

    def sel(pre):
        d = z[pre + "dir"]
        return {"direction": t(d), "v": t(np.where(d[:, None] != 0, z[pre + "v_f"], z[pre + "v_b"]).astype(np.float32))}
    loss, grads, Lx, px = training.notebook_loss_and_grads(dyn, t(z["in_x"]), t(z["in_z"]), rng_x=sel("in_rx_"),
                                                           rng_z=sel("in_rz_"), scale=meta["scale"])
    ref_loss = float(z["out_loss"])
    assert float(loss[0]) == pytest.approx(ref_loss, rel=1e-3)
    worst = 0.0
    for scope in ("XNet", "VNet"):
        for k in training.NAMES:
            ref_g = z["out_grad__%s__%s" % (scope, ref_runner.NET_VARS[k].replace("/", "__"))]
            a = grads[scope][k].double().cpu().numpy().reshape(ref_g.shape)
            worst = max(worst, float(np.abs(a - ref_g).max() / max(1e-12, np.abs(ref_g).max())))
    assert worst < 1e-3, worst   # fp32 noise floor of this gradient: 0.4 .. 2e-5
    ga = float(np.asarray(z["out_grad__alpha"]).reshape(-1)[0])
    assert float(grads["alpha"][0]) == pytest.approx(ga, rel=5e-3, abs=1e-4 * abs(ref_loss))
    assert float(np.max(np.abs(px.cpu().numpy() - z["out_px"]))) <= 2e-4


def test_sharded_internal_draws_equal_the_single_rank_draws():
    """Data-parallel training draws direction bits, momenta, z noise and accept uniforms from Philox keyed by the GLOBAL
    chain id: two shards with chain_offset and count = N_global add up to the single-rank gradient, and train_step's
    Metropolis output is the same chain by chain."""
    P, x, _, _ = _setup("c1_scg2", 200)
    xt = torch.as_tensor(x, device=DEV)
    dyn = P.product(seed=11)
    dyn._ensure_ctx()
    c0 = dyn._counter
    loss, g, Lx, px = training.loss_and_grads(dyn, xt)
    parts = []
    for lo, hi in ((0, 90), (90, 200)):
        dyn._counter = c0                       # every rank holds the same Dynamics (seed and call counter)
        parts.append(training.loss_and_grads(dyn, xt[lo:hi].contiguous(), count=200, chain_offset=lo))
    assert torch.equal(torch.cat([p[2] for p in parts]), Lx) and torch.equal(torch.cat([p[3] for p in parts]), px)
    assert float(parts[0][0][0] + parts[1][0][0]) == pytest.approx(float(loss[0]), rel=1e-5)
    for key in ("XNet", "VNet"):
        for k in training.NAMES:
            s = parts[0][1][key][k] + parts[1][1][key][k]
            assert float((s - g[key][k]).abs().max()) <= 1e-4 * max(1e-12, float(g[key][k].abs().max())), (key, k)
    # without the offset the second shard would repeat the first shard's draws
    dyn._counter = c0
    wrong = training.loss_and_grads(dyn, xt[90:200].contiguous(), count=200)
    assert not torch.equal(wrong[2], Lx[90:200])


def test_launch_sequence_gradients_are_bit_reproducible():
    """The sums over chains (weight-gradient products, bias column sums, loss) are split over CTAs and added in part order
    (train.cuh: k_reduce_add), no atomics: the same inputs give the same bits, with more chains than one part holds."""
    P, x, d, v = _setup("c2_scg50", 1000)
    rng = {"direction": torch.as_tensor(d, device=DEV), "v": torch.as_tensor(v, device=DEV)}
    xt = torch.as_tensor(x, device=DEV)
    dyn = P.product()
    a = training.loss_and_grads(dyn, xt, rng=rng, scale=0.1)
    b = training.loss_and_grads(dyn, xt, rng=rng, scale=0.1)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    for key in ("XNet", "VNet"):
        for k in training.NAMES:
            assert torch.equal(a[1][key][k], b[1][key][k]), (key, k)
    assert torch.equal(a[1]["eps"], b[1]["eps"])


@pytest.mark.parametrize("name,n", [("c1_scg2", 200), ("c3_mog2", 333), ("funnel3", 130)])
def test_fused_small_net_training_kernel_is_one_launch_and_equals_the_launch_sequence(name, n, monkeypatch):
    """x_dim <= 4 / width <= 16 (the notebook's nets): l2hmc_loss_grad is ONE launch of small_train_kernel (one chain per
    thread, tape in local memory, gradients reduced once per block); it must agree with the launch-sequence path
    (L2HMC_TRAIN_FUSED=0: ~200 launches per leapfrog step) and with the oracle's hand-written reverse sweep."""
    P, x, d, v = _setup(name, n)
    rng = {"direction": torch.as_tensor(d, device=DEV), "v": torch.as_tensor(v, device=DEV)}
    xt = torch.as_tensor(x, device=DEV)
    monkeypatch.delenv("L2HMC_TRAIN_FUSED", raising=False)
    dyn = P.product()
    dyn._ensure_ctx()
    l0 = dyn.launch_count
    loss, grads, Lx, px = training.loss_and_grads(dyn, xt, rng=rng, scale=0.1)
    assert dyn.launch_count - l0 == 1
    monkeypatch.setenv("L2HMC_TRAIN_FUSED", "0")
    dyn2 = P.product()
    dyn2._ensure_ctx()
    l0 = dyn2.launch_count
    loss2, grads2, Lx2, px2 = training.loss_and_grads(dyn2, xt, rng=rng, scale=0.1)
    assert dyn2.launch_count - l0 > 100
    monkeypatch.delenv("L2HMC_TRAIN_FUSED", raising=False)
    assert float(loss[0]) == pytest.approx(float(loss2[0]), rel=2e-4)
    assert U.max_rel(Lx.cpu().numpy(), Lx2.cpu().numpy()) < 2e-5 and float((px - px2).abs().max()) < 2e-4
    loss_o, acc = _oracle(P, x, d, v, 0.1)
    assert float(loss[0]) == pytest.approx(loss_o, rel=1e-3)
    assert _worst(grads, acc) < 1e-3
    assert float(grads["eps"][0]) == pytest.approx(float(acc.eps), rel=5e-3, abs=1e-3 * abs(loss_o))
    for key in ("XNet", "VNet"):
        for k in training.NAMES:
            a, b = grads[key][k], grads2[key][k]
            assert float((a - b).abs().max()) <= 2e-3 * max(1e-12, float(b.abs().max())), (key, k)
    # 'standard' is separable too; 'inverse' needs a batch statistic and stays on the launch sequence
    l0 = dyn.launch_count
    training.loss_and_grads(dyn, xt, rng=rng, loss="standard")
    assert dyn.launch_count - l0 == 1
    l0 = dyn.launch_count
    training.loss_and_grads(dyn, xt, rng=rng, loss="inverse")
    assert dyn.launch_count - l0 > 100
