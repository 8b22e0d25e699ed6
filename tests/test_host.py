"""CPU tests of the host side: Philox twin, layer descriptions, mask init, the C-ABI library's exports
(no compute calls without a GPU), argument validation that happens before any CUDA call."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import util as U
from l2hmc_b200 import _lib, layers, philox
from l2hmc_b200.dynamics import Dynamics, init_mask


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def kat(c, k):
        c = [np.array([x], dtype=np.uint32) for x in c]
        return [int(x[0]) for x in philox.philox4x32_10(*c, k[0], k[1])]
    assert kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_streams_statistics_and_sharding():
    v = philox.normals(3, 5, 20000, 6)
    assert abs(v.mean()) < 0.02 and abs(v.std() - 1) < 0.02
    d, u = philox.direction_and_uniform(3, 5, 20000)
    assert abs(d.mean() - 0.5) < 0.02 and 0 <= u.min() and u.max() < 1
    # keyed by global chain id: a shard regenerates the same numbers
    v2 = philox.normals(3, 5, 100, 6, chain_offset=700)
    assert np.array_equal(v2, v[700:800])
    d2, u2 = philox.direction_and_uniform(3, 5, 100, chain_offset=700)
    assert np.array_equal(d2, d[700:800]) and np.array_equal(u2, u[700:800])
    # different call counters decorrelate
    assert not np.array_equal(philox.normals(3, 6, 100, 6), v[:100])


def test_init_mask_follows_reference():
    """floor(D/2) ones per step (utils/dynamics.py:84-93)."""
    for D in (2, 50, 32, 5):
        m = init_mask(D, 7, np.random.default_rng(0))
        assert m.shape == (7, D) and m.dtype == np.float32
        assert set(np.unique(m)) <= {0.0, 1.0}
        assert (m.sum(1) == int(D / 2)).all()


def test_linear_init_is_tf_variance_scaling():
    """variance_scaling_initializer(factor=2f, FAN_IN, normal): truncated normal, std sqrt(1.3*2f/fan_in)."""
    layers.manual_seed(0)
    l = layers.Linear(400, 300, factor=0.5)
    std = np.sqrt(1.3 * 2 * 0.5 / 400)
    w = l.W.numpy()
    assert np.abs(w).max() <= 2 * std + 1e-7
    assert abs(w.std() - 0.88 * std) < 0.03 * std  # a +-2 sigma truncated normal has 0.88 of the std
    assert float(l.b.abs().max()) == 0.0


def test_compile_net_roundtrip_and_rejects_other_structures():
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    net = P.net_factory()(2, "XNet", 2.0)
    got = layers.compile_stq_net(net, 2)
    for k, v in P.xnet.items():
        assert np.array_equal(got[k], v), k
    # the description is also callable, like the reference's layer objects
    a = torch.randn(5, 2)
    S, T, Q = net([a, a, torch.ones(5, 2), None])
    So, To, Qo = U.O.net_apply(U.O.net_cast(P.xnet, torch.float32), a, a, torch.ones(5, 2))
    assert torch.allclose(S, So, atol=1e-6) and torch.allclose(T, To, atol=1e-6) and torch.allclose(Q, Qo, atol=1e-6)
    with pytest.raises(layers.NetStructureError):
        layers.compile_stq_net(layers.Sequential([layers.Linear(2, 2)]), 2)
    bad = P.net_factory()(2, "XNet", 2.0)
    bad.layers[0].layers[3] = layers.Linear(784, 10)  # non-zero aux branch
    with pytest.raises(layers.NetStructureError):
        layers.compile_stq_net(bad, 2)


def test_dynamics_host_state_and_loud_failure_without_gpu():
    P = U.Problem(**U.CONFIGS["c1_scg2"])
    d = Dynamics(2, P.dist.get_energy_function(), T=10, eps=0.1, net_factory=P.net_factory())
    assert d.mask.shape == (10, 2) and d.width == 10 and abs(d.eps - 0.1) < 1e-7 and not d.hmc
    d.mask = P.mask
    with pytest.raises(ValueError):
        d.mask = np.zeros((3, 2), np.float32)
    with pytest.raises(TypeError):
        Dynamics(2, lambda x: x.sum(1), T=10, eps=0.1, net_factory=P.net_factory())  # opaque callable
    with pytest.raises(ValueError):
        Dynamics(3, P.dist.get_energy_function(), T=10, eps=0.1, net_factory=P.net_factory())
    if not torch.cuda.is_available():
        with pytest.raises(_lib.L2HMCLibraryError):
            d.forward(torch.zeros(4, 2))


def test_library_builds_loads_and_exports_every_declared_symbol():
    """nvcc cross-compiles for sm_100a without a GPU; every function include/l2hmc.h declares must be
    exported by libl2hmc.so and bound by the ctypes layer."""
    _lib.build()
    lib = _lib.load()
    hdr = open(os.path.join(U.ROOT, "include", "l2hmc.h")).read()
    declared = set(re.findall(r"\b(l2hmc_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"l2hmc_ctx"}
    bound = {name for name, _, _ in _lib.EXPORTS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.l2hmc_version()


def test_library_reports_errors_without_exceptions():
    lib = _lib.load()
    ctx = C.c_void_p()
    cfg = _lib.Config(0, 10, 10, 0, 0, 0, 0.1)  # x_dim = 0
    assert lib.l2hmc_create(C.byref(cfg), C.byref(ctx)) == 1  # L2HMC_EINVAL
    assert b"x_dim" in lib.l2hmc_last_error(None)
    if not torch.cuda.is_available():
        cfg = _lib.Config(2, 10, 10, 0, 0, 0, 0.1)
        assert lib.l2hmc_create(C.byref(cfg), C.byref(ctx)) == 2  # L2HMC_ECUDA: no device, no fallback
        assert b"no CUDA device" in lib.l2hmc_last_error(None)


def test_sass_is_sm100a_only():
    """The shipped library holds sm_100a code and nothing else (no multi-arch fallback)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_shard_bounds_partition():
    from l2hmc_b200.sharding import shard_bounds
    for n in (0, 1, 7, 64, 1 << 18, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in b]
            assert max(sizes) - min(sizes) <= 1


def test_ais_mixed_gaussian_is_the_annealed_energy_up_to_a_constant():
    """ais._mixed_gaussian: (1 - beta) U0 + beta U1 of two Gaussian energies is one Gaussian energy plus a constant
    (utils/ais.py:44-45); checked against the oracle's MixedEnergy on random points (energy differences and gradients)."""
    import torch
    import util as U
    from l2hmc_b200.ais import _mixed_gaussian
    from l2hmc_b200.distributions import Gaussian
    rng = np.random.default_rng(0)
    D = 5
    A = rng.standard_normal((D, D))
    g0 = Gaussian(rng.standard_normal(D), np.eye(D) * 1.5)
    g1 = Gaussian(rng.standard_normal(D), A @ A.T / D + 0.2 * np.eye(D))
    e0, e1 = g0.get_energy_function(), g1.get_energy_function()
    x = torch.as_tensor(rng.standard_normal((64, D)))
    for beta in (0.0, 0.3, 1.0):
        m = _mixed_gaussian(e0, e1, beta)
        mine = U.O.GaussianEnergy(m.mu[0], m.S[0], torch.float64)
        ref = U.O.MixedEnergy(U.O.GaussianEnergy(e0.mu[0], e0.S[0], torch.float64), U.O.GaussianEnergy(e1.mu[0], e1.S[0], torch.float64), beta)
        d_mine = mine.energy(x) - mine.energy(x[:1])
        d_ref = ref.energy(x) - ref.energy(x[:1])
        assert float((d_mine - d_ref).abs().max()) < 1e-5 * max(1.0, float(d_ref.abs().max()))
        assert float((mine.grad(x) - ref.grad(x)).abs().max()) < 1e-5 * max(1.0, float(ref.grad(x).abs().max()))


def test_training_adam_is_tensorflow_adam_with_the_notebook_schedule():
    """training.Adam (tf.train.AdamOptimizer + exponential_decay(1e-3, step, 1000, 0.96, staircase),
    SCGExperiment.ipynb:183-186) against torch.optim.Adam on the same gradients; parameters flow back into the layer
    objects and alpha = log(eps) is trained with them (utils/dynamics.py:50-58).  Host logic only: no GPU involved."""
    from l2hmc_b200 import training
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    dyn = P.product()
    opt = training.Adam(dyn)
    assert opt.learning_rate == pytest.approx(1e-3)
    opt.global_step = 2500
    assert opt.learning_rate == pytest.approx(1e-3 * 0.96 ** 2)
    opt.global_step = 0
    rng = np.random.default_rng(0)
    ref_p = {(i, k): torch.tensor(np.asarray(dyn._net_params[i][k], np.float32)).clone().requires_grad_(True)
             for i in range(2) for k in training.NAMES}
    ref_alpha = torch.tensor([float(dyn.alpha)], requires_grad=True)
    topt = torch.optim.Adam(list(ref_p.values()) + [ref_alpha], lr=1e-3, eps=1e-8)
    for step in range(3):
        grads = {"loss": torch.zeros(1), "eps": torch.tensor([float(rng.standard_normal())], dtype=torch.float32)}
        for i, key in enumerate(("XNet", "VNet")):
            grads[key] = {k: torch.tensor(rng.standard_normal(np.shape(dyn._net_params[i][k])).astype(np.float32))
                          for k in training.NAMES}
        eps_now = dyn.eps
        for (i, k), p in ref_p.items():
            p.grad = grads[("XNet", "VNet")[i]][k].reshape(p.shape).clone()
        ref_alpha.grad = grads["eps"] * eps_now
        topt.step()
        opt.apply(dyn, grads)
    assert opt.global_step == 3
    for (i, k), p in ref_p.items():
        got = np.asarray(dyn._net_params[i][k], np.float32).reshape(p.shape)
        assert np.allclose(got, p.detach().numpy(), rtol=0, atol=2e-6), (i, k)
    assert np.log(dyn.eps) == pytest.approx(float(ref_alpha.detach()[0]), abs=2e-6)
    # the layer objects hold the new weights (what a checkpoint of the nets would save)
    assert np.allclose(np.asarray(dyn.XNet.layers[3].W), np.asarray(dyn._net_params[0]["W4"]))


def test_losses_module_is_utils_losses():
    """l2hmc_b200.losses (utils/losses.py:26-59, values) against the oracle's restatement, and get_loss's table."""
    from l2hmc_b200 import losses
    rng = np.random.default_rng(1)
    x, X = torch.as_tensor(rng.standard_normal((9, 3))), torch.as_tensor(rng.standard_normal((9, 3)))
    p = torch.as_tensor(rng.random(9))
    # the oracle's fp64 twin keeps the fp32-rounded 1e-4 / scale of the reference's fp32 graph: equal to ~1e-11 here
    assert torch.allclose(losses.loss_vec(x, X, p), U.O.loss_vec(x, X, p), rtol=0, atol=1e-10)
    for name, ref in (("mixed", lambda: U.O.loss_mixed(x, X, p, 1.0)), ("standard", lambda: U.O.loss_std(x, X, p)),
                      ("inverse", lambda: U.O.loss_inverse(x, X, p)), ("logsumexp", lambda: U.O.loss_logsumexp(x, X, p))):
        assert float(losses.get_loss(name)(x, X, p)) == pytest.approx(float(ref()), rel=1e-7)
    assert float(losses.loss_mixed(x, X, p, scale=0.1)) == pytest.approx(float(U.O.loss_mixed(x, X, p, 0.1)), rel=1e-7)
    with pytest.raises(KeyError):
        losses.get_loss("nope")
