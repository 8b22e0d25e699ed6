"""GPU parity: the CUDA path (through the Python mirror -> ctypes -> C ABI) against the CPU oracle.

Tolerances (written here, per the task statement; north_star: "within 1e-5 relative fp32 tolerance"):
  * per chain -- samples Lx, Lv as max|kernel - fp64 oracle| / max(1, max|ref|), accept probability absolute:
    samples within O32 + 1e-5, accept probability within 4 * O32 + 1e-5 (see _check), O32 being the error the fp32
    twin of the reference-pinned oracle makes against its fp64 twin on the same inputs (the fp32 path is only defined
    up to summation order);
  * mean accept probability over chains within 1e-5 of the fp64 oracle (BASELINE.json "accept-prob delta");
  * Metropolis decisions identical except where |p - u| is inside that noise (<= 1e-4).
Tests of derived properties (round trips, sub-sampled full-size runs, fused multi-transition chains) state their own,
wider figures next to the assertion.
"""
import numpy as np
import pytest
import torch

import util as U

pytestmark = pytest.mark.gpu

SAMPLE_TOL = 2e-5
P_TOL = 5e-5
P_MEAN_TOL = 1e-5
NORTH_STAR_TOL = 1e-5


def _check(rep):
    """north_star: within 1e-5 of the reference on identical seeds.  O32 = the error of the fp32 twin of the
    (reference-pinned) oracle against its fp64 twin on the same inputs, i.e. the rounding noise the fp32 path carries
    whatever the implementation:
      * samples: within O32 + 1e-5;
      * mean accept probability: within 1e-5, flat (or O32's own mean error where that is larger: chaotic targets);
      * per-chain accept probability: within 4 * O32 + 1e-5.  It is the MAXIMUM over a few hundred chains of a
        cancellation error (Hamiltonians of O(100) subtracted in fp32); two evaluation orders of the same arithmetic
        (Eigen, torch, the kernels) differ by up to ~4x in that maximum while agreeing to 1e-6 in the mean
        (profiles/r02_parity_noise.txt lists kernel vs O32 for every configuration and engine)."""
    assert rep["Lx_kernel"] <= rep["Lx_o32"] + NORTH_STAR_TOL, rep
    assert rep["Lv_kernel"] <= rep["Lv_o32"] + NORTH_STAR_TOL, rep
    assert rep["px_kernel"] <= 4 * rep["px_o32"] + NORTH_STAR_TOL, rep
    assert rep["px_mean_kernel"] <= max(NORTH_STAR_TOL, rep["px_mean_o32"]), rep
    assert rep["accept_flips_outside_noise"] == 0, rep


@pytest.mark.parametrize("name,n,regime", [
    ("c1_scg2", 200, "init"),          # BASELINE config 1 exactly (SCGExperiment.ipynb settings)
    ("c1_scg2", 200, "stress"),
    ("c2_scg50", 320, "init"),         # BASELINE config 2 at a chain count the oracle finishes in seconds
    ("c2_scg50", 320, "stress"),
    ("c3_mog2", 512, "stress"),        # BASELINE config 3 (2 modes, var 0.1), Lf=25
    ("c4_rw32", 320, "stress"),        # BASELINE config 4 target, easy and hard variants
    ("c4_rw32_hard", 320, "stress"),
    ("funnel3", 256, "stress"),
])
def test_propose_matches_oracle(name, n, regime):
    P = U.Problem(regime=regime, **U.CONFIGS[name])
    rep, _ = U.parity_report(P, n)
    _check(rep)


@pytest.mark.parametrize("n", [1, 7, 63, 64, 65, 129])
def test_ragged_chain_counts(n):
    """Tiles are 64 chains; every remainder must behave."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    rep, _ = U.parity_report(P, n, seed=n)
    _check(rep)


def test_empty_batch():
    P = U.Problem(**U.CONFIGS["c1_scg2"])
    dyn = P.product()
    x = torch.empty((0, 2), device="cuda")
    X, V, p = dyn.forward(x)
    assert X.shape == (0, 2) and p.shape == (0,)


@pytest.mark.parametrize("name", ["c1_scg2", "c2_scg50"])
def test_log_jac_mode(name):
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    rep, _ = U.parity_report(P, 192, log_jac=True)
    assert rep["Lx_kernel"] <= max(SAMPLE_TOL, 4 * rep["Lx_o32"]), rep
    assert rep["px_kernel"] <= max(2e-5 * 10, 4 * rep["px_o32"]), rep  # px is log|J| (O(1..10)) here


@pytest.mark.parametrize("kind,D", [("gaussian", 2), ("gaussian", 50), ("roughwell", 32), ("gmm", 2)])
def test_hmc_mode(kind, D):
    """hmc=True: zero nets, forward only, init_v forwarded (utils/sampler.py:29-31, utils/dynamics.py:73-76)."""
    P = U.Problem(kind=kind, D=D, T=10, eps=0.05, hmc=True)
    rep, _ = U.parity_report(P, 300)
    _check(rep)


@pytest.mark.parametrize("name", ["c1_scg2", "c2_scg50", "c4_rw32"])
def test_forward_backward_methods(name):
    """Dynamics.forward / backward with init_v against the oracle, and the exact-inverse property."""
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    dyn = P.product()
    d = P.draws(192)
    o64 = P.oracle(torch.float64)
    x, v = torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["v_f"]).cuda()
    X, V, lj = dyn.forward(x, init_v=v, log_jac=True)
    Xo, Vo, ljo = o64.forward(U.t64(d["x"]), U.t64(d["v_f"]), log_jac=True)
    assert U.max_rel(X.cpu().numpy(), Xo.numpy()) <= SAMPLE_TOL
    assert U.max_rel(V.cpu().numpy(), Vo.numpy()) <= SAMPLE_TOL
    assert U.max_rel(lj.cpu().numpy(), ljo.numpy()) <= SAMPLE_TOL
    Xb, Vb, ljb = dyn.backward(x, init_v=v, log_jac=True)
    Xbo, Vbo, ljbo = o64.backward(U.t64(d["x"]), U.t64(d["v_f"]), log_jac=True)
    assert U.max_rel(Xb.cpu().numpy(), Xbo.numpy()) <= SAMPLE_TOL
    assert U.max_rel(ljb.cpu().numpy(), ljbo.numpy()) <= SAMPLE_TOL
    # backward(forward(x, v)) == (x, v), log|J| cancels (utils/dynamics.py:159-201 inverts :115-157)
    x2, v2, lj2 = dyn.backward(X, init_v=V, log_jac=True)
    assert U.max_rel(x2.cpu().numpy(), d["x"]) <= 5e-5
    assert U.max_rel(v2.cpu().numpy(), d["v_f"]) <= 5e-5
    assert float((lj + lj2).abs().max()) <= 5e-5 * max(1.0, float(lj.abs().max()))
    # p_accept from forward() equals the component call on its own outputs
    _, _, p = dyn.forward(x, init_v=v)
    p2 = dyn.p_accept(x, v, X, V, lj)
    assert float((p - p2).abs().max()) <= P_TOL  # two fp32 orders of the same cancellation-prone Hamiltonian difference


def test_components_match_oracle():
    for name in ("c1_scg2", "c2_scg50", "c3_mog2", "c4_rw32_hard", "funnel3"):
        P = U.Problem(regime="stress", **U.CONFIGS[name])
        dyn = P.product()
        o = P.oracle(torch.float64)
        d = P.draws(256)
        x, v = torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["v_f"]).cuda()
        e = dyn.energy(x).cpu().numpy()
        assert U.max_rel(e, o.energy(U.t64(d["x"])).numpy()) <= 1e-5, name
        g = dyn.grad_energy(x).cpu().numpy()
        # sin(x / 0.01): one fp32 ulp of the argument (|arg| ~ 300) is already 3e-5 in the sine
        gtol = 5e-5 if name == "c4_rw32_hard" else 1e-5
        assert U.max_rel(g, o.grad_energy(U.t64(d["x"])).numpy()) <= gtol, name
        k = dyn.kinetic(v).cpu().numpy()
        assert U.max_rel(k, o.kinetic(U.t64(d["v_f"])).numpy()) <= 1e-6, name
        h = dyn.hamiltonian(x, v).cpu().numpy()
        assert U.max_rel(h, o.hamiltonian(U.t64(d["x"]), U.t64(d["v_f"])).numpy()) <= 1e-5, name
        # energy closure called directly, like the reference's fn(x)
        e2 = P.dist.get_energy_function()(x).cpu().numpy()
        assert np.array_equal(e, e2)
        for which, net in (("XNet", o.xnet), ("VNet", o.vnet)):
            S, T, Q = dyn.net_apply(which, x, v, 3.0)
            So, To, Qo = U.O.net_apply(net, U.t64(d["x"]), U.t64(d["v_f"]), o.format_time(3.0, 256))
            for a, b in ((S, So), (T, To), (Q, Qo)):
                assert U.max_rel(a.cpu().numpy(), b.numpy()) <= 1e-5, (name, which)


def test_temperature():
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    from l2hmc_b200 import Dynamics
    dyn = Dynamics(P.D, P.dist.get_energy_function(), T=P.T, eps=P.eps, net_factory=P.net_factory(), use_temperature=True)
    dyn.mask = P.mask
    dyn.temperature = 2.5
    d = P.draws(128)
    o = P.oracle(torch.float64, temperature=2.5)
    x, v = torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["v_f"]).cuda()
    X, V, p = dyn.forward(x, init_v=v)
    Xo, Vo, po = o.forward(U.t64(d["x"]), U.t64(d["v_f"]))
    assert U.max_rel(X.cpu().numpy(), Xo.numpy()) <= SAMPLE_TOL
    assert float(np.max(np.abs(p.cpu().numpy() - po.numpy()))) <= P_TOL


def test_mask_is_assignable():
    """eval_sampler.py:156 re-injects dynamics.mask after construction."""
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    dyn = P.product()
    d = P.draws(64)
    x, v = torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["v_f"]).cuda()
    X1, _, _ = dyn.forward(x, init_v=v)
    new_mask = 1.0 - P.mask
    dyn.mask = new_mask
    X2, _, _ = dyn.forward(x, init_v=v)
    assert not torch.equal(X1, X2)
    P.mask = new_mask
    Xo, _, _ = P.oracle(torch.float64).forward(U.t64(d["x"]), U.t64(d["v_f"]))
    assert U.max_rel(X2.cpu().numpy(), Xo.numpy()) <= SAMPLE_TOL


def test_tf_accept_and_nan_semantics():
    from l2hmc_b200 import tf_accept
    x = torch.zeros((4, 3), device="cuda")
    Lx = torch.ones((4, 3), device="cuda")
    px = torch.tensor([0.0, 0.5, 1.0, float("nan")], device="cuda")
    u = torch.tensor([0.0, 0.6, 0.999, 0.0], device="cuda")
    out = tf_accept(x, Lx, px, u=u).cpu().numpy()
    # px - u >= 0 -> take Lx (utils/sampler.py:53-55); NaN compares false
    assert out[:, 0].tolist() == [1.0, 0.0, 1.0, 0.0]
    # p_accept maps non-finite to 0 (utils/dynamics.py:309)
    P = U.Problem(**U.CONFIGS["c1_scg2"])
    dyn = P.product()
    z = torch.zeros((3, 2), device="cuda")
    bad = torch.tensor([[float("inf"), 0.0], [float("nan"), 0.0], [0.0, 0.0]], device="cuda")
    p = dyn.p_accept(z, z, bad, z, torch.zeros(3, device="cuda")).cpu().numpy()
    assert p[0] == 0.0 and p[1] == 0.0 and p[2] == 1.0


def test_philox_device_matches_host_twin_and_reinjection():
    """In-kernel Philox == l2hmc_b200.philox (bits/uniforms exact, normals to libm ulps), and a
    transition with in-kernel randomness is bit-identical to one with those arrays injected."""
    import ctypes as C
    from l2hmc_b200 import philox, propose
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.product(seed=1234)
    n, D = 300, P.D
    v = torch.empty((n, D), device="cuda")
    dirb = torch.empty((n,), dtype=torch.uint8, device="cuda")
    u = torch.empty((n,), device="cuda")
    dyn._chk(dyn._lib.l2hmc_philox_fill(dyn._ctx, n, 0, 1234, 7, v.data_ptr(), dirb.data_ptr(), u.data_ptr(), None))
    torch.cuda.synchronize()
    hd, hu = philox.direction_and_uniform(1234, 7, n)
    hv = philox.normals(1234, 7, n, D)
    assert np.array_equal(dirb.cpu().numpy(), hd)
    assert np.array_equal(u.cpu().numpy(), hu)
    assert np.max(np.abs(v.cpu().numpy() - hv)) <= 2e-6
    x = torch.as_tensor(P.draws(n)["x"]).cuda()
    a = dyn._transition(x, dir_mode=3, do_mh=True, counter=7)
    b = dyn._transition(x, v=v, direction=dirb, u=u, do_mh=True, counter=99)
    for k in ("Lx", "Lv", "px", "x_next", "accepted"):
        assert torch.equal(a[k], b[k]), k
    # sharding invariance: the second half of the chains with chain_offset gives the same numbers
    h = n // 2
    c = dyn._transition(x[h:].contiguous(), dir_mode=3, do_mh=True, counter=7, chain_offset=h)
    assert torch.equal(c["Lx"], a["Lx"][h:]) and torch.equal(c["px"], a["px"][h:])


def test_multi_transition_equals_host_loop():
    """n_transitions=K in one launch == K single launches fed back by the host
    (the reference's loop of sess.run, SCGExperiment.ipynb:291-298)."""
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    dyn = P.product(seed=5)
    x0 = torch.as_tensor(P.draws(200)["x"]).cuda()
    K = 6
    a = dyn._transition(x0, dir_mode=3, do_mh=True, n_transitions=K, counter=100)
    x = x0
    for t in range(K):
        b = dyn._transition(x, dir_mode=3, do_mh=True, counter=100 + t)
        x = b["x_next"]
    assert torch.equal(a["x_next"], b["x_next"])
    assert torch.equal(a["px"], b["px"])


def test_transition_host_equals_device_path():
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.product(seed=9)
    d = P.draws(200)
    dev = dyn._transition(torch.as_tensor(d["x"]).cuda(), dir_mode=3, do_mh=True, counter=3)
    host = dyn.transition_host(d["x"], counter=3)
    torch.cuda.synchronize()
    assert np.array_equal(host["Lx"], dev["Lx"].cpu().numpy())
    assert np.array_equal(host["px"], dev["px"].cpu().numpy())
    assert np.array_equal(host["x_next"], dev["x_next"].cpu().numpy())
    assert np.array_equal(host["accepted"], dev["accepted"].cpu().numpy())


@pytest.mark.parametrize("name,n", [("c2_scg50", 40000), ("c1_scg2", 33001)])
def test_transition_host_chunk_pipeline_equals_device_path(name, n):
    """From 32768 chains on, l2hmc_transition_host cuts the batch into chunks pipelined over streams (H2D / kernel / D2H
    of different chunks overlap); chains are independent and Philox is keyed by the global chain id, so the result
    must be bit-identical to one device-resident launch -- with in-kernel and with injected randomness."""
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    dyn = P.product(seed=13)
    d = P.draws(n, seed=4)
    x = torch.as_tensor(d["x"]).cuda()
    dev = dyn._transition(x, dir_mode=3, do_mh=True, counter=7, chain_offset=1000)
    host = dyn.transition_host(d["x"], counter=7, chain_offset=1000)
    for k in ("Lx", "Lv", "px", "x_next", "accepted"):
        assert np.array_equal(host[k], dev[k].cpu().numpy()), k
    v_sel = np.where(d["dir"][:, None] != 0, d["v_f"], d["v_b"]).astype(np.float32)
    g = lambda a: torch.as_tensor(a).cuda()  # noqa: E731
    dev = dyn._transition(x, v=g(v_sel), direction=g(d["dir"]), u=g(d["u"]), do_mh=True, counter=8)
    host = dyn.transition_host(d["x"], v=v_sel, direction=d["dir"], u=d["u"], counter=8)
    for k in ("Lx", "Lv", "px", "x_next", "accepted"):
        assert np.array_equal(host[k], dev[k].cpu().numpy()), k


def test_chain_operator_matches_oracle():
    from l2hmc_b200 import chain_operator
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    dyn = P.product()
    n, steps = 128, 3
    rngs, dirs, vfs, vbs = [], [], [], []
    g = lambda a: torch.as_tensor(a).cuda()
    base = P.draws(n, seed=3)
    for s in range(steps):
        d = P.draws(n, seed=10 + s)
        sel = np.where(d["dir"][:, None] != 0, d["v_f"], d["v_b"]).astype(np.float32)
        rngs.append({"direction": g(d["dir"]), "v": g(sel)})
        dirs.append(torch.as_tensor(d["dir"].astype(np.float32)))
        vfs.append(torch.as_tensor(d["v_f"]))
        vbs.append(torch.as_tensor(d["v_b"]))
    rngs.append({"u": g(base["u"])})
    fx, fv, p, outs = chain_operator(g(base["x"]), dyn, steps, init_v=g(base["v_f"]), do_mh_step=True, rng=rngs)
    ox, ov, op, oo = U.O.chain_operator(U.t64(base["x"]), P.oracle(torch.float64), steps, init_v=U.t64(base["v_f"]),
                                         directions=dirs, v_fs=vfs, v_bs=vbs, u=U.t64(base["u"]), do_mh_step=True)
    assert U.max_rel(fx.cpu().numpy(), ox.numpy()) <= 5e-5
    assert U.max_rel(fv.cpu().numpy(), ov.numpy()) <= 5e-5
    assert float(np.max(np.abs(p.cpu().numpy() - op.numpy()))) <= 2e-4


def test_golden_fixtures():
    """Committed fp64-oracle vectors (tests/golden/make_golden.py) reproduced by the CUDA path."""
    import glob
    import os
    files = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                   if not os.path.basename(f).startswith("ref_"))  # ref_*: test_reference_pin.py / test_gpu_reference_pin.py
    assert files, "no golden fixtures committed"
    import golden_io
    for f in files:
        P, d, ref = golden_io.load(f)
        rk = U.run_kernel_propose(P, d)
        assert U.max_rel(rk["Lx"], ref["Lx"]) <= SAMPLE_TOL, f
        assert U.max_rel(rk["Lv"], ref["Lv"]) <= SAMPLE_TOL, f
        assert float(np.max(np.abs(rk["px"] - ref["px"]))) <= 2e-4, f
        assert abs(float(rk["px"].mean()) - float(ref["px"].mean())) <= 2e-5, f


def test_full_size_properties_config2():
    """BASELINE config 2 at full size (2^18 chains, 50-d, width 100, Lf=10): size-independent
    properties instead of an oracle run -- exact inverse, log|J| cancellation, p in [0,1], and the
    mean accept probability of a random subsample against the oracle."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.product(seed=3)
    n = 1 << 18
    rng = np.random.default_rng(0)
    x = torch.as_tensor(P.x0(n, rng)).cuda()
    v = torch.randn((n, P.D), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    X, V, lj = dyn.forward(x, init_v=v, log_jac=True)
    x2, v2, lj2 = dyn.backward(X, init_v=V, log_jac=True)
    scale = float(x.abs().max())
    assert float((x2 - x).abs().max()) / scale <= 1e-4
    assert float((v2 - v).abs().max()) / max(1.0, float(v.abs().max())) <= 1e-4
    assert float((lj + lj2).abs().max()) <= 1e-4 * max(1.0, float(lj.abs().max()))
    _, _, p = dyn.forward(x, init_v=v)
    assert float(p.min()) >= 0.0 and float(p.max()) <= 1.0 and bool(torch.isfinite(p).all())
    idx = rng.choice(n, 256, replace=False)
    o = P.oracle(torch.float64)
    _, _, po = o.forward(U.t64(x[idx].cpu().numpy()), U.t64(v[idx].cpu().numpy()))
    assert float(np.max(np.abs(p[idx].cpu().numpy() - po.numpy()))) <= 1e-4


# ---- tensor-core kernel (tcgen05, 3xTF32) ---------------------------------------------------------------
@pytest.mark.parametrize("name,n,regime", [
    ("c2_scg50", 320, "init"),
    ("c2_scg50", 320, "stress"),
    ("c2_scg50", 129, "stress"),   # ragged: one full 128-chain tile + 1
    ("c4_rw32", 320, "stress"),
    ("c4_rw32_hard", 200, "stress"),
])
def test_tc_kernel_matches_oracle(name, n, regime):
    P = U.Problem(regime=regime, **U.CONFIGS[name])
    dyn = P.product(kernel="tc")
    rep, _ = U.parity_report(P, n, dyn=dyn)
    assert dyn.kernel_name in ("tc_3xf16", "tc_3xtf32")
    _check(rep)
    assert not dyn.fp16_range_exceeded()


def test_tc_kernel_multi_transition_and_philox():
    """Fused transitions equal the host loop, and in-kernel Philox equals injected randomness (TC kernel)."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.product(seed=5, kernel="tc")
    x0 = torch.as_tensor(P.draws(300)["x"]).cuda()
    K = 3
    a = dyn._transition(x0, dir_mode=3, do_mh=True, n_transitions=K, counter=100)
    x = x0
    for t in range(K):
        b = dyn._transition(x, dir_mode=3, do_mh=True, counter=100 + t)
        x = b["x_next"]
    assert torch.equal(a["x_next"], b["x_next"]) and torch.equal(a["px"], b["px"])
    # same numbers as the tile kernel's generator -> same decisions up to rounding
    tile = P.product(seed=5, kernel="tile")
    c = tile._transition(x0, dir_mode=3, do_mh=True, counter=100)
    d = dyn._transition(x0, dir_mode=3, do_mh=True, counter=100)
    assert U.max_rel(d["Lx"].cpu().numpy(), c["Lx"].cpu().numpy()) <= 2e-5
    assert float((d["px"] - c["px"]).abs().max()) <= 2e-4


def _fixed_inputs(P, n):
    d = P.draws(n)
    return (torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["v_f"]).cuda(), torch.as_tensor(d["dir"]).cuda(),
            torch.as_tensor(d["u"]).cuda())


@pytest.mark.parametrize("name,n", [("c2_scg50", 700), ("c4_rw32", 300)])
def test_tc_specialised_and_generic_kernels_agree(name, n, monkeypatch):
    """The shape-specialised compute path (kernel_tc_s.cuh: split heads GEMM, overlapped epilogues, interleaved net
    input) and the generic tensor-core kernel run the same transition: same injected randomness -> same samples and
    accept probabilities up to fp32 reordering, same Metropolis decisions outside that noise."""
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    dyn = P.product(kernel="tc")
    x, v, dr, u = _fixed_inputs(P, n)
    kw = dict(v=v, direction=dr, u=u, do_mh=True)
    monkeypatch.delenv("L2HMC_TC_GENERIC", raising=False)
    a = dyn._transition(x, **kw)
    monkeypatch.setenv("L2HMC_TC_GENERIC", "1")
    b = dyn._transition(x, **kw)
    monkeypatch.delenv("L2HMC_TC_GENERIC", raising=False)
    assert not torch.equal(a["Lx"], b["Lx"]) or name != "c2_scg50"  # two different code paths really ran
    assert U.max_rel(a["Lx"].cpu().numpy(), b["Lx"].cpu().numpy()) <= SAMPLE_TOL
    assert U.max_rel(a["Lv"].cpu().numpy(), b["Lv"].cpu().numpy()) <= SAMPLE_TOL
    dp = (a["px"] - b["px"]).abs().cpu().numpy()
    assert float(dp.max()) <= 2 * P_TOL
    flips = (a["accepted"] != b["accepted"]).cpu().numpy()
    margin = np.abs(a["px"].cpu().numpy() - u.cpu().numpy())
    assert not np.any(flips & (margin > 1e-4))


def test_tc_specialised_kernel_follows_eps_and_temperature():
    """The specialised kernel's pre-multiplied head constants depend on eps: l2hmc_set_eps must repack them; the
    temperature enters through grad U and the Hamiltonian (utils/dynamics.py:203-212)."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    from l2hmc_b200 import Dynamics
    dyn = Dynamics(P.D, P.dist.get_energy_function(), T=P.T, eps=0.3, net_factory=P.net_factory(), use_temperature=True,
                   kernel="tc")
    dyn.mask = P.mask
    dyn.eps = P.eps          # changed after the nets were packed
    dyn.temperature = 1.7
    d = P.draws(256)
    o = P.oracle(torch.float64, temperature=1.7)
    x, v = torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["v_f"]).cuda()
    X, V, p = dyn.forward(x, init_v=v)
    assert dyn.kernel_name in ("tc_3xf16", "tc_3xtf32")
    Xo, Vo, po = o.forward(U.t64(d["x"]), U.t64(d["v_f"]))
    assert U.max_rel(X.cpu().numpy(), Xo.numpy()) <= SAMPLE_TOL
    assert U.max_rel(V.cpu().numpy(), Vo.numpy()) <= SAMPLE_TOL
    assert float(np.max(np.abs(p.cpu().numpy() - po.numpy()))) <= P_TOL
    Xb, Vb, pb = dyn.backward(x, init_v=v)
    Xob, Vob, pob = o.backward(U.t64(d["x"]), U.t64(d["v_f"]))
    assert U.max_rel(Xb.cpu().numpy(), Xob.numpy()) <= SAMPLE_TOL
    assert float(np.max(np.abs(pb.cpu().numpy() - pob.numpy()))) <= P_TOL


@pytest.mark.parametrize("kind,D,H,T,mu_shift", [
    ("gaussian", 52, 104, 6, 0.7),   # no pad dimension / hidden unit: explicit biases in the specialised kernel
    ("gaussian", 49, 97, 7, -0.4),   # three pad dimensions, seven pad hidden units, odd Lf, non-zero mean
    ("gaussian", 51, 103, 5, 0.0),   # one pad dimension only: no room for the direction one-hot -> explicit biases
    ("roughwell", 30, 100, 4, 0.0),  # the 8 x 13 instantiation with pad dimensions
    ("gaussian", 29, 98, 5, 0.3),    # Gaussian grad GEMM on the 8 x 13 instantiation
])
def test_tc_specialised_kernel_edge_shapes(kind, D, H, T, mu_shift):
    """Every (x_dim, width) that maps to a specialised instantiation (x_dim padded to 52 or 32, width padded to 104), with
    and without room for the biases in the GEMMs, against the oracle."""
    kw = dict(mu=np.full(D, mu_shift)) if kind == "gaussian" else dict(easy=True)
    P = U.Problem(kind=kind, D=D, H=H, T=T, eps=0.1, regime="stress", **kw)
    dyn = P.product(kernel="tc")
    rep, _ = U.parity_report(P, 200, dyn=dyn)
    assert dyn.kernel_name in ("tc_3xf16", "tc_3xtf32")
    _check(rep)
    assert not dyn.fp16_range_exceeded()


def test_tc_fp16_and_tf32_splits_agree_and_range_flag(monkeypatch):
    """The specialised kernel's fp16 operand split (three kind::f16 MMAs per product) against its tf32 split on the same
    inputs; and the sticky range flag when an operand leaves the fp16 range."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    x, v, dr, u = _fixed_inputs(P, 600)
    kw = dict(v=v, direction=dr, u=u, do_mh=True)
    monkeypatch.delenv("L2HMC_TC_F16", raising=False)
    a = P.product(kernel="tc")._transition(x, **kw)
    monkeypatch.setenv("L2HMC_TC_F16", "0")       # read when the context is created
    b = P.product(kernel="tc")._transition(x, **kw)
    monkeypatch.delenv("L2HMC_TC_F16", raising=False)
    assert not torch.equal(a["Lx"], b["Lx"])
    assert U.max_rel(a["Lx"].cpu().numpy(), b["Lx"].cpu().numpy()) <= SAMPLE_TOL
    assert U.max_rel(a["Lv"].cpu().numpy(), b["Lv"].cpu().numpy()) <= SAMPLE_TOL
    assert float((a["px"] - b["px"]).abs().max()) <= 2 * P_TOL
    # the synchronous host-buffer entry point notices an out-of-range activation, repeats the call with the tf32 split
    # and stays on it
    big = x.clone()
    big[0, 0] = 1.0e5                              # outside the fp16 range
    ref = P.product(kernel="tc")
    monkeypatch.setenv("L2HMC_TC_F16", "0")
    ref32 = P.product(kernel="tc")
    r32 = ref32._transition(big, **kw)
    monkeypatch.delenv("L2HMC_TC_F16", raising=False)
    st = np.zeros(2, np.float64)
    host = ref.transition_host(big.cpu().numpy(), v=v.cpu().numpy(), direction=dr.cpu().numpy(), u=u.cpu().numpy(), do_mh=True,
                               stats=st)
    assert ref.kernel_name == "tc_3xtf32" and ref.fp16_range_exceeded()
    assert np.array_equal(host["x_next"], r32["x_next"].cpu().numpy())
    assert st[1] == float(host["accepted"].sum())   # the repeated call is not counted twice

def test_fp16_range_status_is_per_context_and_later_launches_fall_back_to_tf32():
    """The range flag lives in the CONTEXT (pinned status word), not in a process-wide device symbol: a context that met
    an out-of-range activation reports it and runs the tf32 split from its next launch on, without any caller action;
    other contexts are untouched.  The chains of the tripping launch that were affected are rejected (p = 0), never
    accepted with garbage."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    x, v, dr, u = _fixed_inputs(P, 600)
    kw = dict(v=v, direction=dr, u=u, do_mh=True)
    a, b = P.product(kernel="tc"), P.product(kernel="tc")
    big = x.clone()
    big[5, 3] = 2.0e5
    o = a._transition(big, **kw)
    assert a.fp16_range_exceeded() and not b.fp16_range_exceeded()
    assert a.status_flags() & 1 and not (b.status_flags() & 1)
    bad = ~torch.isfinite(o["Lx"]).all(dim=1)
    assert bool(bad[5]) and bool((o["px"][bad] == 0).all()) and bool((o["accepted"][bad] == 0).all())
    assert torch.equal(o["x_next"][bad], big[bad])
    # next launch of the same context: tf32 split, finite results for the same out-of-range input
    o2 = a._transition(big, **kw)
    assert a.kernel_name == "tc_3xtf32" and bool(torch.isfinite(o2["Lx"]).all())
    b._transition(x, **kw)
    assert b.kernel_name == "tc_3xf16"
    # clearing the word re-arms the fp16 split
    assert a.status_flags(clear=True) & 1 and a.status_flags() == 0
    a._transition(x, **kw)
    assert a.kernel_name == "tc_3xf16" and not a.fp16_range_exceeded()


@pytest.mark.parametrize("wscale,xscale", [(1e-4, 1.0), (1.0, 1e-5), (1e-3, 1e-3)])
def test_tc_fp16_split_with_small_operands(wscale, xscale, monkeypatch):
    """Underflow side of the fp16 operand split.  hi / lo fp16 parts bottom out at the fp16 subnormal spacing, so every
    operand element carries an ABSOLUTE error of up to 2^-25 = 3e-8 (fp32-accurate relative to O(1) values -- the metric
    of every parity test: error / max(1, |ref|)).  Measured here with head weights of 1e-4 x their usual size (hi / lo
    in the subnormal range) and with states / momenta of 1e-5 and 1e-3 x their usual size:
      * in the parity metric the kernel stays at the fp32 oracle's own level in all three cases;
      * relative to the results' OWN magnitude the floor shows once states are ~1e-3 (6e-5 measured): bounded here by
        1e-4, and the tf32 split (L2HMC_TC_F16=0: fp32 exponent range) is at the oracle's level there too."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    for net in (P.xnet, P.vnet):
        for k in ("Ws", "Wt", "Wq", "bs", "bt", "bq"):
            net[k] = (net[k] * wscale).astype(np.float32)
    d = P.draws(256, 5)
    for k in ("x", "v_f", "v_b"):
        d[k] = (d[k] * xscale).astype(np.float32)
    r64 = U.run_oracle_propose(P, d, torch.float64)
    r32 = U.run_oracle_propose(P, d, torch.float32)
    monkeypatch.delenv("L2HMC_TC_F16", raising=False)
    dyn = P.product(kernel="tc")
    rk = U.run_kernel_propose(P, d, dyn=dyn)
    assert dyn.kernel_name == "tc_3xf16" and not dyn.fp16_range_exceeded()
    monkeypatch.setenv("L2HMC_TC_F16", "0")
    dyn32 = P.product(kernel="tc")
    rt = U.run_kernel_propose(P, d, dyn=dyn32)
    monkeypatch.delenv("L2HMC_TC_F16", raising=False)
    assert dyn32.kernel_name == "tc_3xtf32"
    for k in ("Lx", "Lv"):
        assert U.max_rel(rk[k], r64[k]) <= U.max_rel(r32[k], r64[k]) + NORTH_STAR_TOL, k
        own = max(float(np.abs(r64[k]).max()), 1e-30)
        err_k, err_t, err_o = (float(np.abs(r[k] - r64[k]).max()) / own for r in (rk, rt, r32))
        assert err_k <= 1e-4, (k, err_k, err_o)
        assert err_t <= err_o + NORTH_STAR_TOL, (k, err_t, err_o)
    o32 = float(np.abs(r32["px"] - r64["px"]).max())
    assert float(np.abs(rk["px"] - r64["px"]).max()) <= 4 * o32 + NORTH_STAR_TOL
    assert float(np.abs(rt["px"] - r64["px"]).max()) <= 4 * o32 + NORTH_STAR_TOL


@pytest.mark.parametrize("name,n,kernel", [("c1_scg2", 333, "small"), ("c3_mog2", 200, "tile"), ("c2_scg50", 300, "tc"),
                                           ("c2_scg50", 200, "tile"), ("c4_rw32", 200, "layered")])
def test_accept_statistics_and_trace_come_from_the_kernel(name, n, kernel):
    """l2hmc_transition_args.stats / .trace: (sum of px, number accepted) reduced in the kernel over every chain and every
    fused transition, and the Metropolis output of every fused transition -- against separate single-transition calls."""
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    dyn = P.product(kernel=kernel, seed=9)
    x = torch.as_tensor(P.x0(n, np.random.default_rng(2))).cuda()
    K = 3
    c0 = 40
    stats = torch.zeros(2, dtype=torch.float64, device="cuda")
    trace = torch.empty((K, n, P.D), dtype=torch.float32, device="cuda")
    l0 = dyn.launch_count
    o = dyn._transition(x, dir_mode=3, do_mh=True, n_transitions=K, counter=c0, stats=stats, trace=trace)
    fused_launches = dyn.launch_count - l0
    cur, sum_p, n_acc = x, 0.0, 0
    for t in range(K):
        s = dyn._transition(cur, dir_mode=3, do_mh=True, counter=c0 + t)
        assert torch.equal(trace[t], s["x_next"]), t
        sum_p += float(s["px"].double().sum())
        n_acc += int(s["accepted"].sum())
        cur = s["x_next"]
    assert torch.equal(o["x_next"], cur)
    assert float(stats[1]) == n_acc
    assert abs(float(stats[0]) - sum_p) <= 1e-4 * max(1.0, sum_p)
    if kernel != "layered":
        assert fused_launches == 1
    # the accumulators add over calls
    dyn._transition(x, dir_mode=3, do_mh=True, counter=c0, stats=stats)
    assert float(stats[1]) >= n_acc


def test_host_entry_point_returns_accept_statistics():
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.product(seed=4)
    for n in (5000, 40000):   # single launch / chunk pipeline over three streams
        x = P.x0(n, np.random.default_rng(3))
        st = np.zeros(2, np.float64)
        out = dyn.transition_host(x, counter=7, stats=st)
        assert st[1] == float(out["accepted"].sum())
        assert abs(st[0] - float(out["px"].astype(np.float64).sum())) <= 1e-4 * max(1.0, st[0])


@pytest.mark.parametrize("kind,D,H,T", [("gaussian", 20, 64, 6), ("gaussian", 40, 100, 5), ("gaussian", 16, 32, 8), ("roughwell", 50, 64, 4),
                                        ("gaussian", 8, 100, 7), ("roughwell", 24, 48, 5), ("gaussian", 33, 72, 3), ("gaussian", 50, 88, 4)])
def test_tc_run_time_shape_kernel(kind, D, H, T, monkeypatch):
    """Every (x_dim <= 52, width <= 104) runs the overlapped tensor-core pipeline of kernel_tc_s.cuh: the chunk counts of
    the two benchmark shapes are compile-time constants, all other shapes use the instantiation that reads them from the
    launch arguments (instead of the 2.2x slower generic kernel).  Parity against the oracle, and a different result in the
    last bits than the generic kernel's (L2HMC_TC_GENERIC=1) proves which one ran."""
    kw = dict(mu=np.full(D, 0.3)) if kind == "gaussian" else dict(easy=True)
    P = U.Problem(kind=kind, D=D, H=H, T=T, eps=0.1, regime="stress", **kw)
    monkeypatch.delenv("L2HMC_TC_GENERIC", raising=False)
    dyn = P.product(kernel="tc")
    rep, (d, r64, r32, rk) = U.parity_report(P, 300, dyn=dyn)
    assert dyn.kernel_name in ("tc_3xf16", "tc_3xtf32")
    _check(rep)
    assert not dyn.fp16_range_exceeded()
    monkeypatch.setenv("L2HMC_TC_GENERIC", "1")
    gen = U.run_kernel_propose(P, d, dyn=P.product(kernel="tc"))
    monkeypatch.delenv("L2HMC_TC_GENERIC", raising=False)
    assert U.max_rel(gen["Lx"], rk["Lx"]) <= SAMPLE_TOL and not np.array_equal(gen["Lx"], rk["Lx"])
    # fused multi-transition launch and chain mode on a run-time shape
    x = torch.as_tensor(d["x"]).cuda()
    o3 = dyn._transition(x, dir_mode=3, do_mh=True, n_transitions=2, counter=5)
    s1 = dyn._transition(x, dir_mode=3, do_mh=True, counter=5)
    s2 = dyn._transition(s1["x_next"], dir_mode=3, do_mh=True, counter=6)
    assert torch.equal(o3["x_next"], s2["x_next"])


@pytest.mark.parametrize("name,n", [("c1_scg2", 400), ("c3_mog2", 512)])
def test_small_kernel_fast_and_libm_math_agree(name, n, monkeypatch):
    """The one-chain-per-thread kernel evaluates exp / tanh of the updates with ex2.approx / rcp.approx (as the tensor-core
    epilogues do); L2HMC_SMALL_FAST_MATH=0 selects expf / tanhf.  Both meet the parity bar, and they differ in the last
    bits (which proves that the switch selects another instantiation)."""
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    monkeypatch.delenv("L2HMC_SMALL_FAST_MATH", raising=False)
    rep, (d, r64, r32, rk) = U.parity_report(P, n, dyn=P.product(kernel="small"))
    _check(rep)
    monkeypatch.setenv("L2HMC_SMALL_FAST_MATH", "0")
    rep0, (_, _, _, rk0) = U.parity_report(P, n, dyn=P.product(kernel="small"))
    _check(rep0)
    assert U.max_rel(rk0["Lx"], rk["Lx"]) <= SAMPLE_TOL and not np.array_equal(rk0["Lx"], rk["Lx"])


@pytest.mark.parametrize("f16", ["1", "0"])
@pytest.mark.parametrize("kind,D,H,T", [("roughwell", 32, 100, 10), ("gaussian", 40, 100, 5), ("gaussian", 8, 100, 4), ("gaussian", 48, 50, 3),
                                        ("gaussian", 30, 100, 4), ("roughwell", 18, 60, 5)])
def test_tc_biases_in_the_gemms_on_every_shape_with_a_pad_hidden_unit(kind, D, H, T, f16, monkeypatch):
    """Biases ride in the GEMMs (kernel_tc_s.cuh, BIASG) wherever the width leaves a pad hidden unit: the direction one-hot
    that selects the time-embedding bias row sits in the pad dimensions of the last 4-dim chunk when x_dim leaves two
    (config 2; here 30-d, 18-d), else in a K step of its own behind the net input (config 4's 32-d shape, 40-d, 8-d, 48-d).
    Parity against the oracle on both operand splits, and agreement with the explicit-bias kernel (L2HMC_TC_BIASG=0) up
    to the summation order -- the last bits differ, which proves that the other code path ran."""
    kw = dict(mu=np.full(D, 0.3)) if kind == "gaussian" else dict(easy=True)
    P = U.Problem(kind=kind, D=D, H=H, T=T, eps=0.1, regime="stress", **kw)
    monkeypatch.setenv("L2HMC_TC_F16", f16)
    monkeypatch.delenv("L2HMC_TC_BIASG", raising=False)
    dyn = P.product(kernel="tc")
    rep, (d, r64, r32, rk) = U.parity_report(P, 300, dyn=dyn)
    assert dyn.kernel_name == ("tc_3xf16" if f16 == "1" else "tc_3xtf32")
    _check(rep)
    assert not dyn.fp16_range_exceeded()
    monkeypatch.setenv("L2HMC_TC_BIASG", "0")
    exp = U.run_kernel_propose(P, d, dyn=P.product(kernel="tc"))
    monkeypatch.delenv("L2HMC_TC_BIASG", raising=False)
    assert U.max_rel(exp["Lx"], rk["Lx"]) <= SAMPLE_TOL and U.max_rel(exp["Lv"], rk["Lv"]) <= SAMPLE_TOL
    assert not np.array_equal(exp["Lx"], rk["Lx"])
    # both directions in one batch were exercised (forward chains read tb[t], backward chains tb[T-1-t])
    assert 0 < int(np.asarray(d["dir"]).sum()) < len(d["dir"])
