"""GPU parity of the layered engine (csrc/layered.cuh) against the CPU oracle, through the Python mirror -> C ABI.

Covers BASELINE config 5 (the MNIST-VAE posterior target: decoder energy + aux-conditioned width-200 nets,
mnist_vae.py:104-178) at the reference's layer sizes and in miniature / ragged shapes, and the same engine on the
closed-form targets (where it must agree with the fused kernels' oracle too).

Tolerances: as tests/test_gpu_parity.py -- samples within O32 + 1e-5, per-chain accept probability within 4 * O32 + 1e-5 (O32: the fp32 oracle twin's own error on the
same inputs; at the full layer sizes the energy is a sum over 784 pixels of O(500), so fp32 Hamiltonian differences
carry ~1e-4 of noise in either implementation), the mean accept probability within 1e-5 (or the fp32 twin's own).
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


@pytest.mark.parametrize("name,n", [
    ("c5_vae_mini", 300),
    ("c5_vae_ragged", 131),     # no dimension a multiple of 8, chains not a multiple of the 128-row tile
    ("c5_vae_noenc", 200),      # decoder energy with the notebook's zero aux branch in the nets
    ("c5_vae_full", 96),        # mnist_vae.py layer sizes: 50 / 1024-1024-784 / 512-512-200 / width 200 / Lf=15
])
def test_vae_target_matches_oracle(name, n):
    P = U.VaeProblem(**U.VAE_CONFIGS[name])
    dyn = P.product()
    assert dyn.kernel_name.startswith("layered")
    rep, _ = U.parity_report(P, n, dyn=dyn)
    _check(rep)


@pytest.mark.parametrize("name,n", [("c5_vae_mini", 200), ("c5_vae_ragged", 131), ("c5_vae_full", 64)])
def test_vae_target_on_fma_gemms(name, n):
    """kernel='layered_fma': the same engine with the fp32-FMA GEMMs instead of the tcgen05 3xTF32 ones."""
    P = U.VaeProblem(**U.VAE_CONFIGS[name])
    dyn = P.product(kernel="layered_fma")
    assert dyn.kernel_name == "layered_fma"
    rep, _ = U.parity_report(P, n, dyn=dyn)
    _check(rep)


def test_default_layered_gemms_are_tensor_core_fp16_split_with_a_range_guard(monkeypatch):
    """The layered engine's GEMMs default to the fp16 x3 operand split (half the tensor time and half the shared-memory
    operand bytes of the tf32 x3 split).  Its A path tracks |activation|: an out-of-range value raises the context's
    sticky status bit, and from the next call on the context uses the tf32 images kept beside the fp16 ones."""
    monkeypatch.delenv("L2HMC_LAYERED_GEMM", raising=False)
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])
    dyn = P.product()
    assert dyn.kernel_name == "layered_tc3xf16"
    d = P.draws(160)
    ok = U.run_kernel_propose(P, d, dyn=dyn)
    assert np.isfinite(ok["Lx"]).all() and not dyn.fp16_range_exceeded()
    monkeypatch.setenv("L2HMC_LAYERED_GEMM", "tf32")
    ref32 = P.product()
    assert ref32.kernel_name == "layered_tc3xtf32"
    monkeypatch.delenv("L2HMC_LAYERED_GEMM", raising=False)
    r32 = U.run_kernel_propose(P, d, dyn=ref32)
    assert U.max_rel(ok["Lx"], r32["Lx"]) <= SAMPLE_TOL and float(np.abs(ok["px"] - r32["px"]).max()) <= 2 * P_TOL
    big = dict(d)
    big["x"] = d["x"].copy()
    big["x"][7, 1] = 3.0e5                       # outside the fp16 range
    U.run_kernel_propose(P, big, dyn=dyn)
    assert dyn.fp16_range_exceeded() and dyn.kernel_name == "layered_tc3xtf32"
    again = U.run_kernel_propose(P, big, dyn=dyn)  # now on the tf32 images: equal to a tf32-only context
    want = U.run_kernel_propose(P, big, dyn=ref32)
    assert np.array_equal(again["x_next"], want["x_next"]) and np.array_equal(again["px"], want["px"])
    assert not ref32.fp16_range_exceeded()


@pytest.mark.parametrize("name,n", [("c5_vae_mini", 300), ("c5_vae_ragged", 131), ("c5_vae_full", 520)])
def test_presplit_operand_images_change_nothing(monkeypatch, name, n):
    """The decoder's activations / gradients travel between its GEMMs as pre-split fp16 hi / lo operand images written by
    the producing kernel and fetched by TMA (layered::SplitImage); L2HMC_LAYERED_PRESPLIT=0 makes every GEMM convert its
    own A operand instead.  Same split, same MMA order: bit-identical results (ragged widths, chains not a multiple of the
    256-row tile, more than one tile per SM column).  The energy VALUE sums its 784 terms in another order in the kernel
    that also writes the image, so the accept probability agrees to fp32 rounding only."""
    P = U.VaeProblem(**U.VAE_CONFIGS[name])
    d = P.draws(n)
    monkeypatch.delenv("L2HMC_LAYERED_PRESPLIT", raising=False)
    a = U.run_kernel_propose(P, d, dyn=P.product())
    monkeypatch.setenv("L2HMC_LAYERED_PRESPLIT", "0")
    b = U.run_kernel_propose(P, d, dyn=P.product())
    for k in ("Lx", "Lv"):
        assert np.array_equal(a[k], b[k]), k
    # Hamiltonians of O(500) (784 pixels) subtracted in fp32: per-chain maxima of ~1e-4 between two summation orders
    assert float(np.abs(a["px"] - b["px"]).max()) <= 10 * P_TOL and float(np.abs(a["px"] - b["px"]).mean()) <= 1e-5
    assert np.isfinite(a["Lx"]).all()


def test_presplit_kernel_variants_agree(monkeypatch):
    """L2HMC_LAYERED_PRESPLIT=1 keeps the 256-row GEMM kernel with a TMA-fed A; the default (2) is tc_gemm_pre_kernel (128-row
    tiles, two accumulators, epilogue overlapped with the next tile's MMAs).  Same k order, same epilogue: identical."""
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_full"])
    d = P.draws(300)
    monkeypatch.setenv("L2HMC_LAYERED_PRESPLIT", "1")
    a = U.run_kernel_propose(P, d, dyn=P.product())
    monkeypatch.setenv("L2HMC_LAYERED_PRESPLIT", "2")
    b = U.run_kernel_propose(P, d, dyn=P.product())
    for k in ("Lx", "Lv", "px", "x_next"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name,n", [("c5_vae_mini", 300), ("c5_vae_ragged", 131), ("c5_vae_noenc", 200), ("c5_vae_full", 700)])
def test_fused_net_kernel_equals_the_three_gemms(monkeypatch, name, n):
    """tc_net_kernel (one launch per S/T/Q net call, hidden activations as operand images in shared memory) against the
    three tc_gemm_pre_kernel launches it replaces (L2HMC_LAYERED_FUSED_NET=0): same split, k order and epilogue code, so
    the same bits; and fewer launches."""
    P = U.VaeProblem(**U.VAE_CONFIGS[name])
    d = P.draws(n)
    monkeypatch.delenv("L2HMC_LAYERED_FUSED_NET", raising=False)
    da = P.product()
    a = U.run_kernel_propose(P, d, dyn=da)
    la = da.launch_count
    monkeypatch.setenv("L2HMC_LAYERED_FUSED_NET", "0")
    db = P.product()
    b = U.run_kernel_propose(P, d, dyn=db)
    for k in ("Lx", "Lv", "px", "x_next"):
        assert np.array_equal(a[k], b[k]), k
    assert np.isfinite(a["Lx"]).all() and la < db.launch_count


def test_vae_log_jac_mode():
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])
    rep, _ = U.parity_report(P, 192, log_jac=True)
    assert rep["Lx_kernel"] <= max(SAMPLE_TOL, 4 * rep["Lx_o32"]), rep
    assert rep["px_kernel"] <= max(2e-4, 4 * rep["px_o32"]), rep  # px is log|J| here


def test_vae_components_match_oracle():
    """Dynamics.energy / grad_energy / hamiltonian / p_accept / net call with aux= (utils/dynamics.py:203-218,302)."""
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])
    dyn = P.product()
    n = 77
    d = P.draws(n, seed=3)
    o64, ae = P.oracle_for(d, torch.float64)
    g = lambda a: torch.as_tensor(np.asarray(a)).cuda()  # noqa: E731
    x, v, aux = g(d["x"]), g(d["v_f"]), g(d["aux"])
    x64, v64 = U.t64(d["x"]), U.t64(d["v_f"])
    e_ref = o64.energy(x64).numpy()
    assert U.max_rel(dyn.energy(x, aux=aux).cpu().numpy(), e_ref) <= 2e-6
    assert U.max_rel(dyn.grad_energy(x, aux=aux).cpu().numpy(), o64.grad_energy(x64).numpy()) <= 5e-6
    assert U.max_rel(dyn.hamiltonian(x, v, aux=aux).cpu().numpy(), o64.hamiltonian(x64, v64).numpy()) <= 2e-6
    x1, v1 = g(d["x"] * 0.9 + 0.05), g(d["v_b"])
    lj = g(0.01 * d["u"])
    p_ref = o64.p_accept(x64, v64, U.t64(d["x"] * 0.9 + 0.05), U.t64(d["v_b"]), U.t64(0.01 * d["u"])).numpy()
    assert np.max(np.abs(dyn.p_accept(x, v, x1, v1, lj, aux=aux).cpu().numpy() - p_ref)) <= P_TOL
    # the energy descriptor is callable like the reference's closure: energy(z, aux=aux)
    assert U.max_rel(dyn._fn(x, aux=aux).cpu().numpy(), e_ref) <= 2e-6
    # net call with the aux branch
    S, T, Q = dyn.net_apply("XNet", v, x, 2.0, aux=aux)
    tau = o64.format_time(2.0, n)
    Sr, Tr, Qr = U.O.net_apply(o64.xnet, v64, x64, tau, ae)
    for a, b in ((S, Sr), (T, Tr), (Q, Qr)):
        assert U.max_rel(a.cpu().numpy(), b.numpy()) <= 5e-6
    # aux is mandatory for this target
    with pytest.raises(ValueError):
        dyn.energy(x)
    with pytest.raises(ValueError):
        dyn.forward(x)


def test_vae_forward_backward_roundtrip_and_host_path():
    """backward(forward(x)) returns to x with log|J| cancelling (the reference's exact inverse), through aux-conditioned
    nets; and l2hmc_transition_host (host buffers incl. aux) equals the device path."""
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])
    dyn = P.product()
    d = P.draws(150, seed=5)
    g = lambda a: torch.as_tensor(np.asarray(a)).cuda()  # noqa: E731
    x, v, aux = g(d["x"]), g(d["v_f"]), g(d["aux"])
    X, V, j1 = dyn.forward(x, init_v=v, aux=aux, log_jac=True)
    x2, v2, j2 = dyn.backward(X, init_v=V, aux=aux, log_jac=True)
    assert U.max_rel(x2.cpu().numpy(), d["x"]) <= 2e-5
    assert U.max_rel(v2.cpu().numpy(), d["v_f"]) <= 2e-5
    assert float((j1 + j2).abs().max()) <= 2e-4
    o = dyn._transition(x, v=v, direction=g(d["dir"]), u=g(d["u"]), do_mh=True, aux=aux, counter=0)
    h = dyn.transition_host(d["x"], v=d["v_f"], direction=d["dir"], u=d["u"], do_mh=True, aux=d["aux"], counter=0)
    for k in ("Lx", "Lv", "px", "x_next"):
        np.testing.assert_array_equal(o[k].cpu().numpy(), h[k])


def test_vae_multi_transition_and_philox_sharding():
    """n_transitions > 1 on the layered engine equals the host loop; Philox keyed by global chain id makes a
    sharded run equal the unsharded one."""
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])
    dyn = P.product(seed=11)
    d = P.draws(130, seed=7)
    g = lambda a: torch.as_tensor(np.asarray(a)).cuda()  # noqa: E731
    x, aux = g(d["x"]), g(d["aux"])
    o3 = dyn._transition(x, dir_mode=3, do_mh=True, n_transitions=3, counter=5, aux=aux)
    cur = x
    for k in range(3):
        o1 = dyn._transition(cur, dir_mode=3, do_mh=True, counter=5 + k, aux=aux)
        cur = o1["x_next"]
    np.testing.assert_array_equal(o3["x_next"].cpu().numpy(), cur.cpu().numpy())
    np.testing.assert_array_equal(o3["px"].cpu().numpy(), o1["px"].cpu().numpy())
    lo = dyn._transition(x[:50], dir_mode=3, do_mh=True, counter=9, aux=aux[:50], chain_offset=0)
    hi = dyn._transition(x[50:], dir_mode=3, do_mh=True, counter=9, aux=aux[50:], chain_offset=50)
    al = dyn._transition(x, dir_mode=3, do_mh=True, counter=9, aux=aux)
    np.testing.assert_array_equal(torch.cat([lo["x_next"], hi["x_next"]]).cpu().numpy(), al["x_next"].cpu().numpy())


@pytest.mark.parametrize("name,n,regime", [
    ("c2_scg50", 200, "stress"),
    ("c3_mog2", 130, "stress"),
    ("c4_rw32_hard", 129, "stress"),
    ("funnel3", 100, "stress"),
])
def test_layered_engine_on_closed_form_targets(name, n, regime):
    """kernel='layered' forces the batched engine on shapes the fused kernels also cover."""
    P = U.Problem(regime=regime, **U.CONFIGS[name])
    dyn = P.product(kernel="layered")
    assert dyn.kernel_name.startswith("layered")
    rep, _ = U.parity_report(P, n, dyn=dyn)
    _check(rep)


def test_layered_hmc_mode():
    P = U.Problem(kind="gaussian", D=50, T=10, eps=0.05, hmc=True)
    rep, _ = U.parity_report(P, 200, dyn=P.product(kernel="layered"))
    _check(rep)


@pytest.mark.parametrize("D,H", [(80, 100), (50, 200), (130, 40)])
def test_shapes_beyond_the_fused_kernels(D, H):
    """x_dim > 64 or width > 128: AUTO must route to the layered engine (the fused kernels reject these)."""
    P = U.Problem(kind="gaussian", D=D, H=H, T=5, eps=0.05, regime="stress")
    dyn = P.product()
    assert dyn.kernel_name.startswith("layered")
    rep, _ = U.parity_report(P, 150, dyn=dyn)
    _check(rep)
    with pytest.raises(Exception):
        P.product(kernel="tile")


def test_vae_golden_fixture():
    """Committed fp64-oracle vectors for the miniature config-5 problem (tests/golden/make_golden.py)."""
    import os
    import golden_io
    P, d, gold = golden_io.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c5_vae_mini_n96.npz"))
    rk = U.run_kernel_propose(P, d)
    assert U.max_rel(rk["Lx"], gold["Lx"]) <= 4e-5
    assert U.max_rel(rk["Lv"], gold["Lv"]) <= 4e-5
    assert np.max(np.abs(rk["px"] - gold["px"])) <= P_TOL


# ---- device-side diagnostics (utils/func_utils.py:45-54,114-120) ------------------------------------------------
def test_acl_spectrum_and_ess_match_the_numpy_reference():
    from l2hmc_b200 import diagnostics as G
    rng = np.random.default_rng(3)
    for S, N, Dd in ((50, 37, 2), (33, 300, 50), (8, 5000, 7)):
        X = rng.standard_normal((S, N, Dd)).astype(np.float32).cumsum(0).astype(np.float32) * 0.1
        scale = float(np.sqrt(Dd) * 1.7)
        ref = U.O.acl_spectrum(X, scale)
        got = G.acl_spectrum(torch.as_tensor(X).cuda(), scale).cpu().numpy()
        assert got.shape == ref.shape
        # the reference's numpy keeps float32 products / per-step sums; the device sums in fp64
        assert np.max(np.abs(got - ref)) <= 2e-6 * max(1.0, np.abs(ref).max())
        assert abs(G.ESS(got) - U.O.ESS(ref)) <= 1e-5
        assert abs(G.autocovariance(torch.as_tensor(X).cuda(), 3) - U.O.autocovariance(X, 3)) <= 2e-6 * abs(U.O.autocovariance(X, 3))
    part = G.acl_spectrum(torch.as_tensor(X).cuda(), scale, n_lags=3).cpu().numpy()
    assert np.allclose(part, ref[:3], rtol=0, atol=2e-6)


def test_sample_trace_is_the_notebook_loop_on_device():
    """sample_trace == repeated propose(do_mh_step=True) (SCGExperiment.ipynb:291-298), and the L2HMC / HMC ESS comparison
    of the notebook (:331-334,388) runs end to end on the device."""
    from l2hmc_b200 import diagnostics as G, propose
    P = U.Problem(regime="init", **U.CONFIGS["c1_scg2"])
    dyn = P.product(seed=5)
    x0 = torch.as_tensor(P.x0(200, np.random.default_rng(1))).cuda()
    c0 = dyn._counter
    tr = G.sample_trace(x0, dyn, 6)
    dyn._counter = c0
    cur = x0
    for t in range(6):
        _, _, _, out = propose(cur, dyn, do_mh_step=True)
        cur = out[0]
        np.testing.assert_array_equal(tr[t].cpu().numpy(), cur.cpu().numpy())
    # the notebook's evaluation, shortened: spectra normalised by sqrt(trace(cov)), ESS of both samplers
    scale = float(np.sqrt(np.trace(U.scg2_cov())))
    A = G.acl_spectrum(G.sample_trace(x0, dyn, 300), scale)
    H = U.Problem(kind="gaussian", D=2, T=10, eps=0.15, hmc=True).product(seed=6)
    B = G.acl_spectrum(G.sample_trace(x0, H, 300), scale)
    assert A.shape == (299,) and 0.5 < float(A[0]) < 2.0 and 0.5 < float(B[0]) < 2.0
    assert 0.0 < G.ESS(A) <= 1.0 and 0.0 < G.ESS(B) <= 1.0


def test_ais_on_the_vae_posterior_matches_oracle():
    """eval_vae.py:49-65: annealing from the standard normal prior to the decoder posterior with HMC-mode Dynamics; the
    annealed energy is the decoder energy with the likelihood weighted by beta (l2hmc_set_likelihood_scale)."""
    from l2hmc_b200.ais import ais_estimate
    from l2hmc_b200.distributions import Gaussian
    P = U.VaeProblem(regime="stress", **U.VAE_CONFIGS["c5_vae_noenc"])
    n, steps, L, eps = 192, 6, 4, 0.08
    d = P.draws(n)
    rng = np.random.default_rng(7)
    r = {"v0": rng.standard_normal((n, P.D)).astype(np.float32), "v": rng.standard_normal((steps, n, P.D)).astype(np.float32),
         "u": rng.random((steps, n)).astype(np.float32)}
    e1 = U.O.DecoderBernoulliEnergy(P.dec_W, P.dec_b, d["aux"], torch.float64)
    e0 = U.O.GaussianEnergy(np.zeros(P.D), np.eye(P.D))
    est_o, alpha_o, x_o, w_o = U.O.ais_estimate(e0, e1, steps, d["x"], step_size=eps, leapfrogs=L, v0=r["v0"], v_refresh=r["v"],
                                                u=r["u"], num_splits=3)
    dyn = P.product()                                  # only to get the DecoderEnergy descriptor of this problem
    final_energy = dyn._fn
    prior = Gaussian(np.zeros(P.D), np.eye(P.D)).get_energy_function()
    est, alpha, x, w = ais_estimate(prior, final_energy, steps, torch.as_tensor(d["x"]).cuda(), aux=torch.as_tensor(d["aux"]).cuda(),
                                    step_size=eps, leapfrogs=L, x_dim=P.D, num_splits=3, rng=r, return_state=True)
    same = np.abs(x.cpu().numpy() - x_o.numpy()).max(1) < 1e-3
    assert same.mean() > 0.98
    assert U.max_rel(x.cpu().numpy()[same], x_o.numpy()[same]) <= 5e-5
    assert float(np.abs(w.cpu().numpy()[same] - w_o.numpy()[same]).max()) <= 5e-4 * max(1.0, float(np.abs(w_o.numpy()).max()))
    assert abs(alpha - float(alpha_o)) <= 2e-4
    # the weight changes what the sampler's energy calls return: put it back
    dyn.set_likelihood_scale(1.0)
