"""GPU parity against the reference's OWN outputs (tests/golden/ref_*.npz, tests/golden/ref/*.npz: the unmodified
/root/reference sources run on oracle/tf_shim, see tests/golden/make_ref_golden.py).  Nothing here executes the oracle:
the CUDA path (Python mirror -> ctypes -> C ABI) is compared with committed reference vectors.

Tolerances -- BASELINE.json north_star: "samples / accept-probs matching the reference TF1 CPU path on identical seeds
within 1e-5 relative fp32 tolerance":
  * mean accept probability within 1e-5 of the reference's fp64 run;
  * samples Lx, Lv (relative to max|ref|): within REF32 + 1e-5, where REF32 is the error of the reference's own
    float32 run (`out32_*`) against its float64 run on the same inputs -- the fp32 path is only defined up to that
    rounding noise, so an implementation cannot be asked to land closer to fp64 than 1e-5 beyond it;
  * per-chain accept probability (absolute): within 4 * REF32 + 1e-5 -- the maximum over chains of a cancellation error
    (fp32 Hamiltonians of O(100) subtracted), which differs by up to ~4x between two evaluation orders of the same
    arithmetic while the mean agrees to 1e-6 (profiles/r02_parity_noise.txt).
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

import util as U

pytestmark = pytest.mark.gpu

NORTH_STAR_TOL = 1e-5
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROPOSE_FILES = sorted(glob.glob(os.path.join(GOLD, "ref_*.npz")))
_ids = [os.path.basename(f)[:-4] for f in PROPOSE_FILES]


def _meta(z):
    return json.loads(bytes(z["meta"]).decode())


def _check_against_reference(rk, z, what):
    rep = {}
    for k in ("Lx", "Lv"):
        ref32 = U.max_rel(z["out32_" + k], z["out_" + k])
        rep[k] = (U.max_rel(rk[k], z["out_" + k]), ref32)
        assert rep[k][0] <= ref32 + NORTH_STAR_TOL, (what, k, rep)
    ok = np.isfinite(z["out_px"])
    ref32 = float(np.max(np.abs(z["out32_px"][ok] - z["out_px"][ok])))
    err = float(np.max(np.abs(rk["px"][ok] - z["out_px"][ok])))
    rep["px"] = (err, ref32)
    assert err <= 4 * ref32 + NORTH_STAR_TOL, (what, "px", rep)
    mean_err = abs(float(rk["px"][ok].astype(np.float64).mean()) - float(z["out_px"][ok].mean()))
    rep["px_mean"] = mean_err
    assert mean_err <= NORTH_STAR_TOL, (what, "px_mean", rep)
    # Metropolis decisions identical except where |p - u| is inside the fp32 noise
    acc_ref = np.all(z["out_x_next"] == z["out_Lx"], axis=1)
    acc_k = np.all(rk["x_next"] == rk["Lx"], axis=1)
    margin = np.abs(z["out_px"] - z["in_u"])
    assert int(((acc_ref != acc_k) & (margin > 1e-4)).sum()) == 0, (what, "accept decisions")
    return rep


@pytest.mark.parametrize("path", PROPOSE_FILES, ids=_ids)
def test_propose_matches_the_reference(path):
    """propose(x, dynamics, do_mh_step=True[, log_jac]) -- utils/sampler.py:28-55 over utils/dynamics.py:246-309."""
    import golden_io
    P, d, _ = golden_io.load(path)
    z = np.load(path)
    rk = U.run_kernel_propose(P, d, log_jac=P.meta["log_jac"])
    if P.meta["log_jac"]:
        # px holds log|J| (sampler.py:44 with log_jac=True): relative figure, no accept decision to compare
        for k in ("Lx", "Lv", "px"):
            ref32 = U.max_rel(z["out32_" + k], z["out_" + k])
            assert U.max_rel(rk[k], z["out_" + k]) <= ref32 + NORTH_STAR_TOL, (path, k)
        return
    rep = _check_against_reference(rk, z, os.path.basename(path))
    print("reference pin", os.path.basename(path), {k: v for k, v in rep.items()})


@pytest.mark.parametrize("kernel", ["tile", "tc", "layered"])
def test_every_engine_matches_the_reference_on_config2(kernel):
    """BASELINE config 2's shape through each engine that can run it (FMA tile, tcgen05 fused, layered GEMMs)."""
    import golden_io
    path = os.path.join(GOLD, "ref_c2_scg50_n128_stress.npz")
    P, d, _ = golden_io.load(path)
    z = np.load(path)
    rk = U.run_kernel_propose(P, d, dyn=P.product(kernel=kernel))
    _check_against_reference(rk, z, "c2/" + kernel)


@pytest.mark.parametrize("path", [f for f in PROPOSE_FILES if "logjac" not in f],
                         ids=[i for i in _ids if "logjac" not in i])
def test_dynamics_methods_match_the_reference(path):
    """Dynamics.energy / grad_energy / kinetic / hamiltonian / forward / backward / p_accept (utils/dynamics.py:107-108,
    203-218, 246-309) and one raw S/T/Q call per net against the reference's outputs for the same method calls."""
    import golden_io
    P, d, _ = golden_io.load(path)
    z = np.load(path)
    m = {k[2:]: z[k] for k in z.files if k.startswith("m_")}
    m32 = {k[4:]: z[k] for k in z.files if k.startswith("m32_")}
    dyn = P.product()
    x, v = torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["v_f"]).cuda()

    def close(a, key, absolute=False):
        """within REF32 + 1e-5 of the reference's float64 output (REF32: its own float32 run on the same call)"""
        a = a.cpu().numpy()
        if absolute:
            err, ref32 = float(np.max(np.abs(a - m[key]))), float(np.max(np.abs(m32[key] - m[key])))
        else:
            err, ref32 = U.max_rel(a, m[key]), U.max_rel(m32[key], m[key])
        assert err <= (4 * ref32 if absolute else ref32) + NORTH_STAR_TOL, (path, key, err, ref32)
    close(dyn.energy(x), "energy")
    close(dyn.grad_energy(x), "grad_energy")
    close(dyn.kinetic(v), "kinetic")
    close(dyn.hamiltonian(x, v), "hamiltonian")
    X, V, lj = dyn.forward(x, init_v=v, log_jac=True)
    close(X, "fwd_x"); close(V, "fwd_v"); close(lj, "fwd_logjac")
    Xb, Vb, ljb = dyn.backward(x, init_v=v, log_jac=True)
    close(Xb, "bwd_x"); close(Vb, "bwd_v"); close(ljb, "bwd_logjac")
    close(dyn.forward(x, init_v=v)[2], "fwd_p", absolute=True)
    close(dyn.backward(x, init_v=v)[2], "bwd_p", absolute=True)
    f32c = lambda a: torch.as_tensor(a).float().cuda()  # noqa: E731
    close(dyn.p_accept(x, v, f32c(m["fwd_x"]), f32c(m["fwd_v"]), f32c(m["fwd_logjac"])), "p_accept", absolute=True)
    if not P.hmc:
        mk = torch.as_tensor(P.mask[1]).cuda()
        for key, val in zip(("vnet_S", "vnet_T", "vnet_Q"), dyn.net_apply("VNet", x, f32c(m["grad_energy"]), 1.0)):
            close(val, key)
        for key, val in zip(("xnet_S", "xnet_T", "xnet_Q"), dyn.net_apply("XNet", v, mk * x, 1.0)):
            close(val, key)


@pytest.mark.parametrize("name,kernel,launches", [("chain_operator_c1_n64", "auto", 1), ("chain_operator_c1_n64", "tile", 1),
                                                  ("chain_operator_c2_n32", "auto", 1), ("chain_operator_c2_n32", "tile", 1),
                                                  ("chain_operator_c2_n32", "layered", None)])
def test_chain_operator_matches_the_reference(name, kernel, launches):
    """utils/sampler.py:57-85 as the reference runs it -- in ONE kernel launch on the fused kernels (small, tile, the
    shape-specialised tensor-core kernel: l2hmc_transition_args.chain), as a loop of launches on the layered engine."""
    from l2hmc_b200 import chain_operator
    z = np.load(os.path.join(GOLD, "ref", name + ".npz"))
    meta = _meta(z)
    P = U.Problem(regime=meta["regime"], **meta["kw"])
    P.mask = z["mask"]
    P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
    P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
    g = lambda a: torch.as_tensor(np.asarray(a)).cuda()  # noqa: E731
    steps = meta["nb_steps"]
    rngs = []
    for s in range(steps):
        sel = np.where(z["in_dirs"][s][:, None] != 0, z["in_v_fs"][s], z["in_v_bs"][s]).astype(np.float32)
        rngs.append({"direction": g(z["in_dirs"][s]), "v": g(sel)})
    rngs.append({"u": g(z["in_u"])})
    dyn = P.product(kernel=kernel)
    dyn._ensure_ctx()
    l0 = dyn.launch_count
    fx, fv, p, outs = chain_operator(g(z["in_x"]), dyn, steps, init_v=g(z["in_init_v"]), do_mh_step=True, rng=rngs)
    if launches is not None:
        assert dyn.launch_count - l0 == launches, (dyn.kernel_name, dyn.launch_count - l0)
    assert U.max_rel(fx.cpu().numpy(), z["out_final_x"]) <= 3e-5
    assert U.max_rel(fv.cpu().numpy(), z["out_final_v"]) <= 3e-5
    assert float(np.max(np.abs(p.cpu().numpy() - z["out_p_accept"]))) <= 2e-4
    acc_ref = np.all(z["out_x_next"] == z["out_final_x"], axis=1)
    acc_k = np.all(outs[0].cpu().numpy() == fx.cpu().numpy(), axis=1)
    margin = np.abs(z["out_p_accept"] - z["in_u"])
    assert int(((acc_ref != acc_k) & (margin > 1e-3)).sum()) == 0
    # rejected chains keep init_x
    xn = outs[0].cpu().numpy()
    assert np.array_equal(xn[~acc_k], z["in_x"][~acc_k])


def test_chain_operator_with_in_kernel_randomness_is_one_launch_and_self_consistent():
    """No injected randomness: directions, momenta, init_v and the closing uniforms come from the Philox stream inside the
    one launch; re-injecting exactly those draws (l2hmc_philox_fill at the same counters) reproduces the result."""
    from l2hmc_b200 import chain_operator
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.product(seed=21)
    n, K = 300, 3
    x = torch.as_tensor(P.x0(n, np.random.default_rng(4))).cuda()
    dyn._ensure_ctx()
    c0, l0 = dyn._counter, dyn.launch_count
    fx, fv, p, outs = chain_operator(x, dyn, K, do_mh_step=True)
    assert dyn.launch_count - l0 == 1 and dyn._counter == c0 + K
    import ctypes as C
    vs, ds = [], []
    for s in range(K + 1):   # counters c0 .. c0+K-1: the sub-proposals; c0+K: init_v
        v = torch.empty((n, P.D), device="cuda")
        d = torch.empty((n,), dtype=torch.uint8, device="cuda")
        u = torch.empty((n,), device="cuda")
        dyn._chk(dyn._lib.l2hmc_philox_fill(dyn._ctx, n, 0, dyn.seed, c0 + s, v.data_ptr(), d.data_ptr(), u.data_ptr(), dyn._stream()))
        vs.append(v); ds.append(d)
        if s == K - 1:
            u_last = u
    rngs = [{"direction": ds[s], "v": vs[s]} for s in range(K)] + [{"u": u_last}]
    fx2, fv2, p2, outs2 = chain_operator(x, dyn, K, init_v=vs[K], do_mh_step=True, rng=rngs)
    assert torch.equal(fx, fx2) and torch.equal(fv, fv2) and torch.equal(p, p2) and torch.equal(outs[0], outs2[0])


def _load_vae(name):
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
    assert abs(U.vae_weight_checksum(P) - meta["weight_checksum"]) <= 1e-9 * abs(meta["weight_checksum"])
    return P, {k[3:]: z[k] for k in z.files if k.startswith("in_")}, z


@pytest.mark.parametrize("name", ["c5_vae_mini_n96", "c5_vae_full_n32"])
def test_vae_target_matches_the_reference(name):
    """BASELINE config 5: propose(init_x, dynamics, aux=inp, do_mh_step=True) (mnist_vae.py:204) on the decoder-Bernoulli
    posterior with aux-conditioned nets; c5_vae_full is mnist_vae.py's own text (:104-111,122-126,130-178) at its own
    layer sizes (50 -> 1024 -> 1024 -> 784 decoder, 784 -> 512 -> 512 -> 200 encoder, width-200 nets, Lf=15)."""
    P, d, z = _load_vae(name)
    dyn = P.product()
    x, aux = torch.as_tensor(d["x"]).cuda(), torch.as_tensor(d["aux"]).cuda()
    e = dyn.energy(x, aux=aux).cpu().numpy()
    assert U.max_rel(e, z["out_energy"]) <= 1e-5
    gr = dyn.grad_energy(x, aux=aux).cpu().numpy()
    assert U.max_rel(gr, z["out_grad_energy"]) <= 2e-5
    rk = U.run_kernel_propose(P, d, dyn=dyn)
    for k in ("Lx", "Lv"):
        ref32 = U.max_rel(z["out32_" + k], z["out_" + k])
        assert U.max_rel(rk[k], z["out_" + k]) <= ref32 + NORTH_STAR_TOL, (name, k, U.max_rel(rk[k], z["out_" + k]), ref32)
    # U is a sum over 784 pixels (O(500)): the Hamiltonian difference carries fp32 noise of ~1e-4 in the reference's own
    # float32 run, which is what bounds the per-chain accept probability here
    ref32 = float(np.max(np.abs(z["out32_px"] - z["out_px"])))
    assert float(np.max(np.abs(rk["px"] - z["out_px"]))) <= 4 * ref32 + NORTH_STAR_TOL, (name, ref32)
    assert abs(float(rk["px"].astype(np.float64).mean()) - float(z["out_px"].mean())) <= max(NORTH_STAR_TOL, ref32 / 4)


def test_ais_matches_the_reference():
    """utils/ais.py:30-82 between two Gaussians, every particle update on the GPU."""
    from l2hmc_b200.ais import ais_estimate
    from l2hmc_b200.distributions import Gaussian
    z = np.load(os.path.join(GOLD, "ref", "ais_gauss3.npz"))
    meta = _meta(z)
    D = meta["D"]
    g0, g1 = Gaussian(np.zeros(D), np.eye(D)), Gaussian(z["mu1"], z["cov1"])
    r = {"v0": z["in_v0"], "v": z["in_v_refresh"], "u": z["in_u"]}
    est, alpha = ais_estimate(g0.get_energy_function(), g1.get_energy_function(), meta["anneal_steps"],
                              torch.as_tensor(z["in_x"]).cuda(), step_size=meta["step_size"], leapfrogs=meta["leapfrogs"],
                              x_dim=D, rng=r)
    assert abs(float(alpha) - float(z["out_mean_accept"])) <= 1e-4
    assert abs(float(est) - float(z["out_estimate"])) <= 2e-3  # a flipped Metropolis decision moves one of 64 weights


def test_ais_between_unlike_energies_matches_the_reference():
    """utils/ais.py:44-45 for a pair whose mixture is not a closed form (Gaussian -> rough well): the annealed energy
    (1 - beta) U0 + beta U1 is evaluated per chain inside the fused kernel (l2hmc_set_energy_mixed)."""
    from l2hmc_b200.ais import ais_estimate
    from l2hmc_b200.distributions import Gaussian, RoughWell
    z = np.load(os.path.join(GOLD, "ref", "ais_roughwell4.npz"))
    meta = _meta(z)
    D = meta["D"]
    g0, g1 = Gaussian(np.zeros(D), z["cov0"]), RoughWell(D, meta["rw_eps"], easy=meta["easy"])
    r = {"v0": z["in_v0"], "v": z["in_v_refresh"], "u": z["in_u"]}
    est, alpha = ais_estimate(g0.get_energy_function(), g1.get_energy_function(), meta["anneal_steps"],
                              torch.as_tensor(z["in_x"]).cuda(), step_size=meta["step_size"], leapfrogs=meta["leapfrogs"],
                              x_dim=D, rng=r)
    assert abs(float(alpha) - float(z["out_mean_accept"])) <= 1e-4
    assert abs(float(est) - float(z["out_estimate"])) <= 2e-3


@pytest.mark.parametrize("D,final", [(2, "gmm"), (6, "gmm"), (3, "funnel"), (12, "roughwell")])
def test_mixed_energy_components_match_the_oracle(D, final):
    """distributions.MixedEnergy on the device (energy, grad U, an HMC-mode transition; small kernel for x_dim <= 4, tile
    kernel above) against the oracle's MixedEnergy (utils/ais.py:44-45)."""
    from l2hmc_b200 import Dynamics
    from l2hmc_b200.distributions import Gaussian, GMM, GaussianFunnel, RoughWell, MixedEnergy
    rng = np.random.default_rng(D)
    cov0 = np.eye(D) * 1.3
    g0 = Gaussian(np.zeros(D), cov0)
    e0 = U.O.GaussianEnergy(np.zeros(D), np.linalg.inv(cov0).astype(np.float32))
    if final == "gmm":
        mus = [rng.standard_normal(D), rng.standard_normal(D)]
        sig = [0.5 * np.eye(D), 0.8 * np.eye(D)]
        g1 = GMM(mus, sig, [0.5, 0.5])
        e1 = U.O.GMMEnergy(mus, g1.i_sigmas, g1.constants)
    elif final == "funnel":
        g1 = GaussianFunnel(dim=D)
        e1 = U.O.FunnelEnergy(g1.sigma, g1.clip)
    else:
        g1 = RoughWell(D, 0.4, easy=True)
        e1 = U.O.RoughWellEnergy(0.4, True)
    beta = 0.35
    mixed = MixedEnergy(g0.get_energy_function(), g1.get_energy_function(), beta)
    om = U.O.MixedEnergy(e0, e1, beta).to(torch.float64)
    n = 200
    x = rng.standard_normal((n, D)).astype(np.float32)
    v = rng.standard_normal((n, D)).astype(np.float32)
    dyn = Dynamics(D, mixed, T=6, eps=0.15, hmc=True)
    xt, vt = torch.as_tensor(x).cuda(), torch.as_tensor(v).cuda()
    assert U.max_rel(dyn.energy(xt).cpu().numpy(), om.energy(U.t64(x)).numpy()) <= 1e-5
    assert U.max_rel(dyn.grad_energy(xt).cpu().numpy(), om.grad(U.t64(x)).numpy()) <= 1e-5
    Lx, Lv, px = dyn.forward(xt, init_v=vt)
    od = U.O.OracleDynamics(D, 6, 0.15, om, np.zeros((6, D), np.float32), hmc=True, dtype=torch.float64)
    ox, ov, op = od.forward(U.t64(x), U.t64(v))
    assert U.max_rel(Lx.cpu().numpy(), ox.numpy()) <= 2e-5 and U.max_rel(Lv.cpu().numpy(), ov.numpy()) <= 2e-5
    assert float(np.max(np.abs(px.cpu().numpy() - op.numpy()))) <= 1e-4
    dyn.set_mix_beta(0.8)   # beta moves without re-sending the parameters
    om2 = U.O.MixedEnergy(e0, e1, 0.8).to(torch.float64)
    assert U.max_rel(dyn.energy(xt).cpu().numpy(), om2.energy(U.t64(x)).numpy()) <= 1e-5


def test_diagnostics_match_the_reference():
    """utils/func_utils.py:45-54,114-120 (numpy in the reference) on the device."""
    from l2hmc_b200 import diagnostics as G
    z = np.load(os.path.join(GOLD, "ref", "losses_diagnostics.npz"))
    X = torch.as_tensor(z["in_trace"]).cuda()
    spec = G.acl_spectrum(X, float(z["in_scale"])).cpu().numpy()
    assert np.max(np.abs(spec - z["acl_spectrum"])) <= 2e-6 * np.abs(z["acl_spectrum"]).max()
    assert abs(G.ESS(spec) - float(z["ess"])) <= 1e-5
    assert abs(G.autocovariance(X, 3) - float(z["autocov_3"])) <= 2e-6 * abs(float(z["autocov_3"]))
