"""CPU checks of the oracle itself (no GPU).

The reference ships no tests or golden vectors, so the oracle is pinned by the
properties its code implies, by the committed fp64 fixtures, and by the independent C restatement.
"""
import glob
import os

import numpy as np
import pytest
import torch

import util as U

O = U.O


@pytest.mark.parametrize("name", ["c1_scg2", "c2_scg50", "c3_mog2", "c4_rw32", "c4_rw32_hard", "funnel3"])
def test_backward_inverts_forward(name):
    """_backward_step is the algebraic inverse of _forward_step with steps reversed
    (utils/dynamics.py:159-201 vs :115-157, :285) and log|J| cancels (:166,178,186,195)."""
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    dyn = P.oracle(torch.float64)
    d = P.draws(32)
    x, v = U.t64(d["x"]), U.t64(d["v_f"])
    X, V, j = dyn.forward(x, v, log_jac=True)
    x2, v2, j2 = dyn.backward(X, V, log_jac=True)
    assert float((x2 - x).abs().max()) < 1e-7 * max(1, float(x.abs().max()))  # hard rough well amplifies fp64 rounding by ~1/eps^3
    assert float((v2 - v).abs().max()) < 1e-7 * max(1, float(X.abs().max()))
    assert float((j + j2).abs().max()) < 1e-8
    assert float(j.abs().max()) > 1e-3  # the stress regime actually exercises log|J|


def test_logjac_equals_autograd_logdet():
    """log|J| == log|det d(x',v')/d(x,v)| (utils/dynamics.py:155; utils/func_utils.py:56-57 is the
    reference's helper for exactly this check)."""
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    dyn = P.oracle(torch.float64)
    d = P.draws(4)
    for i in range(4):
        z0 = torch.cat([U.t64(d["x"][i]), U.t64(d["v_f"][i])])

        def f(z):
            X, V, _ = dyn.forward(z[None, :2], z[None, 2:], log_jac=True)
            return torch.cat([X[0], V[0]])
        J = torch.autograd.functional.jacobian(f, z0)
        _, _, lj = dyn.forward(z0[None, :2], z0[None, 2:], log_jac=True)
        assert abs(float(torch.linalg.slogdet(J)[1]) - float(lj[0])) < 1e-8
    # and for one backward trajectory
    z0 = torch.cat([U.t64(d["x"][0]), U.t64(d["v_b"][0])])

    def fb(z):
        X, V, _ = dyn.backward(z[None, :2], z[None, 2:], log_jac=True)
        return torch.cat([X[0], V[0]])
    J = torch.autograd.functional.jacobian(fb, z0)
    _, _, lj = dyn.backward(z0[None, :2], z0[None, 2:], log_jac=True)
    assert abs(float(torch.linalg.slogdet(J)[1]) - float(lj[0])) < 1e-8


def test_hmc_limit_is_plain_leapfrog():
    """Zero nets => classic leapfrog, log|J| = 0 (utils/dynamics.py:73-76)."""
    P = U.Problem(kind="gaussian", D=2, T=7, eps=0.05, hmc=True)
    dyn = P.oracle(torch.float64)
    d = P.draws(16)
    x, v = U.t64(d["x"]), U.t64(d["v_f"])
    X, V, j = dyn.forward(x, v, log_jac=True)
    assert float(j.abs().max()) == 0.0
    eps = float(dyn._eps)
    S = P.energy.to(torch.float64).S
    xr, vr = x.clone(), v.clone()
    for _ in range(7):  # masks only split the position update in two halves that sum to a full one
        vr = vr - 0.5 * eps * (xr @ S)
        xr = xr + eps * vr
        vr = vr - 0.5 * eps * (xr @ S)
    assert float((X - xr).abs().max()) < 1e-10 and float((V - vr).abs().max()) < 1e-10


@pytest.mark.parametrize("name", ["c1_scg2", "c3_mog2", "c4_rw32_hard", "funnel3"])
def test_analytic_gradients_match_autodiff(name):
    """grad() must be what tf.gradients(energy(x), x) yields (utils/dynamics.py:217-218)."""
    P = U.Problem(**U.CONFIGS[name])
    en = P.energy.to(torch.float64)
    x = U.t64(P.draws(64)["x"])
    if name == "funnel3":  # also cover both clipped branches (utils/distributions.py:176-177)
        x[0, 0], x[1, 0] = 9.0, -9.5
    assert float((en.grad(x) - en.grad_autodiff(x)).abs().max()) < 1e-9


def test_accept_prob_range_and_nonfinite():
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    dyn = P.oracle(torch.float32)
    d = P.draws(64)
    _, _, p = dyn.forward(torch.as_tensor(d["x"]), torch.as_tensor(d["v_f"]))
    assert float(p.min()) >= 0 and float(p.max()) <= 1
    z = torch.zeros((2, 2))
    bad = torch.tensor([[float("inf"), 0.0], [float("nan"), 0.0]])
    assert dyn.p_accept(z, z, bad, z, torch.zeros(2)).tolist() == [0.0, 0.0]  # utils/dynamics.py:309


def test_selected_direction_equals_blend():
    """Running only the selected direction (what the fused kernel does) equals the reference's
    compute-both-and-blend (utils/sampler.py:35-44) whenever the discarded branch is finite."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.oracle(torch.float64)
    d = P.draws(48)
    Lx, Lv, px, _ = O.propose(U.t64(d["x"]), dyn, direction=torch.as_tensor(d["dir"].astype(np.float32)),
                              v_f=U.t64(d["v_f"]), v_b=U.t64(d["v_b"]), init_v=U.t64(d["v_f"]))
    sel = np.where(d["dir"][:, None] != 0, d["v_f"], d["v_b"])
    Sx, Sv, sp = O.propose_selected(U.t64(d["x"]), dyn, direction=torch.as_tensor(d["dir"]), v=U.t64(sel))
    assert torch.equal(Lx, Sx) and torch.equal(Lv, Sv) and torch.equal(px, sp)


def test_fp32_twin_noise_floor():
    """Documents the fp32 reordering noise the GPU tolerances are built on."""
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    d = P.draws(128)
    r64 = U.run_oracle_propose(P, d, torch.float64)
    r32 = U.run_oracle_propose(P, d, torch.float32)
    assert U.max_rel(r32["Lx"], r64["Lx"]) < 2e-5
    assert np.max(np.abs(r32["px"] - r64["px"])) < 2e-4


def test_golden_fixtures_reproduce():
    import golden_io
    files = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                   if not os.path.basename(f).startswith("ref_"))  # ref_*: test_reference_pin.py / test_gpu_reference_pin.py
    assert len(files) >= 6
    for f in files:
        P, d, ref = golden_io.load(f)
        got = U.run_oracle_propose(P, d, torch.float64, P.meta["log_jac"])
        for k in ("Lx", "Lv", "px", "x_next"):
            assert np.allclose(got[k], ref[k], rtol=0, atol=1e-9 * max(1.0, np.abs(ref[k]).max())), (f, k)


@pytest.mark.parametrize("name", list(U.CONFIGS))
def test_c_restatement_agrees_with_torch_oracle(name):
    """Two independently written restatements (torch op-for-op, plain C scalar loops) must agree:
    fp64 twins to rounding, fp32 twins each within the fp32 noise floor of the fp64 result."""
    import c_oracle
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    d = P.draws(48)
    co = c_oracle.COracle(P)
    c64, t64 = co.propose(d, np.float64), U.run_oracle_propose(P, d, torch.float64)
    assert U.max_rel(c64["Lx"], t64["Lx"]) < 1e-7 and U.max_rel(c64["Lv"], t64["Lv"]) < 1e-7
    assert np.max(np.abs(c64["px"] - t64["px"])) < 1e-5
    c32 = co.propose(d, np.float32)
    assert U.max_rel(c32["Lx"], t64["Lx"]) < 2e-5
    assert np.max(np.abs(c32["px"] - t64["px"])) < 2e-4
    lj_c = co.propose(d, np.float64, log_jac=True)["px"]
    lj_t = U.run_oracle_propose(P, d, torch.float64, log_jac=True)["px"]
    assert np.max(np.abs(lj_c - lj_t)) < 1e-6


def test_c_restatement_hmc():
    import c_oracle
    P = U.Problem(kind="gaussian", D=2, T=10, eps=0.15, hmc=True)
    d = P.draws(64)
    c64 = c_oracle.COracle(P).propose(d, np.float64)
    t64 = U.run_oracle_propose(P, d, torch.float64)
    assert U.max_rel(c64["Lx"], t64["Lx"]) < 1e-10 and np.max(np.abs(c64["px"] - t64["px"])) < 1e-10


def test_statistical_hmc_recovers_covariance():
    """HMC-mode chains on BASELINE config 1's target recover its covariance (sanity of the whole
    transition: energy, leapfrog, accept)."""
    P = U.Problem(kind="gaussian", D=2, T=10, eps=0.15, hmc=True)
    dyn = P.oracle(torch.float32)
    rng = np.random.default_rng(0)
    n = 512
    x = torch.as_tensor(P.x0(n, rng))
    acc = []
    xs = []
    for t in range(150):
        v = torch.as_tensor(rng.standard_normal((n, 2)).astype(np.float32))
        u = torch.as_tensor(rng.random(n).astype(np.float32))
        Lx, _, px, outs = O.propose(x, dyn, init_v=v, u=u, do_mh_step=True)
        x = outs[0]
        acc.append(float(px.mean()))
        if t >= 50:
            xs.append(x.numpy().copy())
    X = np.concatenate(xs)
    cov = np.cov(X.T)
    assert np.allclose(cov, U.scg2_cov(), rtol=0.25, atol=3.0), cov
    assert 0.5 < np.mean(acc) <= 1.0


def test_untrained_accept_rate_near_notebook():
    """Untrained nets on config 1: the notebook logged a mean accept-prob of 0.90 at step 0
    (SCGExperiment.ipynb:200; unseeded there, so only the neighbourhood is checked)."""
    vals = []
    for s in range(4):
        P = U.Problem(seed=s, **U.CONFIGS["c1_scg2"])
        d = P.draws(200, seed=s + 100)
        d["x"] = np.random.default_rng(s).standard_normal((200, 2)).astype(np.float32)  # notebook :257
        vals.append(float(U.run_oracle_propose(P, d, torch.float32)["px"].mean()))
    assert 0.75 < np.mean(vals) <= 1.0, vals


# ---- BASELINE config 5: decoder-Bernoulli target + aux-conditioned nets (mnist_vae.py:104-178) ------------
@pytest.mark.parametrize("name", ["c5_vae_mini", "c5_vae_ragged"])
def test_vae_energy_gradient_is_reverse_mode_of_energy(name):
    """Dynamics.grad_energy is tf.gradients(energy(x, aux), x) (utils/dynamics.py:217-218): the written-out reverse
    pass of the decoder energy must equal autograd through energy()."""
    P = U.VaeProblem(**U.VAE_CONFIGS[name])
    d = P.draws(33)
    dyn, _ = P.oracle_for(d, torch.float64)
    x = U.t64(d["x"])
    g = dyn.energy_obj.grad(x)
    assert torch.allclose(g, dyn.energy_obj.grad_autodiff(x), rtol=0, atol=1e-12)
    # sigmoid_cross_entropy_with_logits(labels=z, logits=l) == -(z log s(l) + (1-z) log(1-s(l)))
    l = U.O.softplus_mlp(dyn.energy_obj.Ws, dyn.energy_obj.bs, x)
    a = dyn.energy_obj.aux
    bce = -(a * torch.nn.functional.logsigmoid(l) + (1 - a) * torch.nn.functional.logsigmoid(-l)).sum(1)
    assert torch.allclose(dyn.energy(x), bce + 0.5 * (x * x).sum(1), rtol=0, atol=1e-10)


def test_vae_backward_inverts_forward_with_aux_branch():
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])
    d = P.draws(40)
    dyn, ae = P.oracle_for(d, torch.float64)
    assert ae is not None and ae.shape == (40, P.H)
    x, v = U.t64(d["x"]), U.t64(d["v_f"])
    X, V, j1 = dyn.forward(x, v, log_jac=True, ae_x=ae, ae_v=ae)
    x2, v2, j2 = dyn.backward(X, V, log_jac=True, ae_x=ae, ae_v=ae)
    assert float((x2 - x).abs().max()) < 1e-9 and float((v2 - v).abs().max()) < 1e-9
    assert float((j1 + j2).abs().max()) < 1e-9
    # the aux branch matters: dropping it changes the proposal
    X0, _, _ = dyn.forward(x, v, log_jac=True)
    assert float((X0 - X).abs().max()) > 1e-4


def test_vae_selected_direction_equals_blend():
    P = U.VaeProblem(**U.VAE_CONFIGS["c5_vae_mini"])
    d = P.draws(50)
    dyn, ae = P.oracle_for(d, torch.float64)
    ref = U.run_oracle_propose(P, d, torch.float64)
    v_sel = np.where(d["dir"][:, None] != 0, d["v_f"], d["v_b"])
    Lx, Lv, px = U.O.propose_selected(U.t64(d["x"]), dyn, direction=torch.as_tensor(d["dir"]), v=U.t64(v_sel),
                                      ae_x=ae, ae_v=ae)
    assert np.allclose(Lx.numpy(), ref["Lx"], atol=1e-12) and np.allclose(px.numpy(), ref["px"], atol=1e-12)


# ---- diagnostics (utils/func_utils.py:45-54,114-120) ---------------------------------------------------------
def test_diagnostics_on_a_known_process():
    """AR(1) chains x_t = a x_{t-1} + sqrt(1-a^2) n_t have autocovariance d * a^tau per chain; ESS follows."""
    rng = np.random.default_rng(0)
    S, N, Dd, a = 400, 3000, 2, 0.8
    X = np.empty((S, N, Dd))
    X[0] = rng.standard_normal((N, Dd))
    for t in range(1, S):
        X[t] = a * X[t - 1] + np.sqrt(1 - a * a) * rng.standard_normal((N, Dd))
    X = X.astype(np.float32)
    assert abs(U.O.autocovariance(X, 0) - Dd) < 0.02 and abs(U.O.autocovariance(X, 3) - Dd * a ** 3) < 0.02
    A = U.O.acl_spectrum(X, np.sqrt(Dd))
    assert A.shape == (S - 1,) and abs(A[0] - 1.0) < 0.01 and abs(A[5] - a ** 5) < 0.01
    ess = U.O.ESS(A)
    lags = np.arange(1, 60)
    expect = 1.0 / (1.0 + 2.0 * np.sum((a ** lags)[a ** lags > 0.05]))
    assert abs(ess - expect) / expect < 0.05


def _ais_problem(D=3, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((D, D))
    cov1 = A @ A.T / D + 0.3 * np.eye(D)
    mu1 = rng.standard_normal(D) * 0.5
    S0 = np.eye(D)
    S1 = np.linalg.inv(cov1)
    e0 = U.O.GaussianEnergy(np.zeros(D), S0)
    e1 = U.O.GaussianEnergy(mu1, S1)
    log_ratio = -0.5 * (np.linalg.slogdet(S1)[1] - np.linalg.slogdet(S0)[1])  # log Z1 / Z0 of exp(-U)
    return e0, e1, mu1, cov1, log_ratio


def test_ais_oracle_recovers_the_normaliser_ratio():
    """utils/ais.py:30-82 restated: annealing N(0, I) -> N(mu1, cov1) estimates log Z1/Z0 = 0.5 log det(cov1)."""
    D, n, steps, L = 3, 4000, 40, 5
    e0, e1, mu1, cov1, log_ratio = _ais_problem(D)
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal((n, D))
    est, alpha, x, w = U.O.ais_estimate(e0, e1, steps, x0, step_size=0.3, leapfrogs=L, v0=rng.standard_normal((n, D)),
                                        v_refresh=rng.standard_normal((steps, n, D)), u=rng.random((steps, n)))
    assert abs(float(est) - log_ratio) < 0.1, (float(est), log_ratio)
    assert 0.5 < float(alpha) <= 1.0
    # the final particles target N(mu1, cov1) (weights aside: close already, the annealing is slow)
    assert np.abs(x.numpy().mean(0) - mu1).max() < 0.15


def _train_setup(n=64, seed=5):
    P = U.Problem(regime="stress", **U.CONFIGS["c1_scg2"])
    dyn = P.oracle(torch.float64)
    rng = np.random.default_rng(seed)

    def draw():
        return {"direction": torch.as_tensor(rng.integers(0, 2, n).astype(np.float64)),
                "v_f": torch.as_tensor(rng.standard_normal((n, P.D))), "v_b": torch.as_tensor(rng.standard_normal((n, P.D)))}
    x = torch.as_tensor(P.x0(n, rng)).double()
    z = torch.as_tensor(rng.standard_normal((n, P.D)))
    return P, dyn, x, z, draw


def test_training_losses_and_gradient_through_the_dynamics():
    """utils/losses.py and the notebook objective on the oracle: autograd through the unrolled leapfrog (second-order in
    the energy) agrees with central finite differences of the same loss."""
    P, dyn, x, z, draw = _train_setup()
    rx, rz = draw(), draw()
    params = U.O.trainable_parameters(dyn)
    loss = U.O.notebook_loss(x, z, dyn, rx, rz)
    assert torch.isfinite(loss)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    assert any(g is not None and float(g.abs().max()) > 0 for g in grads)
    # finite differences on a few entries of the largest-gradient tensor
    idx = max(range(len(params)), key=lambda i: 0.0 if grads[i] is None else float(grads[i].abs().max()))
    w, g = params[idx], grads[idx]
    flat = g.reshape(-1)
    for j in torch.topk(flat.abs(), 3).indices.tolist():
        h = 1e-6
        with torch.no_grad():
            w.reshape(-1)[j] += h
            lp = float(U.O.notebook_loss(x, z, dyn, rx, rz))
            w.reshape(-1)[j] -= 2 * h
            lm = float(U.O.notebook_loss(x, z, dyn, rx, rz))
            w.reshape(-1)[j] += h
        fd = (lp - lm) / (2 * h)
        assert abs(fd - float(flat[j])) <= 1e-5 * max(1.0, abs(fd)), (fd, float(flat[j]))
    # the library of losses on the same proposal
    with torch.no_grad():
        Lx, _, px, _ = U.O.propose(x, dyn, direction=rx["direction"], v_f=rx["v_f"], v_b=rx["v_b"])
        v = U.O.loss_vec(x, Lx, px)
    assert float(U.O.loss_std(x, Lx, px)) == pytest.approx(-float(v.mean()))
    assert float(U.O.loss_mixed(x, Lx, px, 0.1)) == pytest.approx(float((0.1 / v).mean() - (v / 0.1).mean()))
    assert float(U.O.loss_inverse(x, Lx, px)) < 0 and torch.isfinite(U.O.loss_logsumexp(x, Lx, px))


def test_a_few_adam_steps_lower_the_notebook_loss():
    """The notebook's training loop (SCGExperiment.ipynb:254-270) in miniature on the oracle: Adam on both nets with fixed
    randomness lowers the objective."""
    P, dyn, x, z, draw = _train_setup(n=128)
    rx, rz = draw(), draw()
    params = U.O.trainable_parameters(dyn)
    opt = torch.optim.Adam(params, lr=1e-3)
    first = None
    for it in range(12):
        opt.zero_grad()
        loss = U.O.notebook_loss(x, z, dyn, rx, rz)
        loss.backward()
        opt.step()
        first = float(loss.detach()) if first is None else first
    assert float(U.O.notebook_loss(x, z, dyn, rx, rz).detach()) < first


@pytest.mark.parametrize("name,temperature", [("c1_scg2", 1.0), ("c3_mog2", 1.0), ("c4_rw32", 1.0), ("funnel3", 1.0),
                                              ("c1_scg2", 1.7)])
def test_hand_written_reverse_pass_equals_autograd(name, temperature):
    """oracle/l2hmc_reverse.py (the reverse sweep a CUDA backward kernel would follow: per sub-update vector-Jacobian
    products, Hessian-vector products of the energy, no autograd) gives the gradient torch.autograd gets through the
    restated dynamics -- every parameter tensor of both nets and d/d(eps) -- for mixed directions."""
    import l2hmc_reverse as R
    kw = dict(U.CONFIGS[name])
    kw["T"] = min(kw["T"], 4)
    kw["H"] = min(kw["H"], 12)
    P = U.Problem(regime="stress", **kw)
    dyn = P.oracle(torch.float64)
    dyn.temperature = temperature
    n = 24
    rng = np.random.default_rng(11)

    def draw():
        return {"direction": torch.as_tensor(rng.integers(0, 2, n).astype(np.float64)),
                "v_f": torch.as_tensor(rng.standard_normal((n, P.D))), "v_b": torch.as_tensor(rng.standard_normal((n, P.D)))}
    x = torch.as_tensor(P.x0(n, rng)).double()
    z = torch.as_tensor(rng.standard_normal((n, P.D)))
    rx, rz = draw(), draw()

    loss_h, g_h = R.notebook_loss_and_grads(x, z, dyn, rx, rz)
    assert not any(t.requires_grad for t in dyn.xnet.values())

    params = U.O.trainable_parameters(dyn)
    dyn._eps = dyn._eps.clone().requires_grad_(True)
    loss_a = U.O.notebook_loss(x, z, dyn, rx, rz)
    grads = torch.autograd.grad(loss_a, params + [dyn._eps])
    assert float(loss_h) == pytest.approx(float(loss_a.detach()), rel=1e-12)
    keys = sorted(dyn.xnet)
    flat_h = [g_h["xnet"][k] for k in keys] + [g_h["vnet"][k] for k in keys] + [g_h["eps"]]
    worst = 0.0
    for gh, ga in zip(flat_h, grads):
        assert gh.shape == ga.shape
        worst = max(worst, float((gh - ga).abs().max()) / max(1e-12, float(ga.abs().max())))
    assert worst < 1e-9, worst
    assert float(grads[-1].abs()) > 0 and all(float(g.abs().max()) > 0 for g in grads)
    assert float(g_h["alpha"]) == pytest.approx(float((grads[-1] * dyn._eps).detach()), rel=1e-9)


@pytest.mark.parametrize("name", ["c1_scg2", "c3_mog2", "c4_rw32", "c4_rw32_hard", "funnel3"])
def test_closed_form_hessian_vector_products_equal_autograd(name):
    """l2hmc_reverse.energy_hvp (the forms the training kernels use) against autograd of the gradient expression,
    including funnel rows beyond the clip where the coupling to x_0 drops out (utils/distributions.py:161-180)."""
    import l2hmc_reverse as R
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    en = P.energy.to(torch.float64)
    rng = np.random.default_rng(2)
    x = torch.as_tensor(P.x0(12, rng)).double()
    if name == "funnel3":
        x[0, 0], x[1, 0] = 9.5, -8.5
    w = torch.as_tensor(rng.standard_normal(x.shape))
    xr = x.clone().requires_grad_(True)
    (ref,) = torch.autograd.grad((en.grad(xr) * w).sum(), xr)
    got = R.energy_hvp(en, x, w)
    assert float((got - ref).abs().max()) <= 1e-10 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("kind", ["mixed", "standard", "inverse", "logsumexp"])
def test_hand_written_reverse_pass_for_every_loss_of_the_library(kind):
    """get_loss(name) of utils/losses.py:26-59: the hand-written d loss / d v and sweep against autograd of the oracle's
    loss_mixed / loss_std / loss_inverse / loss_logsumexp through the dynamics."""
    import l2hmc_reverse as R
    P = U.Problem(regime="stress", kind="gaussian", D=2, H=8, T=3, eps=0.1)
    dyn = P.oracle(torch.float64)
    rng = np.random.default_rng(3)
    n = 20
    x = torch.as_tensor(P.x0(n, rng)).double()
    r = {"direction": torch.as_tensor(rng.integers(0, 2, n).astype(np.float64)),
         "v_f": torch.as_tensor(rng.standard_normal((n, P.D))), "v_b": torch.as_tensor(rng.standard_normal((n, P.D)))}
    with torch.no_grad():
        acc = R._Acc(dyn)
        loss_h = R.loss_and_grads(x, dyn, r, 0.1, acc, kind=kind)
    params = U.O.trainable_parameters(dyn)
    Lx, _, px, _ = U.O.propose(x, dyn, direction=r["direction"], v_f=r["v_f"], v_b=r["v_b"])
    fn = {"mixed": lambda: U.O.loss_mixed(x, Lx, px, 0.1), "standard": lambda: U.O.loss_std(x, Lx, px),
          "inverse": lambda: U.O.loss_inverse(x, Lx, px), "logsumexp": lambda: U.O.loss_logsumexp(x, Lx, px)}[kind]
    loss_a = fn()
    grads = torch.autograd.grad(loss_a, params)
    assert float(loss_h) == pytest.approx(float(loss_a.detach()), rel=1e-12)
    keys = sorted(dyn.xnet)
    for gh, ga in zip([acc.x[k] for k in keys] + [acc.v[k] for k in keys], grads):
        assert float((gh - ga).abs().max()) <= 1e-9 * max(1e-12, float(ga.abs().max()))
