"""Generate the reference-pinned golden fixtures: python tests/golden/make_ref_golden.py   (CPU, ~1 min)

Every array written here is an output of the UNMODIFIED reference sources under /root/reference
(utils/dynamics.py, sampler.py, layers.py, distributions.py, losses.py, ais.py, func_utils.py; the `network` cell of
SCGExperiment.ipynb; for config 5 the decoder / energy / sampler-net text of mnist_vae.py) executed on the eager
TensorFlow stand-in oracle/tf_shim (see oracle/ref_loader.py, oracle/ref_runner.py) with injected parameters and
randomness.  The reference is run twice per case: in float64 (`out_*`, the ground truth the oracle and the CUDA kernels
are compared with) and in float32 (`out32_*`, which measures the reference's own rounding noise on the same inputs).
The GPU box has no /root/reference; it only reads the committed .npz files.

  tests/golden/ref_<case>.npz      propose() fixtures, same schema as make_golden.py's (tests/golden_io.py loads both)
  tests/golden/ref/<what>.npz      Dynamics methods, chain_operator, training objective + gradients, losses,
                                   diagnostics, AIS, the VAE posterior target
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.dirname(HERE), ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_runner as R  # noqa: E402
import util as U  # noqa: E402

PROPOSE = [
    # (file, config, n, seed, regime, log_jac)
    ("ref_c1_scg2_n200_init", "c1_scg2", 200, 21, "init", False),       # BASELINE config 1, notebook cell verbatim
    ("ref_c1_scg2_n200_stress", "c1_scg2", 200, 22, "stress", False),
    ("ref_c2_scg50_n128_stress", "c2_scg50", 128, 23, "stress", False),  # BASELINE config 2, reduced chain count
    ("ref_c2_scg50_n128_logjac", "c2_scg50", 128, 24, "stress", True),
    ("ref_c3_mog2_n256_stress", "c3_mog2", 256, 25, "stress", False),    # BASELINE config 3
    ("ref_c4_rw32_n128_stress", "c4_rw32", 128, 26, "stress", False),    # BASELINE config 4
    ("ref_c4_rw32_hard_n128_stress", "c4_rw32_hard", 128, 27, "stress", False),
    ("ref_funnel3_n128_stress", "funnel3", 128, 28, "stress", False),
]
HMC = dict(kind="gaussian", D=2, T=10, eps=0.15, hmc=True)  # utils/notebook_utils.py:25-28 as the notebook calls it


def _meta(**kw):
    kw["source"] = "unmodified /root/reference on oracle/tf_shim (tests/golden/make_ref_golden.py)"
    return np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)


def problem_arrays(P):
    arrays = {"mask": P.mask}
    if not P.hmc:
        for k, v in P.xnet.items():
            arrays["xnet_" + k] = v
        for k, v in P.vnet.items():
            arrays["vnet_" + k] = v
    return arrays


def save_propose(fname, kw, n, seed, regime, log_jac):
    P = U.Problem(regime=regime, **kw)
    d = P.draws(n, seed)
    arrays = problem_arrays(P)
    arrays["meta"] = _meta(name=fname, kw=kw, n=n, regime=regime, log_jac=log_jac)
    for k, v in d.items():
        arrays["in_" + k] = v
    for k, v in R.run_propose(P, d, "float64", log_jac).items():
        arrays["out_" + k] = v
    for k, v in R.run_propose(P, d, "float32", log_jac).items():
        arrays["out32_" + k] = v
    if not log_jac:
        for k, v in R.run_methods(P, d, "float64").items():
            arrays["m_" + k] = v
        for k, v in R.run_methods(P, d, "float32").items():
            arrays["m32_" + k] = v
    path = os.path.join(HERE, fname + ".npz")
    np.savez_compressed(path, **arrays)
    return path


def save_chain_operator(fname, cfg, n, nb_steps, seed):
    P = U.Problem(regime="stress", **U.CONFIGS[cfg])
    rng = np.random.default_rng(seed)
    x = P.x0(n, rng)
    f32 = lambda a: a.astype(np.float32)  # noqa: E731
    init_v = f32(rng.standard_normal((n, P.D)))
    dirs = rng.integers(0, 2, (nb_steps, n)).astype(np.uint8)
    v_fs = f32(rng.standard_normal((nb_steps, n, P.D)))
    v_bs = f32(rng.standard_normal((nb_steps, n, P.D)))
    u = f32(rng.random(n))
    arrays = problem_arrays(P)
    arrays["meta"] = _meta(name=fname, kw=U.CONFIGS[cfg], n=n, regime="stress", nb_steps=nb_steps)
    arrays.update(in_x=x, in_init_v=init_v, in_dirs=dirs, in_v_fs=v_fs, in_v_bs=v_bs, in_u=u)
    for k, v in R.run_chain_operator(P, x, nb_steps, init_v, dirs, v_fs, v_bs, u, "float64").items():
        arrays["out_" + k] = v
    path = os.path.join(HERE, "ref", fname + ".npz")
    np.savez_compressed(path, **arrays)
    return path


def save_notebook_loss(fname, cfg, n, seed):
    P = U.Problem(regime="stress", **U.CONFIGS[cfg])
    rng = np.random.default_rng(seed)
    f32 = lambda a: a.astype(np.float32)  # noqa: E731
    x = P.x0(n, rng)
    z = f32(rng.standard_normal((n, P.D)))
    draw = lambda: {"dir": rng.integers(0, 2, n).astype(np.uint8), "v_f": f32(rng.standard_normal((n, P.D))),  # noqa: E731
                    "v_b": f32(rng.standard_normal((n, P.D))), "u": f32(rng.random(n))}
    rx, rz = draw(), draw()
    arrays = problem_arrays(P)
    arrays["meta"] = _meta(name=fname, kw=U.CONFIGS[cfg], n=n, regime="stress", scale=0.1)
    arrays.update(in_x=x, in_z=z)
    for k, v in rx.items():
        arrays["in_rx_" + k] = v
    for k, v in rz.items():
        arrays["in_rz_" + k] = v
    for k, v in R.run_notebook_loss(P, x, z, rx, rz, 0.1, "float64").items():
        arrays["out_" + k.replace("/", "__")] = v
    path = os.path.join(HERE, "ref", fname + ".npz")
    np.savez_compressed(path, **arrays)
    return path


def save_losses_and_diagnostics():
    import ref_loader
    ref = ref_loader.load()
    rng = np.random.default_rng(31)
    x = rng.standard_normal((64, 5)).astype(np.float32)
    X = (x + 0.3 * rng.standard_normal((64, 5))).astype(np.float32)
    p = rng.random(64).astype(np.float32)
    arrays = {"meta": _meta(name="losses_diagnostics"), "in_x": x, "in_X": X, "in_p": p}
    for k, v in R.run_losses(x, X, p, "float64").items():
        arrays["loss_" + k] = v
    # utils/func_utils.py:45-54,114-120 are plain numpy: run them as they are
    trace = np.cumsum(rng.standard_normal((60, 16, 2)), axis=0).astype(np.float32) * 0.1
    scale = np.sqrt(2.0)
    spec = ref.func_utils.acl_spectrum(trace, scale=scale)
    arrays.update(in_trace=trace, in_scale=np.float64(scale), acl_spectrum=spec, ess=np.float64(ref.func_utils.ESS(spec)),
                  autocov_3=np.float64(ref.func_utils.autocovariance(trace, tau=3)))
    path = os.path.join(HERE, "ref", "losses_diagnostics.npz")
    np.savez_compressed(path, **arrays)
    return path


def save_ais():
    import ref_loader
    import contextlib
    import io
    ref = ref_loader.load()
    out = []
    for name in ("ais_gauss3", "ais_roughwell4"):
        rng = np.random.default_rng(32 if name == "ais_gauss3" else 33)
        if name == "ais_gauss3":       # Gaussian -> Gaussian
            D, n, steps, lf, step = 3, 64, 12, 5, 0.3
            A = rng.standard_normal((D, D))
            cov1 = A @ A.T / D + 0.5 * np.eye(D)
            mu1 = rng.standard_normal(D)
            with contextlib.redirect_stdout(io.StringIO()):
                e0 = ref.distributions.Gaussian(np.zeros(D), np.eye(D))
                e1 = ref.distributions.Gaussian(mu1, cov1)
            extra = {"mu1": mu1, "cov1": cov1}
            meta = dict(D=D, n=n, anneal_steps=steps, leapfrogs=lf, step_size=step)
        else:                          # Gaussian -> rough well: a pair whose mixture is not a Gaussian (utils/ais.py:44-45)
            D, n, steps, lf, step = 4, 96, 10, 5, 0.2
            rw_eps = 0.3
            with contextlib.redirect_stdout(io.StringIO()):
                e0 = ref.distributions.Gaussian(np.zeros(D), 1.5 * np.eye(D))
                e1 = ref.distributions.RoughWell(D, rw_eps, easy=True)
            extra = {"cov0": 1.5 * np.eye(D)}
            meta = dict(D=D, n=n, anneal_steps=steps, leapfrogs=lf, step_size=step, rw_eps=rw_eps, easy=True)
        x0 = rng.standard_normal((n, D)).astype(np.float32)
        v0 = rng.standard_normal((n, D)).astype(np.float32)
        v_refresh = rng.standard_normal((steps, n, D)).astype(np.float32)
        u = rng.random((steps, n)).astype(np.float32)
        arrays = {"meta": _meta(name=name, **meta), "in_x": x0, "in_v0": v0, "in_v_refresh": v_refresh, "in_u": u}
        arrays.update(extra)
        for k, v in R.run_ais(e0, e1, steps, x0, v0, v_refresh, u, step, lf, "float64").items():
            arrays["out_" + k] = v
        path = os.path.join(HERE, "ref", name + ".npz")
        np.savez_compressed(path, **arrays)
        out.append(path)
    return out


weight_checksum = U.vae_weight_checksum


def save_vae(fname, cfg, n, seed, store_weights):
    P = U.VaeProblem(**U.VAE_CONFIGS[cfg])
    d = P.draws(n, seed)
    arrays = {"meta": _meta(name=fname, kw=U.VAE_CONFIGS[cfg], n=n, vae=True, log_jac=False, cfg=cfg,
                            weights_stored=store_weights, weight_checksum=weight_checksum(P)),
              "mask": P.mask}
    if store_weights:
        arrays.update(problem_arrays(P))
        for i, (W, b) in enumerate(zip(P.dec_W, P.dec_b)):
            arrays["decW_%d" % i], arrays["decb_%d" % i] = W, b
        if P.use_encoder:
            for i, (W, b) in enumerate(zip(P.enc_W, P.enc_b)):
                arrays["encW_%d" % i], arrays["encb_%d" % i] = W, b
    for k, v in d.items():
        arrays["in_" + k] = v
    for k, v in R.run_vae_propose(P, d, "float64").items():
        arrays["out_" + k] = v
    for k, v in R.run_vae_propose(P, d, "float32").items():
        arrays["out32_" + k] = v
    path = os.path.join(HERE, "ref", fname + ".npz")
    np.savez_compressed(path, **arrays)
    return path


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "ref"), exist_ok=True)
    out = []
    for fname, cfg, n, seed, regime, lj in PROPOSE:
        out.append(save_propose(fname, U.CONFIGS[cfg], n, seed, regime, lj))
    out.append(save_propose("ref_hmc_scg2_n200", HMC, 200, 29, "init", False))
    out.append(save_chain_operator("chain_operator_c1_n64", "c1_scg2", 64, 3, 41))
    out.append(save_chain_operator("chain_operator_c2_n32", "c2_scg50", 32, 2, 42))
    out.append(save_notebook_loss("notebook_loss_c1_n200", "c1_scg2", 200, 43))
    out.append(save_notebook_loss("notebook_loss_c3_n64", "c3_mog2", 64, 44))
    out.append(save_losses_and_diagnostics())
    out += save_ais()
    out.append(save_vae("c5_vae_mini_n96", "c5_vae_mini", 96, 45, True))
    out.append(save_vae("c5_vae_full_n32", "c5_vae_full", 32, 46, False))  # mnist_vae.py's own text, its layer sizes
    for p in out:
        print("%9d  %s" % (os.path.getsize(p), os.path.relpath(p, ROOT)))
