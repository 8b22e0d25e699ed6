"""Read / write the golden fixtures under tests/golden/ (fp64-oracle transitions on fixed inputs).

A fixture is self-contained: target parameters, both nets, masks, the injected randomness and the
oracle's outputs -- nothing is regenerated from a seed at load time.
"""
from __future__ import annotations

import json

import numpy as np

import util as U


def save(path, name, kw, n, seed, regime, log_jac=False):
    import torch
    P = U.Problem(regime=regime, **kw)
    d = P.draws(n, seed)
    ref = U.run_oracle_propose(P, d, torch.float64, log_jac)
    arrays = {"meta": np.frombuffer(json.dumps({"name": name, "kw": kw, "n": n, "regime": regime,
                                                 "log_jac": log_jac}).encode(), dtype=np.uint8),
              "mask": P.mask}
    if not P.hmc:
        for k, v in P.xnet.items():
            arrays["xnet_" + k] = v
        for k, v in P.vnet.items():
            arrays["vnet_" + k] = v
    for k, v in d.items():
        arrays["in_" + k] = v
    for k, v in ref.items():
        arrays["out_" + k] = v  # fp64
    np.savez_compressed(path, **arrays)


def save_vae(path, name, kw, n, seed):
    """Fixture for the decoder-energy / aux-conditioned problem (BASELINE config 5 in miniature)."""
    import torch
    P = U.VaeProblem(**kw)
    d = P.draws(n, seed)
    ref = U.run_oracle_propose(P, d, torch.float64)
    arrays = {"meta": np.frombuffer(json.dumps({"name": name, "kw": kw, "n": n, "vae": True, "log_jac": False}).encode(), dtype=np.uint8),
              "mask": P.mask}
    for k, v in P.xnet.items():
        arrays["xnet_" + k] = v
    for k, v in P.vnet.items():
        arrays["vnet_" + k] = v
    for i, (W, b) in enumerate(zip(P.dec_W, P.dec_b)):
        arrays["decW_%d" % i], arrays["decb_%d" % i] = W, b
    if P.use_encoder:
        for i, (W, b) in enumerate(zip(P.enc_W, P.enc_b)):
            arrays["encW_%d" % i], arrays["encb_%d" % i] = W, b
    for k, v in d.items():
        arrays["in_" + k] = v
    for k, v in ref.items():
        arrays["out_" + k] = v  # fp64
    np.savez_compressed(path, **arrays)


def _load_vae(z, meta):
    P = U.VaeProblem(**meta["kw"])
    P.mask = z["mask"]
    P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
    P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
    P.dec_W = [z["decW_%d" % i] for i in range(len(P.dec_W))]
    P.dec_b = [z["decb_%d" % i] for i in range(len(P.dec_b))]
    if P.use_encoder:
        P.enc_W = [z["encW_%d" % i] for i in range(len(P.enc_W))]
        P.enc_b = [z["encb_%d" % i] for i in range(len(P.enc_b))]
    d = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    P.meta = meta
    return P, d, ref


def load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    if meta.get("vae"):
        return _load_vae(z, meta)
    P = U.Problem(regime=meta["regime"], **meta["kw"])
    P.mask = z["mask"]
    if not P.hmc:
        P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
        P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
    d = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    P.meta = meta
    return P, d, ref
