"""CPU check of the training-path kernels' SOURCE (l2hmc_b200/csrc/train.cuh + train_host.cuh).

tests/emu/train_emu.cpp compiles those two files with g++ and runs every CUDA thread on a host thread, so index
arithmetic, operand strides, the GEMM tiling, the order of the two sweeps and every vector-Jacobian product are checked
here against oracle/l2hmc_reverse.py (which equals torch.autograd through the restated dynamics, see test_oracle.py)
without a GPU.  The GPU run of the same source is tests/test_gpu_training.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import util as U
import l2hmc_reverse as R

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "train_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libtrain_emu.so")
F = C.POINTER(C.c_float)
NAMES = ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "Ws", "bs", "Wt", "bt", "Wq", "bq", "scale_s", "scale_q")
ORACLE_KEY = dict(zip(NAMES, ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "Ws", "bs", "Wt", "bt", "Wq", "bq", "ls", "lq")))


class NetParams(C.Structure):   # l2hmc_net_params (include/l2hmc.h)
    _fields_ = [(k, F) for k in NAMES]


# l2hmc_net_grads / l2hmc_loss_grad_args: the product's own ctypes mirrors, so that their layout is checked against the
# header as g++ compiles it
from l2hmc_b200._lib import NetGrads, LossGradArgs  # noqa: E402


def vptr(a):
    return a.ctypes.data


def build_emu():
    """Compile (if stale) and load the host build of the training kernels; also used by tests/test_distributed_cpu.py."""
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(HERE, "..", "l2hmc_b200", "csrc", f) for f in ("train.cuh", "train_host.cuh")] + \
           [os.path.join(HERE, "..", "include", "l2hmc.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-DL2HMC_TRAIN_EMU", "-DL2HMC_TR_KCHUNK=16", "-DL2HMC_TR_SLAB=16", "-x", "c++",
                        "-Wno-unknown-pragmas", SRC, "-o", OUT], check=True)
    lib = C.CDLL(OUT)
    lib.emu_loss_grad.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def fptr(a):
    return a.ctypes.data_as(F)


def pack(net, cls=NetParams):
    keep = {k: np.ascontiguousarray(np.asarray(net[ORACLE_KEY[k]], dtype=np.float32)) for k in NAMES}
    return cls(**{k: fptr(v) for k, v in keep.items()}), keep


def energy_args(P):
    """(kind, ncomp, mu, S, logc, s0, s1) of a util.Problem for the emulated context."""
    ncomp, logc, s0, s1 = 1, np.zeros(1, np.float32), 0.0, 0.0
    mu, S = np.zeros(P.D, np.float32), np.zeros((P.D, P.D), np.float32)
    if P.kind == "gaussian":
        kind, mu = 0, np.ascontiguousarray(P.energy.mu.numpy(), np.float32)
        S = np.ascontiguousarray(P.energy.S.numpy(), np.float32)
    elif P.kind == "gmm":
        kind, ncomp = 1, len(P.energy.mus)
        mu = np.ascontiguousarray(np.stack([m.numpy() for m in P.energy.mus]), np.float32)
        S = np.ascontiguousarray(np.stack([m.numpy() for m in P.energy.Ss]), np.float32)
        logc = np.log(np.array([float(c) for c in P.energy.cs], np.float64)).astype(np.float32)
    elif P.kind == "roughwell":
        kind = 2
        e, den = P.energy._scale(torch.zeros(1))
        s0, s1 = float(e), float(den)
    else:
        kind, s0, s1 = 3, P.energy.sigma, P.energy.clip
    return kind, ncomp, mu, S, logc, s0, s1


def run_emu(lib, P, x, v, direction, scale, inv_count, temperature=1.0, loss_kind=0):
    n = x.shape[0]
    xp, keep_x = pack(P.xnet)
    vp, keep_v = pack(P.vnet)
    gx = {k: np.zeros_like(a) for k, a in keep_x.items()}
    gv = {k: np.zeros_like(a) for k, a in keep_v.items()}
    loss = np.zeros(1, np.float32)
    d_eps = np.zeros(1, np.float32)
    Lx = np.zeros((n, P.D), np.float32)
    px = np.zeros(n, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    v = np.ascontiguousarray(v, np.float32)
    d8 = np.ascontiguousarray(direction, np.uint8)
    a = LossGradArgs(n=n, x=vptr(x), v=vptr(v), dir=vptr(d8), scale=scale, inv_count=inv_count,
                     loss=vptr(loss), d_eps=vptr(d_eps), grad_xnet=NetGrads(**{k: vptr(g) for k, g in gx.items()}),
                     grad_vnet=NetGrads(**{k: vptr(g) for k, g in gv.items()}), x_out=vptr(Lx), px_out=vptr(px), stream=None,
                     loss_kind=loss_kind)
    mask = np.ascontiguousarray(P.mask, np.float32)
    kind, ncomp, mu, S, logc, s0, s1 = energy_args(P)
    err = C.create_string_buffer(512)
    rc = lib.emu_loss_grad(C.c_int(P.D), C.c_int(P.H), C.c_int(P.T), C.c_float(P.eps), C.c_float(temperature), C.c_int(kind),
                           C.c_int(ncomp), fptr(mu), fptr(S), fptr(logc), C.c_float(s0), C.c_float(s1), fptr(mask), C.byref(xp),
                           C.byref(vp), C.byref(a), err, C.c_int(512))
    assert rc == 0, err.value
    return float(loss[0]), float(d_eps[0]), gx, gv, Lx, px


@pytest.mark.parametrize("kind,D,H,T,n,temperature", [
    ("gaussian", 3, 5, 2, 9, 1.0),       # nothing a multiple of anything
    ("gaussian", 2, 10, 3, 40, 1.7),     # the notebook's shape, T_emp != 1, more than one warp-block of chains
    ("roughwell", 5, 7, 2, 12, 1.0),
    ("gmm", 2, 10, 3, 24, 1.0),          # config 3's target: the mixture's Hessian-vector product
    ("gmm", 3, 6, 2, 10, 1.3),
    ("funnel", 3, 10, 3, 20, 1.0),
    ("gaussian", 9, 70, 1, 70, 1.0),     # width and chain count beyond one 64-wide GEMM tile
    ("gaussian", 2, 1, 1, 1, 1.0),       # one chain, one hidden unit, one leapfrog step
    ("gaussian", 50, 100, 2, 130, 1.0),  # BASELINE config 2's shape (fewer steps and chains)
    ("roughwell", 32, 100, 2, 66, 1.0),  # config 4's shape
])
def test_training_kernels_under_emulation_match_the_hand_written_reverse_pass(emu, kind, D, H, T, n, temperature):
    kw = dict(kind=kind, D=D, H=H, T=T, eps=0.1)
    if kind == "roughwell":
        kw["easy"] = True
    P = U.Problem(regime="stress", **kw)
    rng = np.random.default_rng(4)
    x = P.x0(n, rng)
    r = {"direction": torch.as_tensor(rng.integers(0, 2, n).astype(np.float64)),
         "v_f": torch.as_tensor(rng.standard_normal((n, D)).astype(np.float32)).double(),
         "v_b": torch.as_tensor(rng.standard_normal((n, D)).astype(np.float32)).double()}
    d = r["direction"].numpy().astype(np.uint8)
    v = np.where(d[:, None] == 1, r["v_f"].numpy(), r["v_b"].numpy()).astype(np.float32)
    scale = 0.1

    dyn = P.oracle(torch.float64, temperature=temperature)
    with torch.no_grad():
        acc = R._Acc(dyn)
        loss_o = R.loss_and_grads(torch.as_tensor(x).double(), dyn, r, scale, acc)
        Lx_o, _, px_o = U.O.propose_selected(torch.as_tensor(x).double(), dyn, direction=r["direction"],
                                             v=torch.as_tensor(v).double())
    loss, d_eps, gx, gv, Lx, px = run_emu(emu, P, x, v, d, scale, 1.0 / n, temperature)

    assert np.abs(Lx - Lx_o.numpy()).max() <= 2e-5 * max(1.0, float(Lx_o.abs().max()))
    assert np.abs(px - px_o.numpy()).max() <= 1e-4   # H is O(100) at 50 dimensions: fp32 cancellation in H0 - H1 + log|J|
    assert loss == pytest.approx(float(loss_o), rel=2e-4)
    assert d_eps == pytest.approx(float(acc.eps), rel=2e-3, abs=1e-3 * abs(float(loss_o)))
    worst = 0.0
    for got, ref in ((gx, acc.x), (gv, acc.v)):
        for k in NAMES:
            a, b = got[k].astype(np.float64), ref[ORACLE_KEY[k]].numpy().reshape(got[k].shape)
            assert np.isfinite(a).all()
            worst = max(worst, float(np.abs(a - b).max() / max(1e-12, np.abs(b).max())))
    # fp32 sweep against the fp64 statement: the fp32 noise floor of this gradient is about 1e-5
    assert worst < 2e-4, worst   # measured 1e-6 .. 1.3e-5


def test_emulated_gradients_accumulate_over_two_batches(emu):
    """Outputs are += : the x batch and the z batch of the notebook objective add up (SCGExperiment.ipynb:159-181)."""
    P = U.Problem(regime="stress", kind="gaussian", D=2, H=6, T=2, eps=0.1)
    rng = np.random.default_rng(9)
    n = 8
    x, z = P.x0(n, rng), rng.standard_normal((n, 2)).astype(np.float32)
    d1, d2 = rng.integers(0, 2, n).astype(np.uint8), rng.integers(0, 2, n).astype(np.uint8)
    v1, v2 = rng.standard_normal((n, 2)).astype(np.float32), rng.standard_normal((n, 2)).astype(np.float32)
    la, ea, gxa, gva, _, _ = run_emu(emu, P, x, v1, d1, 0.1, 1.0 / n)
    lb, eb, gxb, gvb, _, _ = run_emu(emu, P, z, v2, d2, 0.1, 1.0 / n)
    lab, eab, gxab, gvab, _, _ = run_emu(emu, P, np.concatenate([x, z]), np.concatenate([v1, v2]), np.concatenate([d1, d2]),
                                         0.1, 1.0 / n)
    assert lab == pytest.approx(la + lb, rel=1e-5)
    assert eab == pytest.approx(ea + eb, rel=1e-4, abs=1e-4)
    for k in NAMES:
        assert np.allclose(gxab[k], gxa[k] + gxb[k], rtol=1e-4, atol=1e-5 * max(1e-6, np.abs(gxab[k]).max()))
        assert np.allclose(gvab[k], gva[k] + gvb[k], rtol=1e-4, atol=1e-5 * max(1e-6, np.abs(gvab[k]).max()))


class _EmuLib:
    """Stands in for libl2hmc.so under a Dynamics whose tensors live on the CPU: l2hmc_loss_grad goes to the emulated
    kernels, so the marshalling in l2hmc_b200/training.py (struct filling, accumulation, alpha, Adam write-back) runs here."""

    def __init__(self, lib, P):
        self.lib, self.P = lib, P
        self.calls = 0

    def l2hmc_loss_grad(self, ctx, a_ref):
        P = self.P
        a = a_ref._obj
        xp, self._kx = pack(self.dyn._net_params[0])
        vp, self._kv = pack(self.dyn._net_params[1])
        mask = np.ascontiguousarray(self.dyn.mask, np.float32)
        mu = np.ascontiguousarray(P.energy.mu.numpy(), np.float32)
        S = np.ascontiguousarray(P.energy.S.numpy(), np.float32)
        err = C.create_string_buffer(512)
        self.calls += 1
        return self.lib.emu_loss_grad(C.c_int(P.D), C.c_int(P.H), C.c_int(P.T), C.c_float(self.dyn.eps), C.c_float(1.0), C.c_int(0),
                                      C.c_int(1), fptr(mu), fptr(S), fptr(np.zeros(1, np.float32)), C.c_float(0), C.c_float(0),
                                      fptr(mask), C.byref(xp), C.byref(vp), C.byref(a), err, C.c_int(512))

    def l2hmc_last_error(self, ctx):
        return b"emulated"

    def l2hmc_set_eps(self, ctx, eps):   # the emulated call reads dyn.eps directly
        return 0


def test_training_module_marshalling_and_loop_over_the_emulated_kernels(emu, monkeypatch):
    from l2hmc_b200 import training
    P = U.Problem(regime="stress", kind="gaussian", D=2, H=6, T=2, eps=0.1)
    dyn = P.product()                      # no GPU here: the object has no library context
    shim = _EmuLib(emu, P)
    shim.dyn = dyn
    dyn._lib, dyn._ctx = shim, C.c_void_p(1)
    monkeypatch.setattr(type(dyn), "_prep", lambda self, t, name, cols=None: t.detach().to(torch.float32).contiguous())
    monkeypatch.setattr(type(dyn), "_stream", lambda self: None)
    monkeypatch.setattr(type(dyn), "_sync_temperature", lambda self: None)
    monkeypatch.setattr(type(dyn), "_ensure_ctx", lambda self: None)
    monkeypatch.setattr(type(dyn), "_push_nets", lambda self: None)
    monkeypatch.setattr(type(dyn), "_chk", lambda self, rc: (_ for _ in ()).throw(RuntimeError(rc)) if rc else None)
    rng = np.random.default_rng(2)
    n = 16

    def draw(dynamics, n_, device, want_u=False, chain_offset=0):
        return (torch.as_tensor(rng.integers(0, 2, n_).astype(np.uint8)), torch.as_tensor(rng.standard_normal((n_, P.D)).astype(np.float32)),
                None)
    monkeypatch.setattr(training, "_draw", draw)
    monkeypatch.setattr(training, "tf_accept", lambda x, Lx, px, u=None, seed=0, counter=0, chain_offset=0: torch.where((px >= 0.5)[:, None], Lx, x))

    x = torch.as_tensor(P.x0(n, rng))
    d = torch.as_tensor(rng.integers(0, 2, n).astype(np.uint8))
    v = torch.as_tensor(rng.standard_normal((n, P.D)).astype(np.float32))
    loss, grads, Lx, px = training.loss_and_grads(dyn, x, rng={"direction": d, "v": v})
    odyn = P.oracle(torch.float64)
    r = {"direction": d.double(), "v_f": v.double(), "v_b": v.double()}
    with torch.no_grad():
        acc = R._Acc(odyn)
        loss_o = R.loss_and_grads(x.double(), odyn, r, 0.1, acc)
    assert float(loss[0]) == pytest.approx(float(loss_o), rel=2e-4)
    for key, ref in (("XNet", acc.x), ("VNet", acc.v)):
        for k in training.NAMES:
            a, b = grads[key][k].double().numpy(), ref[k].numpy().reshape(grads[key][k].shape)
            assert np.abs(a - b).max() <= 2e-4 * max(1e-12, np.abs(b).max()), (key, k)
    assert float(grads["alpha"][0]) == pytest.approx(float(acc.eps) * 0.1, rel=2e-3)
    assert px.shape == (n,) and Lx.shape == (n, P.D) and float(px.min()) >= 0.0 and float(px.max()) <= 1.0

    # the loop: fixed-randomness loss before / after a few Adam steps, parameters and eps move, weights reach the layers
    z = torch.as_tensor(rng.standard_normal((n, P.D)).astype(np.float32))
    fixed = dict(rng_x={"direction": d, "v": v}, rng_z={"direction": d, "v": v})
    first, _, _, _ = training.notebook_loss_and_grads(dyn, x, z, **fixed)
    opt = training.Adam(dyn)
    W4_before, eps_before = np.array(dyn._net_params[0]["W4"]), dyn.eps
    samples = x
    for _ in range(6):
        out = training.train_step(dyn, opt, samples)
        samples = out["samples"]
        assert np.isfinite(out["loss"])
    last, _, _, _ = training.notebook_loss_and_grads(dyn, x, z, **fixed)
    assert float(last[0]) < float(first[0])
    assert opt.global_step == 6 and dyn.eps != eps_before
    assert not np.array_equal(W4_before, np.asarray(dyn._net_params[0]["W4"]))
    assert shim.calls == 1 + 2 + 6 * 2 + 2


@pytest.mark.parametrize("name,temperature", [("c1_scg2", 1.0), ("c3_mog2", 1.3), ("c4_rw32", 1.0), ("c4_rw32_hard", 1.0),
                                              ("funnel3", 2.0)])
def test_hessian_vector_product_kernel_under_emulation(emu, name, temperature):
    """k_hvp alone against the oracle's closed forms, funnel rows beyond the clip included."""
    P = U.Problem(regime="stress", **U.CONFIGS[name])
    rng = np.random.default_rng(6)
    n = 150                                    # more than one block of 128 threads
    x = P.x0(n, rng)
    if name == "funnel3":
        x[0, 0], x[1, 0] = 9.5, -8.5
    w = rng.standard_normal(x.shape).astype(np.float32)
    base = rng.standard_normal(x.shape).astype(np.float32)
    out = base.copy()
    kind, ncomp, mu, S, logc, s0, s1 = energy_args(P)
    emu.emu_hvp.restype = C.c_int
    rc = emu.emu_hvp(C.c_int(P.D), C.c_float(temperature), C.c_int(kind), C.c_int(ncomp), fptr(mu), fptr(S), fptr(logc),
                     C.c_float(s0), C.c_float(s1), C.c_longlong(n), fptr(np.ascontiguousarray(x)), fptr(w), fptr(out))
    assert rc == 0
    ref = R.energy_hvp(P.energy.to(torch.float64), torch.as_tensor(x).double(), torch.as_tensor(w).double()).numpy() / temperature
    assert np.abs((out - base) - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("kind", ["standard", "inverse", "logsumexp"])
def test_library_losses_under_emulation(emu, kind):
    """get_loss(name) of utils/losses.py:26-59 through the emulated kernels (the batch statistics of 'inverse' and
    'logsumexp' come from a one-block reduction), against the hand-written reverse pass."""
    from l2hmc_b200.training import LOSSES
    P = U.Problem(regime="stress", kind="gaussian", D=3, H=6, T=2, eps=0.1)
    rng = np.random.default_rng(12)
    n = 300                                   # more rows than the reduction block has threads
    x = P.x0(n, rng)
    d = rng.integers(0, 2, n).astype(np.uint8)
    v = rng.standard_normal((n, P.D)).astype(np.float32)
    r = {"direction": torch.as_tensor(d.astype(np.float64)), "v_f": torch.as_tensor(v).double(), "v_b": torch.as_tensor(v).double()}
    dyn = P.oracle(torch.float64)
    with torch.no_grad():
        acc = R._Acc(dyn)
        loss_o = R.loss_and_grads(torch.as_tensor(x).double(), dyn, r, 0.1, acc, kind=kind)
    loss, d_eps, gx, gv, _, _ = run_emu(emu, P, x, v, d, 0.1, 1.0 / n, loss_kind=LOSSES[kind])
    assert loss == pytest.approx(float(loss_o), rel=2e-4)
    assert d_eps == pytest.approx(float(acc.eps), rel=2e-3, abs=1e-3 * abs(float(loss_o)))
    for got, ref in ((gx, acc.x), (gv, acc.v)):
        for k in NAMES:
            a, b = got[k].astype(np.float64), ref[ORACLE_KEY[k]].numpy().reshape(got[k].shape)
            assert np.abs(a - b).max() <= 2e-4 * max(1e-12, np.abs(b).max()), (kind, k)


@pytest.mark.parametrize("all_forward", [True, False])
def test_single_direction_batches_under_emulation(emu, all_forward):
    """Dynamics.forward only / Dynamics.backward only (utils/dynamics.py:246-300): every chain in one direction."""
    P = U.Problem(regime="stress", kind="gaussian", D=4, H=6, T=3, eps=0.1)
    rng = np.random.default_rng(21)
    n = 11
    x = P.x0(n, rng)
    d = np.full(n, 1 if all_forward else 0, np.uint8)
    v = rng.standard_normal((n, P.D)).astype(np.float32)
    r = {"direction": torch.as_tensor(d.astype(np.float64)), "v_f": torch.as_tensor(v).double(), "v_b": torch.as_tensor(v).double()}
    dyn = P.oracle(torch.float64)
    with torch.no_grad():
        acc = R._Acc(dyn)
        loss_o = R.loss_and_grads(torch.as_tensor(x).double(), dyn, r, 0.1, acc)
    loss, d_eps, gx, gv, _, _ = run_emu(emu, P, x, v, d, 0.1, 1.0 / n)
    assert loss == pytest.approx(float(loss_o), rel=2e-4)
    for got, ref in ((gx, acc.x), (gv, acc.v)):
        for k in NAMES:
            a, b = got[k].astype(np.float64), ref[ORACLE_KEY[k]].numpy().reshape(got[k].shape)
            assert np.abs(a - b).max() <= 2e-4 * max(1e-12, np.abs(b).max()), k
