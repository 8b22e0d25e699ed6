"""Randomised sweep of the training-path kernels under host emulation (tests/emu) against oracle/l2hmc_reverse.py:
random shapes, energies, temperatures and losses.  Usage (repo root, no GPU needed): python tools/fuzz_train_emu.py [seed] [trials]
Round 1: seeds 1 and 2, 120 trials, 0 failures, worst per-tensor relative gradient error 9e-5."""
import sys, os
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle'); sys.path.insert(0,'.')
import numpy as np, torch, util as U, l2hmc_reverse as R, test_train_emu as E
from l2hmc_b200.training import LOSSES
lib = E.build_emu()
rs = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    kind = rs.choice(["gaussian", "gmm", "roughwell", "funnel"])
    D = int(rs.integers(1 if kind == "roughwell" else 2, 10)); H = int(rs.integers(1, 24)); T = int(rs.integers(1, 5)); n = int(rs.integers(1, 90))
    temp = float(rs.choice([1.0, 0.7, 2.3])); loss = str(rs.choice(list(LOSSES)))
    kw = dict(kind=kind, D=D, H=H, T=T, eps=float(rs.choice([0.05, 0.1, 0.2])), seed=int(rs.integers(0, 1000)))
    if kind == "roughwell": kw["easy"] = bool(rs.integers(0, 2)); 
    if kind == "roughwell" and not kw["easy"]: kw["eps"] = 0.01
    P = U.Problem(regime="stress", **kw)
    rng = np.random.default_rng(int(rs.integers(0, 1 << 30)))
    x = P.x0(n, rng); d = rng.integers(0, 2, n).astype(np.uint8); v = rng.standard_normal((n, D)).astype(np.float32)
    r = {"direction": torch.as_tensor(d.astype(np.float64)), "v_f": torch.as_tensor(v).double(), "v_b": torch.as_tensor(v).double()}
    dyn = P.oracle(torch.float64, temperature=temp)
    with torch.no_grad():
        acc = R._Acc(dyn); lo = float(R.loss_and_grads(torch.as_tensor(x).double(), dyn, r, 0.1, acc, kind=loss))
    l, de, gx, gv, Lx, px = E.run_emu(lib, P, x, v, d, 0.1, 1.0 / n, temp, loss_kind=LOSSES[loss])
    worst = 0.0
    for got, ref in ((gx, acc.x), (gv, acc.v)):
        for k in E.NAMES:
            a, b = got[k].astype(np.float64), ref[E.ORACLE_KEY[k]].numpy().reshape(got[k].shape)
            worst = max(worst, float(np.abs(a - b).max() / max(1e-12, np.abs(b).max())))
    ok = np.isfinite(l) and abs(l - lo) <= 5e-4 * max(1e-6, abs(lo)) + 1e-6 and worst < 5e-3
    if not ok: bad += 1
    print(("ok " if ok else "BAD"), kind, D, H, T, n, temp, loss, "loss %.6g vs %.6g" % (l, lo), "worst %.2e" % worst, flush=True)
print("bad:", bad)
