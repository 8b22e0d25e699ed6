#!/usr/bin/env python
"""GPU diagnostic: parity numbers for every configuration and a quick timing sweep.
Writes human-readable lines to stdout (redirect into gpurun_out/)."""
import argparse
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import util as U  # noqa: E402


def fmt(rep):
    return " ".join("%s=%.3g" % (k, v) if isinstance(v, float) else "%s=%s" % (k, v) for k, v in rep.items())


def parity(quick, kernel):
    cases = [("c1_scg2", 200), ("c2_scg50", 192), ("c3_mog2", 256), ("c4_rw32", 192), ("c4_rw32_hard", 192), ("funnel3", 128)]
    if quick:
        cases = cases[:2]
    if kernel == "tc":  # what the tensor-core kernel covers
        cases = [("c2_scg50", 192), ("c2_scg50", 500), ("c4_rw32", 192), ("c4_rw32_hard", 192)]
    for name, n in cases:
        for regime in ("init", "stress"):
            try:
                P = U.Problem(regime=regime, **U.CONFIGS[name])
                dyn = P.product(kernel=kernel)
                rep, _ = U.parity_report(P, n, dyn=dyn)
                print("PARITY %-13s %-6s %s kernel=%s" % (name, regime, fmt(rep), dyn.kernel_name), flush=True)
            except Exception:
                print("PARITY %s %s FAILED" % (name, regime))
                traceback.print_exc()
    for kind, D in (() if kernel == "tc" else (("gaussian", 2), ("gaussian", 50), ("roughwell", 32))):
        try:
            P = U.Problem(kind=kind, D=D, T=10, eps=0.05, hmc=True)
            rep, _ = U.parity_report(P, 192, dyn=P.product(kernel=kernel))
            print("PARITY hmc-%s-%d %s" % (kind, D, fmt(rep)), flush=True)
        except Exception:
            print("PARITY hmc %s FAILED" % kind)
            traceback.print_exc()


def timing(kernel):
    from l2hmc_b200 import _lib
    cfgs = (("c2_scg50", 1 << 18), ("c4_rw32", 1 << 17), ("c1_scg2", 1 << 18), ("c3_mog2", 1 << 18))
    if kernel == "tc":
        cfgs = cfgs[:2]
    for name, n in cfgs:
        try:
            P = U.Problem(regime="stress", **U.CONFIGS[name])
            dyn = P.product(kernel=kernel)
            x = torch.as_tensor(P.x0(n, np.random.default_rng(0))).cuda()
            for _ in range(2):
                o = dyn._transition(x, dir_mode=_lib.DIR_RANDOM, do_mh=True, want_v=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = int(os.environ.get("L2HMC_DIAG_REPS", "5"))
            e0.record()
            for _ in range(reps):
                o = dyn._transition(o["x_next"], dir_mode=_lib.DIR_RANDOM, do_mh=True, want_v=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print("TIMING %-10s n=%d T=%d kernel=%s ms=%.3f steps/s=%.4g mean_p=%.3f" %
                  (name, n, P.T, dyn.kernel_name, ms, n * P.T / (ms * 1e-3), float(o["px"].mean())), flush=True)
            if dyn.kernel_name.startswith("tc"):
                import ctypes as C
                buf = (C.c_int64 * 56)()
                dyn._chk(dyn._lib.l2hmc_debug_counters(dyn._ctx, buf, 56))
                c = list(buf)
                if sum(c[24:56]) > 0:   # kernel_tc_s.cuh, -DL2HMC_TC_PHASE_ACCOUNTING: compute threads per epilogue kind
                    kn = ("embed", "hidden", "headsV_p0", "headsV_p1", "headsX_p0", "headsX_p1", "grad")
                    calls = (4, 4, 2, 2, 2, 2, 1)   # per leapfrog step
                    for g in (0, 1):
                        o = 24 + 16 * g
                        print("TCEPI   %-10s thread %d of a chain, cycles per call (work + wait for the accumulator): " % (name, g) + "  ".join(
                            "%s %.0f+%.0f" % (kn[k], c[o + k] / (calls[k] * P.T), c[o + 8 + k] / (calls[k] * P.T)) for k in range(7)), flush=True)
                if c[2] > 0 and c[4] > 0:
                    print("TCPHASE %-10s CTA0 cycles: issuer total=%d wait_A=%.1f%% wait_TMA=%.1f%% issue/MMA=%.1f%% | compute total=%d "
                          "wait_acc=%.1f%% | gemms=%d cycles/gemm=%.0f" %
                          (name, c[2], 100.0 * c[0] / c[2], 100.0 * c[1] / c[2], 100.0 * (c[2] - c[0] - c[1]) / c[2], c[4],
                           100.0 * c[3] / c[4], c[5], c[2] / max(c[5], 1)), flush=True)
                    if sum(c[18:23]) > 0:
                        names = ("grad", "embed", "hidden", "heads_a", "heads_b")
                        print("TCKIND  %-10s per GEMM (cycles): " % name + "  ".join(
                            "%s: wait_A=%.0f wait_TMA=%.0f (x%d)" % (names[k], c[8 + k] / max(c[18 + k], 1), c[13 + k] / max(c[18 + k], 1), c[18 + k])
                            for k in range(5)), flush=True)
        except Exception:
            print("TIMING %s FAILED" % name)
            traceback.print_exc()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-timing", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--kernel", default="auto")
    a = ap.parse_args()
    print("device:", torch.cuda.get_device_name(0), flush=True)
    if not a.no_parity:
        parity(a.quick, a.kernel)
    if not a.no_timing and not a.quick:
        timing(a.kernel)
