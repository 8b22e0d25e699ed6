#!/usr/bin/env python
"""Timing of the tensor-core kernels over (x_dim, width) shapes: the shape-specialised / run-time-shape instantiations of
kernel_tc_s.cuh against the generic kernel of kernel_tc.cuh (L2HMC_TC_GENERIC=1), 2^17 chains, Lf = 10.
    python tools/tc_shapes_timing.py > gpurun_out/r02_tc_shapes_timing.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from l2hmc_b200 import _lib, synthetic as S  # noqa: E402

n = 1 << 17
print("# %s, %d chains, Lf=10, CUDA events over 10 transitions after 3 warm-up" % (torch.cuda.get_device_name(0), n))
print("# %-10s %5s %6s | %10s %12s | %10s %12s | %s" % ("target", "x_dim", "width", "tc_s ms", "steps/s", "generic ms", "steps/s", "speed-up"))
for kind, D, H in (("gaussian", 50, 100), ("roughwell", 32, 100), ("gaussian", 50, 64), ("gaussian", 40, 100), ("gaussian", 32, 64), ("gaussian", 20, 64),
                   ("gaussian", 16, 32), ("roughwell", 24, 48), ("gaussian", 8, 100), ("gaussian", 52, 104)):
    row = []
    for generic in (False, True):
        if generic:
            os.environ["L2HMC_TC_GENERIC"] = "1"
        else:
            os.environ.pop("L2HMC_TC_GENERIC", None)
        kw = dict(easy=True) if kind == "roughwell" else {}
        P = S.SyntheticProblem(kind=kind, D=D, H=H, T=10, eps=0.1, regime="stress", **kw)
        dyn = P.product(kernel="tc", seed=1)
        x = torch.as_tensor(P.x0(n, np.random.default_rng(0))).cuda()
        for _ in range(3):
            x = dyn._transition(x, dir_mode=_lib.DIR_RANDOM, do_mh=True, want_v=False)["x_next"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            x = dyn._transition(x, dir_mode=_lib.DIR_RANDOM, do_mh=True, want_v=False)["x_next"]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        row.append((ms, n * 10 / (ms * 1e-3), dyn.kernel_name))
    os.environ.pop("L2HMC_TC_GENERIC", None)
    print("  %-10s %5d %6d | %10.3f %12.4g | %10.3f %12.4g | %.2fx  (%s)" % (kind, D, H, row[0][0], row[0][1], row[1][0], row[1][1], row[1][0] / row[0][0], row[0][2]))
