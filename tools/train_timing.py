"""Times one l2hmc_loss_grad batch (the training path) beside the sampling transition of the same shape.
Usage (GPU box): python tools/train_timing.py [--chains 16384] [--config c2_scg50] [--reps 3]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import util as U  # noqa: E402
from l2hmc_b200 import propose, training  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=1 << 14)
    ap.add_argument("--config", default="c2_scg50")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    P = U.Problem(regime="stress", **U.CONFIGS[a.config])
    dyn = P.product()
    x = torch.as_tensor(P.x0(a.chains, np.random.default_rng(0)), device="cuda:0")

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    l0 = dyn.launch_count
    ms_t = timed(lambda: training.loss_and_grads(dyn, x))
    launches = (dyn.launch_count - l0) // (a.reps + 1)
    ms_s = timed(lambda: propose(x, dyn))
    steps = a.chains * P.T
    print({"config": a.config, "chains": a.chains, "loss_grad_ms": round(ms_t, 3), "launches_per_loss_grad": launches,
           "train_leapfrog_steps_per_s": steps / ms_t * 1e3, "sampling_ms": round(ms_s, 3),
           "sampling_leapfrog_steps_per_s": steps / ms_s * 1e3, "ratio": ms_t / ms_s})


if __name__ == "__main__":
    main()
