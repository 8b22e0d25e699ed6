#!/usr/bin/env python
"""One gradient evaluation of the notebook objective (l2hmc_loss_grad) on BASELINE config 2's shape, for ncu launch lists.
usage: profile_training.py [chains] [config]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from l2hmc_b200 import synthetic as S, training  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
name = sys.argv[2] if len(sys.argv) > 2 else "c2_scg50"
P = S.SyntheticProblem(regime="init", **S.CONFIGS[name])
dyn = P.product(seed=3)
x = torch.as_tensor(P.x0(n, np.random.default_rng(5))).cuda()
out = training.loss_and_grads(dyn, x)
torch.cuda.synchronize()
print(name, n, dyn.launch_count, float(out["loss"]) if "loss" in out else None)
