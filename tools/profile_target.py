#!/usr/bin/env python
"""One BASELINE configuration, a few transitions, nothing else: the process ncu wraps for the per-kernel captures
(tools/run_profiles.sh).  usage: profile_target.py c1|c2|c3|c4|c5 [steps] [chains]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from l2hmc_b200 import _lib, synthetic as S  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
name, n = {"c1": ("c1_scg2", 1 << 18), "c2": ("c2_scg50", 1 << 18), "c3": ("c3_mog2", 1 << 18), "c4": ("c4_rw32", 1 << 18),
           "c5": ("c5_vae_full", 1 << 14)}[cfg]
if len(sys.argv) > 3:
    n = int(sys.argv[3])
aux = None
if cfg == "c5":
    P = S.SyntheticVaeProblem(**S.VAE_CONFIGS[name])
    x = torch.randn((n, P.D), device="cuda")
    aux = (torch.rand((n, P.aux_dim), device="cuda") < 0.5).float()
else:
    P = S.SyntheticProblem(regime="stress", **S.CONFIGS[name])
    x = torch.as_tensor(P.x0(n, np.random.default_rng(0))).cuda()
dyn = P.product(seed=1)
for t in range(steps):
    o = dyn._transition(x, dir_mode=_lib.DIR_RANDOM, do_mh=True, want_v=False, aux=aux)
    x = o["x_next"]
torch.cuda.synchronize()
print(cfg, dyn.kernel_name, float(o["px"].mean()))
