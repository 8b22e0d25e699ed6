"""Config 5 (BASELINE.json configs[4]) on one GPU: the MNIST-VAE posterior target at the reference's layer sizes
(latent 50, decoder 1024-1024-784, aux encoder 512-512-200, width-200 nets, Lf=15; mnist_vae.py:104-178) on the
layered engine.  Prints one JSON line: leapfrog-steps/s, ms per transition, launches, a small parity report against the
CPU oracle (test infrastructure), and the fraction of the fp32-FMA roofline.

    python tools/vae_bench.py [--chains 65536] [--steps 3] [--warmup 1]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--parity-chains", type=int, default=64)
    args = ap.parse_args()
    import util as U
    from l2hmc_b200 import _lib
    kw = U.VAE_CONFIGS["c5_vae_full"]
    P = U.VaeProblem(**kw)
    dyn = P.product(seed=1)
    rep, _ = U.parity_report(P, args.parity_chains, dyn=dyn)
    n = args.chains
    d = P.draws(n, seed=2)
    x = torch.as_tensor(d["x"]).cuda()
    aux = torch.as_tensor(d["aux"]).cuda()
    ctr = 0
    for _ in range(args.warmup):
        x = dyn._transition(x, dir_mode=_lib.DIR_RANDOM, do_mh=True, counter=ctr, aux=aux, want_v=False)["x_next"]
        ctr += 1
    torch.cuda.synchronize()
    l0 = dyn.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        o = dyn._transition(x, dir_mode=_lib.DIR_RANDOM, do_mh=True, counter=ctr, aux=aux, want_v=False)
        x = o["x_next"]
        ctr += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    D, H, T = P.D, P.H, P.T
    dec = P.dec_w
    mac_dec = sum(dec[i] * dec[i + 1] for i in range(len(dec) - 1))
    mac_net = H * (5 * D + H + 2)
    mac_step = 4 * mac_net + 2 * mac_dec          # one grad U per step: forward + reverse pass of the decoder
    mac_once = sum(P.enc_w[i] * P.enc_w[i + 1] for i in range(len(P.enc_w) - 1)) + 2 * mac_dec  # aux encoding + first grad U
    flops = 2.0 * n * (T * mac_step + mac_once)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fma_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    print(json.dumps({
        "workload": "BASELINE configs[4]: MNIST-VAE posterior target, %d chains, Lf=%d, width %d, 1 GPU" % (n, T, H),
        "kernel": dyn.kernel_name, "ms_per_transition": ms, "leapfrog_steps_per_s": n * T / (ms * 1e-3),
        "launches_per_transition": (dyn.launch_count - l0) / args.steps,
        "mac_per_step_per_chain": mac_step, "achieved_tflops": flops / (ms * 1e-3) / 1e12,
        "frac_of_fma_roofline": flops / (ms * 1e-3) / fma_peak,
        "mean_accept_prob": float(o["px"].mean()), "parity": rep}))


if __name__ == "__main__":
    main()
