#!/usr/bin/env python
"""Benchmark of the L2HMC sampling hot path (BASELINE.json metric: leapfrog-steps/sec on 50-d SCG).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one full transition -- propose (Lf augmented leapfrog steps for every chain) + Metropolis
accept -- over one batch of synthetic chains; x_next feeds the next step and stays in HBM.
Headline workload (config.workload): BASELINE.json configs[1], 50-d strongly correlated Gaussian, 2^18 chains
per GPU, Lf=10, width-100 S/T/Q nets, eps=0.1, synthetic perturbed-initialisation weights (not trained; see
l2hmc_b200/synthetic.py), in-kernel Philox.
value = chains x Lf x K x n_gpus / seconds (useful, selected-direction steps; the reference's discarded direction is not
counted).  One JSON line is printed by rank 0.  `other_configs` in the same line carries the other BASELINE.json
configurations (1, 3, 4, 5) measured the same way in the same run -- configs 4 and 5 with the GLOBAL chain count
BASELINE.json states, sharded over the ranks, one NCCL all-gather of the samples inside the timed region.

The product arm imports only l2hmc_b200 (workloads from l2hmc_b200/synthetic.py, parity spot checks against the committed
reference fixtures tests/golden/ref*.npz); oracle/ is executed only by the cpu_baseline leg and by --impl reference.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "leapfrog-steps/sec (chains×Lf/s) on 50-d SCG; accept-prob Δ vs ref"
UNIT = "leapfrog-steps/s"
D, H, LF = 50, 100, 10
CHAINS_PER_GPU = 1 << 18
GOLD = os.path.join(ROOT, "tests", "golden")


def mac_step(D, H, G):
    """Algorithmic MACs per leapfrog step per chain, one direction: four net calls of
    H (5 D + H + 2) MACs and ONE grad U of G MACs."""
    return 4 * H * (5 * D + H + 2) + G


def flop_step(D, H, G):
    return 2 * mac_step(D, H, G) + 56 * D   # + ~40 D elementwise + 16 D transcendentals


MAC_STEP = mac_step(D, H, D * D)            # 143,300
FLOP_STEP = flop_step(D, H, D * D)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("L2HMC_BENCH_SMI_MS", "50")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if not (t0 <= t <= t1 + 0.05):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_problem():
    from l2hmc_b200 import synthetic as S
    return S.SyntheticProblem(regime="stress", **S.CONFIGS["c2_scg50"])


# ---- reference arm: the reference's CPU path, restated (oracle/), on the host cores --------------------------------
def cpu_reference_run(steps, warmup, sample_chains):
    """The reference's own CPU path for this transition (both directions for every chain,
    utils/sampler.py:35-36), restated op-for-op in fp32 torch (oracle/l2hmc_oracle.py, pinned to the reference's own
    code by tests/test_reference_pin.py), all host threads."""
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import util as U  # tests/util.py: the oracle's view of the same synthetic problem (test infrastructure)
    P = U.Problem(regime="stress", **U.CONFIGS["c2_scg50"])
    dyn = P.oracle(torch.float32)
    rng = np.random.default_rng(0)
    n = sample_chains
    x = torch.as_tensor(P.x0(n, rng))
    # "all the host threads it can use": torch's intra-op pool thrashes when the container exposes more
    # logical CPUs than it may run on, so time one transition at a few pool sizes and keep the fastest.
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best = None

    def one(x):
        d = {"direction": torch.as_tensor(rng.integers(0, 2, n).astype(np.float32)),
             "v_f": torch.as_tensor(rng.standard_normal((n, D)).astype(np.float32)),
             "v_b": torch.as_tensor(rng.standard_normal((n, D)).astype(np.float32)),
             "u": torch.as_tensor(rng.random(n).astype(np.float32))}
        _, _, px, outs = U.O.propose(x, dyn, do_mh_step=True, **d)
        return outs[0], px
    for th in sorted({avail, min(avail, 64), min(avail, 32), min(avail, 16), min(avail, 8)}):
        torch.set_num_threads(th)
        t0 = time.perf_counter()
        one(x)
        dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (th, dt)
        elif dt > 1.5 * best[1]:
            break  # larger pools only get slower from here
    torch.set_num_threads(best[0])
    for _ in range(warmup):
        x, _ = one(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        x, px = one(x)
    dt = time.perf_counter() - t0
    return {"value": n * LF * steps / dt, "seconds": dt, "ms_per_step": 1e3 * dt / steps, "cores": torch.get_num_threads(),
            "sample": "%d chains x Lf=%d x %d transitions of the same 50-d SCG / width-100 workload" % (n, LF, steps)}


# ---- parity spot checks against the committed reference vectors (no oracle on the product arm) ----------------------
def _fixture_problem(z, meta):
    from l2hmc_b200 import synthetic as S
    if meta.get("vae"):
        P = S.SyntheticVaeProblem(**meta["kw"])
        if meta.get("weights_stored", True):
            P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
            P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
            P.dec_W = [z["decW_%d" % i] for i in range(len(P.dec_W))]
            P.dec_b = [z["decb_%d" % i] for i in range(len(P.dec_b))]
            if P.use_encoder:
                P.enc_W = [z["encW_%d" % i] for i in range(len(P.enc_W))]
                P.enc_b = [z["encb_%d" % i] for i in range(len(P.enc_b))]
    else:
        P = S.SyntheticProblem(regime=meta["regime"], **meta["kw"])
        P.xnet = {k[5:]: z[k] for k in z.files if k.startswith("xnet_")}
        P.vnet = {k[5:]: z[k] for k in z.files if k.startswith("vnet_")}
    P.mask = z["mask"]
    return P


def parity_vs_reference_fixture(relpath, dev, dyn=None):
    """Kernel vs the outputs of the unmodified reference on the same inputs (tests/golden/make_ref_golden.py)."""
    from l2hmc_b200 import propose
    path = os.path.join(GOLD, relpath)
    if not os.path.exists(path):
        return None
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    P = _fixture_problem(z, meta)
    dyn = dyn or P.product(device=dev.index)
    g = lambda a: torch.as_tensor(np.asarray(a)).to(dev)  # noqa: E731
    d = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    v_sel = np.where(d["dir"][:, None] != 0, d["v_f"], d["v_b"]).astype(np.float32)
    aux = g(d["aux"]) if "aux" in d else None
    Lx, Lv, px, outs = propose(g(d["x"]), dyn, init_v=g(v_sel), aux=aux, do_mh_step=True,
                               rng={"direction": g(d["dir"]), "v": g(v_sel), "u": g(d["u"])})
    torch.cuda.synchronize(dev)

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))
    px = px.cpu().numpy()
    return {"fixture": relpath, "chains": int(d["x"].shape[0]),
            "Lx": rel(Lx.cpu().numpy(), z["out_Lx"]), "Lx_ref_fp32": rel(z["out32_Lx"], z["out_Lx"]),
            "Lv": rel(Lv.cpu().numpy(), z["out_Lv"]), "Lv_ref_fp32": rel(z["out32_Lv"], z["out_Lv"]),
            "px_max": float(np.max(np.abs(px - z["out_px"]))), "px_max_ref_fp32": float(np.max(np.abs(z["out32_px"] - z["out_px"]))),
            "px_mean": abs(float(px.astype(np.float64).mean()) - float(z["out_px"].mean())),
            "mean_accept_prob_ref": float(z["out_px"].mean())}


def ncu_traffic(kernel_regex, profile_glob="r02_*ncu_raw.csv"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, PARSED from the committed `ncu --set full`
    capture of this same workload under profiles/ (ncu -i ... --page raw --csv); None when no capture is committed."""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", profile_glob)), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
            names, units = rows[head], rows[head + 1]
            ik, ir, iw = names.index("Kernel Name"), names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[head + 2:]:
                if len(r) > max(ir, iw) and re.search(kernel_regex, r[ik]):
                    rd = float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0)
                    wr = float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0)
                    return rd + wr, os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def fma_peak_tflops(sm_mhz):
    return 148 * 128 * 2 * sm_mhz * 1e6 / 1e12  # fp32 FMA pipe


class Runner:
    """Times K transitions of one configuration on this rank's shard: CUDA events on the launch stream, barrier +
    synchronize on both sides, max over ranks; one all-gather of the samples at the end when sharded."""

    def __init__(self, dev, rank, world, sampler):
        self.dev, self.rank, self.world, self.sampler = dev, rank, world, sampler

    def run(self, P, n_local, lo, steps, warmup, aux=None, kernel="auto", gather_total=None, dyn=None):
        import torch.distributed as dist
        from l2hmc_b200 import _lib
        from l2hmc_b200.sharding import all_gather_chains
        dev = self.dev
        dyn = dyn or P.product(device=dev.index, seed=1, kernel=kernel)
        rng = np.random.default_rng(100 + self.rank)
        x = torch.as_tensor(P.x0(n_local, rng) if hasattr(P, "x0") else rng.standard_normal((n_local, P.D)).astype(np.float32)).to(dev)
        n = n_local
        Dd = P.D

        def outset():
            return {"Lx": torch.empty((n, Dd), dtype=torch.float32, device=dev), "Lv": None,
                    "px": torch.empty((n,), dtype=torch.float32, device=dev),
                    "x_next": torch.empty((n, Dd), dtype=torch.float32, device=dev),
                    "accepted": torch.empty((n,), dtype=torch.uint8, device=dev)}
        outs = [outset(), outset()]
        stats = torch.zeros(2, dtype=torch.float64, device=dev)
        dir_mode = _lib.DIR_FORWARD if P.hmc else _lib.DIR_RANDOM

        def step(x, counter):
            return dyn._transition(x, dir_mode=dir_mode, do_mh=True, counter=counter, chain_offset=lo, want_v=False,
                                   out=outs[counter & 1], aux=aux, stats=stats)
        ctr = 0
        for _ in range(warmup):
            x = step(x, ctr)["x_next"]
            ctr += 1
        gathered = None
        if gather_total is not None and self.world > 1:
            gathered = torch.empty((gather_total, Dd), dtype=torch.float32, device=dev)
            for _ in range(2):
                all_gather_chains(x, gather_total, out=gathered)
        stats.zero_()
        dyn.timing_enable(True)
        launches0 = dyn.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t_wall0 = time.time()
        e0.record()
        for _ in range(steps):
            x = step(x, ctr)["x_next"]
            ctr += 1
        if gathered is not None:
            all_gather_chains(x, gather_total, out=gathered)  # the single NCCL all-gather of samples at the end
        e1.record()
        torch.cuda.synchronize(dev)
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        launches = dyn.launch_count - launches0
        kern_ms, kern_cnt = dyn.timing_read()
        dyn.timing_enable(False)
        st = stats.cpu().numpy().copy()
        if self.world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            lt = torch.tensor([float(launches), st[0], st[1]], device=dev, dtype=torch.float64)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches, st = int(lt[0]), np.array([float(lt[1]), float(lt[2])])
            dist.barrier()
        clocks = self.sampler.summary(t_wall0, t_wall1) if self.sampler else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"ms": ms, "launches": launches, "kernel_ms": kern_ms, "kernel_launches_timed": kern_cnt, "clocks": clocks,
                "kernel": dyn.kernel_name, "x": x, "ctr": ctr, "dyn": dyn, "sum_p": st[0], "n_accept": st[1],
                "window": (t_wall0, t_wall1)}


def other_config(runner, name, peaks, quick):
    """One BASELINE.json configuration besides the headline one, measured like it.  Returns a dict for `other_configs`."""
    from l2hmc_b200 import synthetic as S
    from l2hmc_b200.sharding import shard_bounds
    world, rank, dev = runner.world, runner.rank, runner.dev
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    spec = {
        # name: (BASELINE.json config, problem, GLOBAL chains, sharded over ranks?, steps, parity fixture, G MACs of grad U)
        "c1_n200": ("configs[0]: 2-d SCG, 200 chains, Lf=10, width-10 nets (SCGExperiment.ipynb settings)", "c1_scg2", 200, False, 50,
                    "ref_c1_scg2_n200_stress.npz", 4),
        "c1_n2e18": ("configs[0] nets and target at 2^18 chains per GPU", "c1_scg2", 1 << 18, False, 20, None, 4),
        "c3_mog2": ("configs[2]: 2-d mixture of 2 Gaussians (covariance 0.1 I), 2^18 chains per GPU, Lf=25, width-10 nets", "c3_mog2",
                    1 << 18, False, 10, "ref_c3_mog2_n256_stress.npz", 8),
        "c4_rw32": ("configs[3]: 32-d rough well (easy=True), 2^20 chains GLOBAL sharded over the ranks, Lf=10, width-100 nets, "
                    "final all-gather", "c4_rw32", 1 << 20, True, 5, "ref_c4_rw32_n128_stress.npz", 0),
        "c5_vae": ("configs[4]: 784-d MNIST-VAE posterior target (mnist_vae.py decoder energy, 50-d latent, aux-conditioned "
                   "width-200 nets), 2^16 chains GLOBAL sharded over the ranks, Lf=15, final all-gather", "c5_vae_full", 1 << 16, True, 3,
                   os.path.join("ref", "c5_vae_full_n32.npz"), 2 * (50 * 1024 + 1024 * 1024 + 1024 * 784)),
    }[name]
    label, cfg, n_global, sharded, steps, fixture, G = spec
    if quick:
        steps = max(2, steps // 3)
    vae = cfg.startswith("c5")
    P = S.SyntheticVaeProblem(**S.VAE_CONFIGS[cfg]) if vae else S.SyntheticProblem(regime="stress", **S.CONFIGS[cfg])
    if sharded:
        lo, hi = shard_bounds(n_global, rank, world)
        n_local, n_total, scaling = hi - lo, n_global, "strong"
    else:
        n_local, lo, n_total, scaling = n_global, rank * n_global, n_global * world, "weak"
    aux = None
    if vae:
        P.x0 = lambda n, rng: rng.standard_normal((n, P.D)).astype(np.float32)   # latent prior, like init_x = latent_q
        aux = torch.as_tensor((np.random.default_rng(7 + rank).random((n_local, P.aux_dim)) < 0.5).astype(np.float32)).to(dev)
    # long enough a timed region for the 50 ms clock sampler to see it (>= ~0.4 s): calibrate on a few steps first
    cal = runner.run(P, n_local, lo, 2, 3, aux=aux, gather_total=None)
    per = max(cal["ms"] / 2.0, 1e-3)
    steps = int(min(20000, max(steps, np.ceil((150.0 if quick else 400.0) / per))))
    if world > 1:  # every rank must time the same number of steps
        import torch.distributed as dist
        t = torch.tensor([steps], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steps = int(t[0])
    r = runner.run(P, n_local, lo, steps, 3, aux=aux, gather_total=n_total if sharded else None, dyn=cal["dyn"])
    value = n_total * P.T * steps / (r["ms"] * 1e-3)
    flops = n_local * P.T * flop_step(P.D, P.H, G)
    kernel = r["kernel"]
    tensor = kernel.startswith("tc") or kernel.startswith("layered_tc")
    achieved = flops / (r["kernel_ms"] * 1e-3) / 1e12 if r["kernel_ms"] and r["kernel_ms"] > 0 and not kernel.startswith("layered") else \
        flops * steps / (r["ms"] * 1e-3) / 1e12   # the layered engine is a launch sequence: use the whole step
    if tensor:
        f16 = kernel.endswith("f16")
        peak = float(peaks["bf16_tflops_sustained"]) / (1.0 if f16 else 2.0)
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "note": "algorithmic fp32 FLOP once; each product costs 3 %s MMAs" % ("fp16" if f16 else "tf32")}
    else:
        peak = fma_peak_tflops(sm_max)
        roof = {"bound": "fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak}
    out = {"name": name, "workload": label, "value": value, "unit": UNIT, "n_gpus": world, "scaling": scaling,
           "global_chains": n_total, "chains_per_gpu": n_local, "x_dim": P.D, "width": P.H, "Lf": P.T, "steps": steps,
           "warmup": 3, "ms_per_step": r["ms"] / steps, "kernel": kernel, "kernel_ms": r["kernel_ms"], "gpu_launches": r["launches"],
           "mac_per_leapfrog_step": mac_step(P.D, P.H, G), "roofline": roof, "clocks": r["clocks"],
           "mean_accept_prob": r["sum_p"] / max(1.0, n_total * steps), "accept_rate": r["n_accept"] / max(1.0, n_total * steps),
           "allgather_in_timed_region": bool(sharded and world > 1)}
    if rank == 0 and fixture:
        out["parity_vs_reference"] = parity_vs_reference_fixture(fixture, dev, dyn=None)
    return out


def training_config(runner, name, n, steps, quick):
    """One iteration of the notebook's training loop (SCGExperiment.ipynb:254-270: the objective on the current samples
    and on fresh noise, its gradient through the unrolled leapfrog, Adam, Metropolis output fed back) per step, through
    l2hmc_b200.training (l2hmc_loss_grad).  Rank 0 only: the path is measured on one GPU."""
    from l2hmc_b200 import synthetic as S, training
    dev = runner.dev
    P = S.SyntheticProblem(regime="init", **S.CONFIGS[name])   # training starts from the notebook's initialisation
    dyn = P.product(device=dev.index, seed=3)
    opt = training.Adam(dyn)
    samples = torch.as_tensor(P.x0(n, np.random.default_rng(5))).to(dev)
    for _ in range(3):
        samples = training.train_step(dyn, opt, samples)["samples"]
    torch.cuda.synchronize(dev)
    steps = max(3, steps // 3) if quick else steps
    launches0 = dyn.launch_count
    t_wall0 = time.time()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = training.train_step(dyn, opt, samples)
        samples = out["samples"]
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    t_wall1 = time.time()
    clocks = runner.sampler.summary(t_wall0, t_wall1) if runner.sampler else None
    return {"name": "train_" + name, "workload": "notebook training loop (loss on samples + noise batch, reverse sweep, Adam), %d chains, "
            "x_dim %d, width %d, Lf=%d" % (n, P.D, P.H, P.T), "chains": n, "steps": steps,
            "training_steps_per_s": steps / dt, "ms_per_step": 1e3 * dt / steps,
            # both batches (x and z) run forward and reverse sweeps of Lf leapfrog steps per chain
            "leapfrog_steps_per_s_forward_equivalent": 2 * n * P.T * steps / dt,
            "loss": out["loss"], "mean_accept_prob": float(out["px"].mean()), "clocks": clocks,
            "timing": "host wall clock around train_step (includes the host-side Adam write-back), device synchronised"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--cpu-chains", type=int, default=1 << 14,
                    help="chains of the CPU arm's bounded sample (its throughput still grows a little with the sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="headline workload only")
    ap.add_argument("--quick", action="store_true", help="fewer timed steps for the other configurations")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = {"workload": "BASELINE configs[1]: 50-d strongly correlated Gaussian, 2^18 chains per GPU, Lf=10, "
                            "width-100 S/T/Q nets, eps=0.1", "x_dim": D, "width": H, "Lf": LF}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.steps, args.warmup, args.cpu_chains)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload, sample_chains=args.cpu_chains),
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    from l2hmc_b200.sharding import init_distributed
    rank, world, local = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    P = build_problem()
    n = args.chains
    n_total = n * world
    lo = rank * n

    # L2HMC_BENCH_NO_SMI=1: development switch to measure what the nvidia-smi polling itself costs (it does perturb the GPU)
    sampler = ClockSampler(local) if rank == 0 and os.environ.get("L2HMC_BENCH_NO_SMI") != "1" else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    runner = Runner(dev, rank, world, sampler)

    # ---- parity spot check outside the timed region: the kernel against the unmodified reference's outputs ----------
    dyn = P.product(device=local, seed=1, kernel=args.kernel)
    rep = parity_vs_reference_fixture("ref_c2_scg50_n128_stress.npz", dev) if rank == 0 else None

    # ---- device-resident throughput: all ranks enter the timed region together; one all-gather of samples at the end
    r = runner.run(P, n, lo, args.steps, args.warmup, kernel=args.kernel, gather_total=n_total, dyn=dyn)
    ms, launches, kern_ms, kern_cnt, clocks = r["ms"], r["launches"], r["kernel_ms"], r["kernel_launches_timed"], r["clocks"]
    x, ctr = r["x"], r["ctr"]
    value = n_total * LF * args.steps / (ms * 1e-3)
    mean_px = r["sum_p"] / (n_total * args.steps)       # accept statistics reduced inside the kernel (stats output)
    accept_rate = r["n_accept"] / (n_total * args.steps)

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region --------
    hx = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
    hx.copy_(x)
    out = {"Lx": None, "Lv": None, "px": torch.empty((n,), dtype=torch.float32, pin_memory=True).numpy(),
           "x_next": torch.empty((n, D), dtype=torch.float32, pin_memory=True).numpy(),
           "accepted": torch.empty((n,), dtype=torch.uint8, pin_memory=True).numpy()}
    hxn = hx.numpy()
    hstats = np.zeros(2, np.float64)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        dyn.transition_host(hxn, counter=ctr, chain_offset=lo, out=out)
        ctr += 1
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cur = hxn
    for _ in range(e2e_steps):
        dyn.transition_host(cur, counter=ctr, chain_offset=lo, out=out, stats=hstats)  # synchronous: returns after D2H
        cur = out["x_next"]
        ctr += 1
    e2e_s = time.perf_counter() - t0
    e2e_rank_s = e2e_s
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = n_total * LF * e2e_steps / e2e_s
    h2d = n * D * 4
    d2h = n * D * 4 + n * 4 + n

    # ---- the other BASELINE.json configurations, same run, same clocks record ------------------------------------------
    others = []
    if not args.no_other_configs:
        del x, hx
        torch.cuda.empty_cache()
        for name in ("c1_n200", "c1_n2e18", "c3_mog2", "c4_rw32", "c5_vae"):
            try:
                others.append(other_config(runner, name, measured_peaks()[0], args.quick))
            except Exception as e:  # noqa: BLE001 -- a failing side configuration must not lose the headline line
                others.append({"name": name, "error": "%s: %s" % (type(e).__name__, str(e)[:300])})
            torch.cuda.empty_cache()
    train = []
    if not args.no_other_configs and rank == 0:
        for name, nn, st in (("c1_scg2", 200, 60), ("c2_scg50", 4096, 12)):
            try:
                train.append(training_config(runner, name, nn, st, args.quick))
            except Exception as e:  # noqa: BLE001
                train.append({"name": "train_" + name, "error": "%s: %s" % (type(e).__name__, str(e)[:300])})

    if sampler:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, src = measured_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fma_peak = fma_peak_tflops(sm_max)  # fp32 FMA pipe at the max SM clock
    launch_flops = n * LF * FLOP_STEP
    achieved = launch_flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else None
    roofline = {"bound": "fma", "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s",
                "frac": (achieved / fma_peak) if achieved else None, "traffic": None,
                "peak_source": "148 SM x 128 fp32 lanes x 2 flop x sm_max_mhz from MEASURED_PEAKS.json (%s); "
                               "the path is FMA-pipe bound, not HBM or tensor (SURVEY.md section 8d)" % src,
                "kernel": dyn.kernel_name, "kernel_ms": kern_ms, "kernel_launches_timed": kern_cnt,
                "algorithmic_flops_per_launch": launch_flops,
                "algorithmic_hbm_bytes_per_launch": n * (3 * D * 4 + 4 + 1),
                "frac_of_measured_tf32_tensor": (achieved / (float(peaks["bf16_tflops_sustained"]) / 2)) if achieved else None}
    if clocks.get("sm_mhz"):
        roofline["frac_at_clock_under_load"] = achieved / fma_peak_tflops(clocks["sm_mhz"]) if achieved else None
    if dyn.kernel_name.startswith("tc") and n == CHAINS_PER_GPU:
        # DRAM bytes of one launch of the dominant kernel, parsed from the committed `ncu --set full` capture of this same
        # workload (below the algorithmic bytes when the 126 MB L2 still holds part of the written lines at the end)
        roofline["traffic"], roofline["traffic_source"] = ncu_traffic(r"tc_transition_kernel_s")
    if dyn.kernel_name.startswith("tc") and achieved:
        # the GEMMs run on the tensor pipe with an error-compensated split: three MMAs per fp32-accurate product, tf32
        # (K = 8 per MMA) or fp16 pairs (K = 16 per MMA, twice the rate) -- l2hmc_kernel_name says which the launch used.
        # Denominator: the measured sustained dense rate of that input type (bf16 figure for fp16, half of it for tf32).
        f16 = dyn.kernel_name == "tc_3xf16"
        t_peak = float(peaks["bf16_tflops_sustained"]) / (1.0 if f16 else 2.0)
        roofline.update({"bound": "tensor", "peak": t_peak, "frac": achieved / t_peak,
                         "frac_of_fma_roofline": achieved / fma_peak,
                         "peak_source": "dense %s = MEASURED_PEAKS.json bf16_tflops_sustained%s (%s); algorithmic fp32 FLOP counted "
                                        "once although each product costs three MMAs over padded shapes (error-compensated split for "
                                        "1e-5 parity): see executed_tensor_tflops for what the tensor pipe really does; the kernel "
                                        "holds the maximum SM clock while the GEMM behind the measured peak runs power-capped"
                                        % ("fp16" if f16 else "TF32", "" if f16 else " / 2", src),
                         "tensor_mma_per_product": 3, "operand_split": "fp16 x3" if f16 else "tf32 x3"})
        # what the tensor pipe executes per leapfrog step and 128-chain tile: 3 MMAs per product over the padded shapes
        # (K 100 -> 104 (tf32) / 112 (fp16), N 100 -> 112, heads 150 -> 96 + 80 columns, grad 50 -> K 56 / 64 x N 64)
        kp, kg = (112, 64) if f16 else (104, 56)
        mac_tile_step = 3 * (4 * 128 * kp * (112 + 112 + 96 + 80) + 128 * kg * 64)
        executed = 2.0 * mac_tile_step * LF * (n / 128.0) / (kern_ms * 1e-3) / 1e12
        roofline["executed_tensor_tflops"] = executed
        roofline["executed_frac_of_measured_peak"] = executed / t_peak
        roofline["executed_frac_at_sm_max_clock"] = executed / (148 * (4096 if f16 else 2048) * 2 * sm_max * 1e6 / 1e12)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rr = cpu_reference_run(steps=5, warmup=1, sample_chains=args.cpu_chains)
        cpu = {"value": rr["value"], "unit": UNIT, "cores": rr["cores"], "kind": "port", "sample": rr["sample"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload, chains_per_gpu=n, global_chains=n_total, parallelism="chains sharded x%d" % world,
                           l2="per-step working set (x, Lx, x_next, px) = %.0f MB > 126 MB L2; no flush" % (3 * n * D * 4 / 1e6),
                           kernel=dyn.kernel_name, weights="synthetic perturbed-initialisation ('stress' regime of l2hmc_b200/synthetic.py), not trained",
                           mean_accept_prob=mean_px, accept_rate=accept_rate,
                           accept_statistics="sum_p / n_accept reduced in the kernel (l2hmc_transition_args.stats)"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "l2hmc_transition_host (pinned host buffers)",
                    "rank0_h2d_gbs": h2d * e2e_steps / e2e_rank_s / 1e9, "rank0_d2h_gbs": d2h * e2e_steps / e2e_rank_s / 1e9,
                    "mean_accept_prob": hstats[0] / (n * e2e_steps)},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": rep,
            "other_configs": others,
            "training": train}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
