#!/usr/bin/env python
"""Benchmark of the L2HMC sampling hot path (BASELINE.json metric: leapfrog-steps/sec on 50-d SCG).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one full transition -- propose (Lf augmented leapfrog steps for every chain) + Metropolis
accept -- over one batch of synthetic chains; x_next feeds the next step and stays in HBM.
Workload (config.workload): BASELINE.json configs[1], 50-d strongly correlated Gaussian, 2^18 chains
per GPU, Lf=10, width-100 S/T/Q nets, eps=0.1, trained-like synthetic weights, in-kernel Philox.
value = chains x Lf x K x n_gpus / seconds (useful, selected-direction steps; the reference's
discarded direction is not counted).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "leapfrog-steps/sec (chains×Lf/s) on 50-d SCG; accept-prob Δ vs ref"
UNIT = "leapfrog-steps/s"
D, H, LF = 50, 100, 10
CHAINS_PER_GPU = 1 << 18
# algorithmic work per leapfrog step per chain, one direction (SURVEY.md section 8d / BASELINE.md section 4)
MAC_STEP = 4 * H * (5 * D + H + 2) + D * D          # 143,300
FLOP_STEP = 2 * MAC_STEP + 56 * D                   # + ~40 D elementwise + 16 D transcendentals


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
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
    import util as U
    return U.Problem(regime="stress", **U.CONFIGS["c2_scg50"]), U


def cpu_reference_run(steps, warmup, sample_chains):
    """The reference's own CPU path for this transition (both directions for every chain,
    utils/sampler.py:35-36), restated op-for-op in fp32 torch (oracle/l2hmc_oracle.py), all host threads."""
    P, U = build_problem()
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
    from l2hmc_b200 import _lib
    from l2hmc_b200.sharding import all_gather_chains, init_distributed
    rank, world, local = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    P, U = build_problem()
    dyn = P.product(device=local, seed=1, kernel=args.kernel)
    n = args.chains
    n_total = n * world
    lo = rank * n
    rng = np.random.default_rng(100 + rank)
    x = torch.as_tensor(P.x0(n, rng)).to(dev)

    # two output sets used alternately: x_next of one step is the x of the next, nothing is allocated per step
    def outset():
        return {"Lx": torch.empty((n, D), dtype=torch.float32, device=dev), "Lv": None,
                "px": torch.empty((n,), dtype=torch.float32, device=dev),
                "x_next": torch.empty((n, D), dtype=torch.float32, device=dev),
                "accepted": torch.empty((n,), dtype=torch.uint8, device=dev)}
    outs = [outset(), outset()]

    def step(x, counter):
        return dyn._transition(x, dir_mode=_lib.DIR_RANDOM, do_mh=True, counter=counter, chain_offset=lo, want_v=False,
                               out=outs[counter & 1])

    # ---- parity spot check outside the timed region (accept-prob delta vs the oracle) -----------------
    rep = None
    if rank == 0:
        rep, _ = U.parity_report(P, 256, dyn=dyn)

    # ---- device-resident throughput ----------------------------------------------------------------------
    ctr = 0
    for _ in range(args.warmup):
        x = step(x, ctr)["x_next"]
        ctr += 1
    if world > 1:
        # the communicator, its channels and the gather buffer are set up by the first collective: do that
        # in the warm-up, like every other first-call cost
        gathered = torch.empty((n_total, D), dtype=torch.float32, device=dev)
        for _ in range(2):
            all_gather_chains(x, n_total, out=gathered)
    # L2HMC_BENCH_NO_SMI=1: development switch to measure what the nvidia-smi polling itself costs (it does perturb the GPU)
    sampler = ClockSampler(local) if rank == 0 and os.environ.get("L2HMC_BENCH_NO_SMI") != "1" else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    dyn.timing_enable(True)
    launches0 = dyn.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # all ranks enter the timed region together (rank 0 alone starts the clock sampler above: without this barrier
    # the other ranks' timed all-gather would sit waiting for it)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record()
    px_sum = 0.0
    for _ in range(args.steps):
        o = step(x, ctr)
        x = o["x_next"]
        ctr += 1
    if world > 1:
        samples = all_gather_chains(x, n_total, out=gathered)  # the single NCCL all-gather of samples at the end
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = dyn.launch_count - launches0
    kern_ms, kern_cnt = dyn.timing_read()
    dyn.timing_enable(False)
    mean_px = float(o["px"].mean())
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        lt = torch.tensor([launches], device=dev, dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])
        dist.barrier()
    value = n_total * LF * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region --------
    hx = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
    hx.copy_(x)
    out = {"Lx": None, "Lv": None, "px": torch.empty((n,), dtype=torch.float32, pin_memory=True).numpy(),
           "x_next": torch.empty((n, D), dtype=torch.float32, pin_memory=True).numpy(),
           "accepted": torch.empty((n,), dtype=torch.uint8, pin_memory=True).numpy()}
    hxn = hx.numpy()
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
        dyn.transition_host(cur, counter=ctr, chain_offset=lo, out=out)  # synchronous: returns after D2H
        cur = out["x_next"]
        ctr += 1
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = n_total * LF * e2e_steps / e2e_s
    h2d = n * D * 4
    d2h = n * D * 4 + n * 4 + n

    if sampler:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    clocks = sampler.summary(t_wall0, t_wall1) if sampler else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    peaks, src = measured_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fma_peak_tflops = 148 * 128 * 2 * sm_max * 1e6 / 1e12  # fp32 FMA pipe at the max SM clock
    launch_flops = n * LF * FLOP_STEP
    achieved = launch_flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else None
    roofline = {"bound": "fma", "achieved": achieved, "peak": fma_peak_tflops, "unit": "TFLOP/s",
                "frac": (achieved / fma_peak_tflops) if achieved else None, "traffic": None,
                "peak_source": "148 SM x 128 fp32 lanes x 2 flop x sm_max_mhz from MEASURED_PEAKS.json (%s); "
                               "the path is FMA-pipe bound, not HBM or tensor (SURVEY.md section 8d)" % src,
                "kernel": dyn.kernel_name, "kernel_ms": kern_ms, "kernel_launches_timed": kern_cnt,
                "algorithmic_flops_per_launch": launch_flops,
                "algorithmic_hbm_bytes_per_launch": n * (3 * D * 4 + 4 + 1),
                "frac_of_measured_tf32_tensor": (achieved / (float(peaks["bf16_tflops_sustained"]) / 2)) if achieved else None}
    if clocks.get("sm_mhz"):
        roofline["frac_at_clock_under_load"] = achieved / (148 * 128 * 2 * clocks["sm_mhz"] * 1e6 / 1e12) if achieved else None
    # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this same workload
    # (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_tc_final_ncu.txt); below the algorithmic bytes because the
    # 126 MB L2 still holds part of the written lines when the launch ends
    if dyn.kernel_name.startswith("tc") and n == CHAINS_PER_GPU:
        roofline["traffic"] = 53.189120e6 + 51.559936e6
        roofline["traffic_source"] = "profiles/r01_tc_final_ncu.txt"
    if dyn.kernel_name.startswith("tc") and achieved:
        # the GEMMs run on the tensor pipe with an error-compensated split: three MMAs per fp32-accurate product, tf32
        # (K = 8 per MMA) or fp16 pairs (K = 16 per MMA, twice the rate) -- l2hmc_kernel_name says which the launch used.
        # Denominator: the measured sustained dense rate of that input type (bf16 figure for fp16, half of it for tf32).
        f16 = dyn.kernel_name == "tc_3xf16"
        t_peak = float(peaks["bf16_tflops_sustained"]) / (1.0 if f16 else 2.0)
        roofline.update({"bound": "tensor", "peak": t_peak, "frac": achieved / t_peak,
                         "frac_of_fma_roofline": achieved / fma_peak_tflops,
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
        r = cpu_reference_run(steps=5, warmup=1, sample_chains=args.cpu_chains)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload, chains_per_gpu=n, global_chains=n_total, parallelism="chains sharded x%d" % world,
                           l2="per-step working set (x, Lx, x_next, px) = %.0f MB > 126 MB L2; no flush" % (3 * n * D * 4 / 1e6),
                           kernel=dyn.kernel_name, mean_accept_prob=mean_px),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "l2hmc_transition_host (pinned host buffers)"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": {k: rep[k] for k in ("Lx_kernel", "Lv_kernel", "px_kernel", "px_mean_kernel", "px_mean_ref", "px_o32")} if rep else None}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
