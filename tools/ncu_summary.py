#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt [--launch 0]

Reads the report with `ncu -i ... --page raw --csv` (headline counters) and `--page source --csv`
(instruction mix by opcode, stall reasons), which works on the CPU-only dev box.
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.per_cycle_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "smsp__inst_executed.sum",
]


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
    lines = ["# ncu summary of %s (launch %d)" % (rep, launch), ""]
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2 + launch]
    col = {h: i for i, h in enumerate(hdr)}
    lines.append("kernel: %s" % vals[col.get("Kernel Name", 4)])
    for k in RAW:
        if k in col:
            lines.append("%-72s %-10s %s" % (k, units[col[k]], vals[col[k]]))
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
    h = next((i for i, r in enumerate(src) if "Source" in r and "# Samples" in r), None)
    if h is not None:
        sh = src[h]
        ix = {k: i for i, k in enumerate(sh)}
        data = [r for r in src[h + 1:] if len(r) == len(sh)]

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        ti = sum(f(r, "Instructions Executed") for r in data) or 1.0
        ts = sum(f(r, "# Samples") for r in data) or 1.0
        byop, bys = collections.Counter(), collections.Counter()
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            op = m.group(2).split(".")[0] if m else "?"
            byop[op] += f(r, "Instructions Executed")
            bys[op] += f(r, "# Samples")
        lines += ["", "instruction mix (warp instructions executed: %.4g)" % ti]
        for op, c in byop.most_common(14):
            lines.append("  %-12s %5.1f%% of instructions  %5.1f%% of stall samples" % (op, 100 * c / ti, 100 * bys[op] / ts))
        st = [k for k in sh if k.startswith("stall_") and "Not Issued" not in k]
        tot = {k: sum(f(r, k) for r in data) for k in st}
        s = sum(tot.values()) or 1.0
        lines += ["", "warp stall reasons (share of samples)"]
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]:
            lines.append("  %-24s %5.1f%%" % (k, 100 * v / s))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
