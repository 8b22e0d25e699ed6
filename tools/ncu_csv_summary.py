#!/usr/bin/env python
"""Summarise the CSV pages of one Nsight Compute capture (`ncu -i x.ncu-rep --page raw --csv`, `--page source --csv`,
written on the GPU box by tools/run_profiles.sh) into a small text file for profiles/.

    python tools/ncu_csv_summary.py gpurun_out/r02_c2_ncu_raw.csv [gpurun_out/r02_c2_ncu_source.csv] > profiles/r02_c2_ncu.txt
"""
import collections
import csv
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.per_cycle_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
]


def raw_page(path):
    rows = list(csv.reader(open(path)))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    out = []
    for r in rows[head + 2:]:
        if len(r) < len(names):
            continue
        d = {n: (v, u) for n, u, v in zip(names, units, r)}
        out.append(d)
    return out


def source_page(path):
    rows = list(csv.reader(open(path)))
    head = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[head]
    ix = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[head + 1:] if len(r) == len(hdr)]
    return rows[0][1] if len(rows[0]) > 1 else "?", ix, data


def main():
    raw = raw_page(sys.argv[1])
    for k in raw:
        print("kernel:", k["Kernel Name"][0])
        for m in RAW:
            if m in k:
                print("  %-72s %s %s" % (m, k[m][0], k[m][1]))
        rd, wr = k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum")
        if rd and wr:
            sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = float(rd[0].replace(",", "")) * sc.get(rd[1], 1) + float(wr[0].replace(",", "")) * sc.get(wr[1], 1)
            print("  dram traffic of the launch (read + write): %.2f MB" % (tot / 1e6))
    if len(sys.argv) > 2:
        name, ix, data = source_page(sys.argv[2])

        def f(r, k):
            try:
                return float(r[ix[k]])
            except (ValueError, KeyError):
                return 0.0
        tot_s = sum(f(r, "# Samples") for r in data)
        tot_e = sum(f(r, "Instructions Executed") for r in data)
        print("\nsource page: %d SASS instructions, %.4g warp instructions executed, %d stall samples" % (len(data), tot_e, tot_s))
        stall = collections.Counter()
        for r in data:
            for k in ix:
                if k.startswith("stall_") and "Not Issued" not in k:
                    stall[k[6:]] += f(r, k)
        print("stall reasons (all warps):", "  ".join("%s=%.1f%%" % (k, 100 * v / max(tot_s, 1)) for k, v in stall.most_common(9)))
        op, ope = collections.Counter(), collections.Counter()
        for r in data:
            toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
            if not toks:
                continue
            name_ = toks[0].split(".")[0]
            op[name_] += f(r, "# Samples")
            ope[name_] += f(r, "Instructions Executed")
        print("%-10s %9s %6s %13s %6s" % ("opcode", "samples", "%", "executed", "%"))
        for k, v in ope.most_common(26):
            print("%-10s %9d %5.1f%% %13d %5.1f%%" % (k, op[k], 100 * op[k] / max(tot_s, 1), v, 100 * v / max(tot_e, 1)))


if __name__ == "__main__":
    main()
