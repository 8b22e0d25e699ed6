#!/usr/bin/env python
"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`) by kernel name and grid.
usage: launch_summary.py x.csv [top]"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault((row["Kernel Name"][:70], row["Grid Size"]), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{k[0]:72s} {k[1]:14s} n={a[0]:4d} {a[1]:9.3f} ms {100 * a[1] / tot:5.1f}%  avg {a[1] / a[0] * 1e3:8.1f} us")
print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
