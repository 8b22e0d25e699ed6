#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the in-tree libl2hmc.so (cuobjdump -sass): the evidence that the hot kernels are
tcgen05 / TMEM / bulk-TMA code for sm_100a (UTCHMMA, LDTM, STTM, UBLKCP) with packed fp32 (FFMA2 / FMUL2 / FADD2).
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "l2hmc_b200", "libl2hmc.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "USETMAXREG", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD",
       "MUFU", "F2FP", "HMMA", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMG", "REDG", "RED", "ATOM"]
cur, hist, order = None, {}, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
arch = re.findall(r"arch = (sm_\w+)", out)
print("# %s: %d kernels, arch %s" % (os.path.basename(lib), len(order), sorted(set(arch))))
print("# demangled name (c++filt) | total SASS instructions | counts of the opcodes that matter")
names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
for mangled, name in sorted(zip(order, names), key=lambda t: -sum(hist[t[0]].values())):
    h = hist[mangled]
    short = re.sub(r"\(.*\)$", "", name).replace("l2hmc::", "")
    print("%-92s %6d | %s" % (short[:92], sum(h.values()), " ".join("%s=%d" % (k, h[k]) for k in KEY if h[k])))
