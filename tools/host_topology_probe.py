#!/usr/bin/env python
"""Host-side topology of a GPU box and what it means for l2hmc_transition_host: NUMA nodes, the cpuset of this
container, each GPU's NUMA node, and H2D / D2H bandwidth of a 52 MB pinned buffer whose pages are bound (mbind) to each
NUMA node in turn.  Writes plain text to stdout (redirect into gpurun_out/)."""
import ctypes
import glob
import mmap
import os
import subprocess
import sys
import time

import numpy as np
import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %s" % e


def read(path):
    try:
        return open(path).read().strip()
    except Exception as e:  # noqa: BLE001
        return "n/a (%s)" % type(e).__name__


print("cpu_count", os.cpu_count(), "affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpuset.cpus.effective", "/sys/fs/cgroup/cpuset.mems.effective", "/sys/devices/system/node/online"):
    print(f, read(f))
for nd in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    print(nd, "cpus", read(nd + "/cpulist"), "|", read(nd + "/meminfo").split("\n")[0])
print(sh("lscpu | egrep 'Model name|Socket|NUMA|Thread|Core'"))
print(sh("nvidia-smi topo -m"))
n_gpu = torch.cuda.device_count()
for i in range(n_gpu):
    bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    q = sh("nvidia-smi -i %d --query-gpu=pci.bus_id --format=csv,noheader" % i).lower()
    q = q[4:] if q.startswith("0000") and len(q) > 12 else q
    print("gpu", i, q, "numa_node", read("/sys/bus/pci/devices/%s/numa_node" % q), "local_cpulist", read("/sys/bus/pci/devices/%s/local_cpulist" % q))

libc = ctypes.CDLL(None, use_errno=True)
SYS_mbind, MPOL_BIND, MPOL_MF_MOVE = 237, 2, 2
cudart = torch.cuda.cudart()
nodes = [int(os.path.basename(p)[4:]) for p in glob.glob("/sys/devices/system/node/node[0-9]*")]
size = 52 << 20
dev = torch.device("cuda", 0)
d = torch.empty(size, dtype=torch.uint8, device=dev)
for node in sorted(nodes) + [None]:
    mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    tag = "default policy"
    if node is not None:
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(SYS_mbind, ctypes.c_void_p(addr), ctypes.c_ulong(size), MPOL_BIND, ctypes.byref(mask), ctypes.c_ulong(64), 0)
        tag = "mbind node %d rc=%d errno=%d" % (node, rc, ctypes.get_errno() if rc else 0)
    a = np.frombuffer(mm, dtype=np.uint8)
    a[:] = 1  # first touch
    rc = int(cudart.cudaHostRegister(addr, size, 0))
    t = torch.from_numpy(a)
    res = []
    for direction in ("h2d", "d2h"):
        for _ in range(3):
            (d.copy_(t, non_blocking=True) if direction == "h2d" else t.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            (d.copy_(t, non_blocking=True) if direction == "h2d" else t.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        res.append("%s %.1f GB/s" % (direction, reps * size / (time.perf_counter() - t0) / 1e9))
    print("pinned 52 MB,", tag, "register rc=%d:" % rc, ", ".join(res), flush=True)
    cudart.cudaHostUnregister(addr)
    del t, a
    try:
        mm.close()
    except BufferError:
        pass
# torch's own pinned allocator for comparison
t = torch.empty(size, dtype=torch.uint8, pin_memory=True)
for direction in ("h2d", "d2h"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        (d.copy_(t, non_blocking=True) if direction == "h2d" else t.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    print("torch pin_memory", direction, "%.1f GB/s" % (20 * size / (time.perf_counter() - t0) / 1e9))
