"""ctypes binding of the plain-C oracle (oracle/l2hmc_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libl2hmc_oracle.so")
NET_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "Ws", "bs", "Wt", "bt", "Wq", "bq", "ls", "lq")
_fp = C.POINTER(C.c_float)


class OracleProblem(C.Structure):
    _fields_ = [("D", C.c_int), ("H", C.c_int), ("T", C.c_int), ("hmc", C.c_int), ("eps", C.c_float),
                ("mask", _fp), ("xnet", _fp * 16), ("vnet", _fp * 16),
                ("energy_kind", C.c_int), ("ncomp", C.c_int), ("mu", _fp), ("S", _fp), ("logc", _fp),
                ("s0", C.c_float), ("s1", C.c_float), ("temperature", C.c_float)]


_lib = None


def load():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, f) for f in ("l2hmc_oracle.c", "l2hmc_oracle_impl.h")]
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(LIB)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class COracle:
    """Holds the arrays of one problem (tests/util.py: Problem) alive and calls the C functions."""

    def __init__(self, P, temperature=1.0):
        self.lib = load()
        self.P = P
        self.keep = []
        op = OracleProblem()
        op.D, op.H, op.T, op.hmc = P.D, (P.H if not P.hmc else 1), P.T, int(P.hmc)
        op.eps = float(np.exp(np.log(np.float32(P.eps))))
        op.mask = self._ptr(P.mask)
        if not P.hmc:
            for i, k in enumerate(NET_KEYS):
                op.xnet[i] = self._ptr(P.xnet[k])
                op.vnet[i] = self._ptr(P.vnet[k])
        e = P.dist.get_energy_function()
        op.energy_kind, op.ncomp = e.kind, e.n_comp
        if e.mu is not None:
            op.mu, op.S = self._ptr(e.mu), self._ptr(e.S)
        if e.logc is not None:
            op.logc = self._ptr(e.logc)
        if e.scalars is not None:
            op.s0, op.s1 = float(e.scalars[0]), float(e.scalars[1])
        op.temperature = float(temperature)
        self.op = op

    def _ptr(self, a):
        a = _f32(a)
        self.keep.append(a)
        return a.ctypes.data_as(_fp)

    def propose(self, d, dtype=np.float64, log_jac=False):
        suf = "_f64" if dtype == np.float64 else "_f32"
        fn = getattr(self.lib, "oracle_propose" + suf)
        n, D = d["x"].shape
        x, vf, vb, u = (np.ascontiguousarray(d[k], dtype=dtype) for k in ("x", "v_f", "v_b", "u"))
        dirb = np.ascontiguousarray(d["dir"], dtype=np.uint8)
        Lx, Lv, xn = (np.empty((n, D), dtype) for _ in range(3))
        px = np.empty((n,), dtype)
        scratch = np.empty((2 * (2 * n * D + n),), dtype)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        fn.restype = None
        fn(C.byref(self.op), C.c_int(n), p(x), p(vf), p(vb), p(dirb), p(u), C.c_int(int(log_jac)), p(Lx), p(Lv), p(px),
           p(xn), p(scratch))
        return {"Lx": Lx, "Lv": Lv, "px": px, "x_next": xn}
