"""ctypes binding of libl2hmc.so (C ABI in include/l2hmc.h) and the in-tree nvcc build.

The library is the product: if it is missing or does not load, importing users get a loud
``L2HMCLibraryError`` -- there is no eager / CPU fallback for the sampling path.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from typing import Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# L2HMC_LIB: development override (kernel variants built side by side); the product library is in-tree
LIB_PATH = os.environ.get("L2HMC_LIB") or os.path.join(PKG_DIR, "libl2hmc.so")
SOURCES = ["l2hmc_api.cu"]
HEADERS = ["common.cuh", "kernel_tile.cuh", "kernel_tc.cuh", "kernel_small.cuh", "layered.cuh", "layered_host.cuh", "tc_gemm.cuh",
           os.path.join("..", "..", "include", "l2hmc.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class L2HMCLibraryError(RuntimeError):
    pass


class L2HMCError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libl2hmc status %d: %s" % (code, msg))
        self.code = code


def _src_hash() -> str:
    """Content hash of everything the library is compiled from (mtimes do not survive a copy of the tree)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    for p in [os.path.join(CSRC, f) for f in names] + [os.path.join(REPO_DIR, "include", "l2hmc.h")]:
        with open(p, "rb") as fh:
            h.update(os.path.basename(p).encode() + b"\0" + fh.read())  # names, not paths: the tree is copied to the GPU box
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(LIB_PATH + ".srchash"):
        return True
    with open(LIB_PATH + ".srchash") as fh:
        return fh.read().strip() != _src_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libl2hmc.so in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise L2HMCLibraryError("nvcc not found; cannot build %s" % LIB_PATH)
    # L2HMC_NVCC_EXTRA: development builds of kernel variants (e.g. -DL2HMC_TC_PHASE_ACCOUNTING); not part of the hash
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("L2HMC_NVCC_EXTRA", "").split() + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise L2HMCLibraryError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    if verbose:
        print(r.stderr)
    with open(LIB_PATH + ".srchash", "w") as fh:
        fh.write(_src_hash())
    return LIB_PATH


# ---- ctypes mirrors of include/l2hmc.h ------------------------------------------------------------
class Config(C.Structure):
    _fields_ = [("x_dim", C.c_int32), ("width", C.c_int32), ("T", C.c_int32), ("hmc", C.c_int32),
                ("device", C.c_int32), ("kernel", C.c_int32), ("eps", C.c_float)]


_fp = C.POINTER(C.c_float)


class NetParams(C.Structure):
    _fields_ = [(k, _fp) for k in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "Ws", "bs", "Wt", "bt",
                                   "Wq", "bq", "scale_s", "scale_q")]


class NetGrads(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "Ws", "bs", "Wt", "bt",
                                          "Wq", "bq", "scale_s", "scale_q")]


class LossGradArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("x", C.c_void_p), ("v", C.c_void_p), ("dir", C.c_void_p), ("scale", C.c_float),
                ("inv_count", C.c_float), ("loss", C.c_void_p), ("d_eps", C.c_void_p), ("grad_xnet", NetGrads),
                ("grad_vnet", NetGrads), ("x_out", C.c_void_p), ("px_out", C.c_void_p), ("stream", C.c_void_p),
                ("loss_kind", C.c_int32)]


class TransitionArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("chain_offset", C.c_int64),
                ("x", C.c_void_p), ("v", C.c_void_p), ("dir", C.c_void_p), ("u", C.c_void_p),
                ("dir_mode", C.c_int32), ("log_jac", C.c_int32), ("do_mh", C.c_int32), ("n_transitions", C.c_int32),
                ("seed", C.c_uint64), ("counter", C.c_uint64),
                ("x_out", C.c_void_p), ("v_out", C.c_void_p), ("px_out", C.c_void_p), ("x_next", C.c_void_p),
                ("accepted", C.c_void_p), ("stream", C.c_void_p), ("aux", C.c_void_p),
                ("stats", C.c_void_p), ("trace", C.c_void_p), ("chain", C.c_int32), ("v0", C.c_void_p)]


ENERGY_GAUSSIAN, ENERGY_GMM, ENERGY_ROUGHWELL, ENERGY_FUNNEL, ENERGY_DECODER, ENERGY_MIXED = 0, 1, 2, 3, 4, 5


class EnergyDesc(C.Structure):   # l2hmc_energy_desc
    _fields_ = [("kind", C.c_int32), ("n_comp", C.c_int32), ("mu", C.POINTER(C.c_float)), ("S", C.POINTER(C.c_float)),
                ("logc", C.POINTER(C.c_float)), ("scalars", C.POINTER(C.c_float)), ("n_scalars", C.c_int32)]
XNET, VNET = 0, 1
DIR_FORWARD, DIR_BACKWARD, DIR_PER_CHAIN, DIR_RANDOM = 0, 1, 2, 3
KERNEL_AUTO, KERNEL_TILE, KERNEL_SMALL, KERNEL_TC, KERNEL_LAYERED, KERNEL_LAYERED_FMA = 0, 1, 2, 3, 4, 5

# every symbol include/l2hmc.h declares: (name, restype, argtypes)
_vp, _i64, _u64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_float
EXPORTS = [
    ("l2hmc_create", C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    ("l2hmc_destroy", None, [_vp]),
    ("l2hmc_last_error", C.c_char_p, [_vp]),
    ("l2hmc_version", C.c_char_p, []),
    ("l2hmc_set_net", C.c_int, [_vp, C.c_int, C.POINTER(NetParams)]),
    ("l2hmc_set_masks", C.c_int, [_vp, _fp]),
    ("l2hmc_set_eps", C.c_int, [_vp, _f32]),
    ("l2hmc_set_temperature", C.c_int, [_vp, _f32]),
    ("l2hmc_set_likelihood_scale", C.c_int, [_vp, _f32]),
    ("l2hmc_set_energy", C.c_int, [_vp, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int]),
    ("l2hmc_set_energy_mixed", C.c_int, [_vp, C.POINTER(EnergyDesc), C.POINTER(EnergyDesc), C.c_float]),
    ("l2hmc_set_mix_beta", C.c_int, [_vp, C.c_float]),
    ("l2hmc_set_energy_decoder", C.c_int, [_vp, C.c_int, C.POINTER(_i32), C.POINTER(_fp), C.POINTER(_fp)]),
    ("l2hmc_set_aux_encoder", C.c_int, [_vp, C.c_int, C.POINTER(_i32), C.POINTER(_fp), C.POINTER(_fp)]),
    ("l2hmc_bind_aux", C.c_int, [_vp, _i64, _vp]),
    ("l2hmc_transition", C.c_int, [_vp, C.POINTER(TransitionArgs)]),
    ("l2hmc_transition_host", C.c_int, [_vp, C.POINTER(TransitionArgs)]),
    ("l2hmc_energy", C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    ("l2hmc_grad_energy", C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    ("l2hmc_kinetic", C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    ("l2hmc_hamiltonian", C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    ("l2hmc_p_accept", C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("l2hmc_net_apply", C.c_int, [_vp, C.c_int, _i64, _vp, _vp, _f32, _vp, _vp, _vp, _vp]),
    ("l2hmc_accept", C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    ("l2hmc_philox_fill", C.c_int, [_vp, _i64, _i64, _u64, _u64, _vp, _vp, _vp, _vp]),
    ("l2hmc_acl_spectrum", C.c_int, [_vp, _i64, _i64, _vp, C.c_double, _i64, _vp, _vp]),
    ("l2hmc_loss_grad", C.c_int, [_vp, C.POINTER(LossGradArgs)]),
    ("l2hmc_kernel_name", C.c_char_p, [_vp]),
    ("l2hmc_launch_count", _i64, [_vp]),
    ("l2hmc_timing_enable", C.c_int, [_vp, C.c_int]),
    ("l2hmc_timing_read", C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    ("l2hmc_debug_counters", C.c_int, [_vp, C.POINTER(_i64), C.c_int]),
    ("l2hmc_status_flags", C.c_int, [_vp, C.POINTER(C.c_uint32), C.c_int]),
]
STATUS_F16_RANGE = 1

_lib: Optional[C.CDLL] = None


def load(rebuild_if_stale: bool = True) -> C.CDLL:
    """dlopen libl2hmc.so and bind every export; raises L2HMCLibraryError if that is not possible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or (rebuild_if_stale and _stale() and shutil.which("nvcc")):
        build()
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise L2HMCLibraryError("cannot load %s: %s" % (LIB_PATH, e)) from e
    for name, res, args in EXPORTS:
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise L2HMCLibraryError("%s does not export %s" % (LIB_PATH, name)) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(lib, ctx, rc: int) -> None:
    if rc != 0:
        msg = lib.l2hmc_last_error(ctx)
        raise L2HMCError(rc, msg.decode() if msg else "?")
