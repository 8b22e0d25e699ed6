"""Dynamics: the augmented-leapfrog operator (mirror of the reference's utils/dynamics.py).

Constructor, method names, argument order and return arities follow
/root/reference/utils/dynamics.py:34-309.  Underneath, every call goes to libl2hmc.so (C ABI in
include/l2hmc.h); PyTorch tensors are only storage (``data_ptr()``), streams come from torch.

Differences a reference user should know (all additive or forced by eager execution):
  * tensors are eager CUDA fp32 ``torch.Tensor``s instead of symbolic TF tensors;
  * ``temperature`` is a plain float attribute (the reference feeds a placeholder, :47);
  * randomness is Philox keyed by ``seed`` and a per-object call counter (the reference is unseeded);
    explicit momentum via ``init_v`` exactly as in the reference, explicit direction bits / accept
    uniforms through the keyword-only ``rng`` argument of ``propose``;
  * ``energy_function`` must come from ``l2hmc_b200.distributions`` (closed-form descriptor) or be a
    ``l2hmc_b200.vae.DecoderEnergy`` (the ``energy(z, aux)`` closure of mnist_vae.py:122-126); an
    arbitrary Python callable cannot be fused into the kernels and is rejected loudly;
  * ``aux`` (mnist_vae.py:196,204 pass the image batch) is a CUDA fp32 [N, aux_dim] tensor; it is required
    exactly when the energy or the nets consume it.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .layers import compile_softplus_mlp, compile_stq_net

TORCH_FLOAT = torch.float32
NP_FLOAT = np.float32

_KERNELS = {"auto": _lib.KERNEL_AUTO, "tile": _lib.KERNEL_TILE, "small": _lib.KERNEL_SMALL, "tc": _lib.KERNEL_TC,
            "layered": _lib.KERNEL_LAYERED, "layered_fma": _lib.KERNEL_LAYERED_FMA}


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def init_mask(x_dim: int, T: int, rng=None) -> np.ndarray:
    """_init_mask (utils/dynamics.py:84-93): per step, ones at the first int(x_dim / 2) entries of a
    random permutation.  rng=None uses numpy's global RNG like the reference."""
    perm = np.random.permutation if rng is None else rng.permutation
    mask_per_step = []
    for _ in range(T):
        ind = perm(np.arange(x_dim))[:int(x_dim / 2)]
        m = np.zeros((x_dim,))
        m[ind] = 1
        mask_per_step.append(m)
    return np.stack(mask_per_step).astype(NP_FLOAT)


class Dynamics(object):
    def __init__(self,
                 x_dim,
                 energy_function,
                 T=25,
                 eps=0.1,
                 hmc=False,
                 net_factory=None,
                 eps_trainable=True,
                 use_temperature=False,
                 *,
                 device=None,
                 seed=0,
                 kernel="auto",
                 mask_rng=None):
        self.x_dim = int(x_dim)
        self.use_temperature = use_temperature
        self.temperature = 1.0
        self.eps_trainable = eps_trainable
        # alpha = log(eps) and eps = exp(alpha) in fp32, as the reference's variable does (:50-58)
        self.alpha = np.log(np.float32(eps)).astype(NP_FLOAT)
        self._fn = energy_function
        self.T = int(T)
        self.hmc = bool(hmc)
        self.seed = int(seed)
        self._counter = 0
        self._kernel = _KERNELS[kernel] if isinstance(kernel, str) else int(kernel)
        self._ctx = None
        self._lib = None
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device_index = torch.device(device).index if not isinstance(device, int) else device
        if self.device_index is None:
            self.device_index = 0

        self._mask = init_mask(self.x_dim, self.T, mask_rng)

        if not hasattr(energy_function, "kind"):
            raise TypeError("energy_function must be a closed-form energy from l2hmc_b200.distributions "
                            "(got %r); arbitrary callables cannot run inside the fused kernel" % (energy_function,))
        if energy_function.dim != self.x_dim:
            raise ValueError("energy function is %d-d but x_dim=%d" % (energy_function.dim, self.x_dim))

        if hmc:
            # if HMC we just return all zeros (:73-76)
            self.XNet = lambda inp: [torch.zeros_like(inp[0]) for t in range(3)]
            self.VNet = lambda inp: [torch.zeros_like(inp[0]) for t in range(3)]
            self.width = 0
            self._net_params = None
        else:
            if net_factory is None:
                raise ValueError("net_factory is required unless hmc=True")
            self.XNet = net_factory(x_dim, scope='XNet', factor=2.0)
            self.VNet = net_factory(x_dim, scope='VNet', factor=1.0)
            self._net_params = [compile_stq_net(self.XNet, self.x_dim), compile_stq_net(self.VNet, self.x_dim)]
            self.width = int(self._net_params[0]["W4"].shape[0])
            if self._net_params[1]["W4"].shape[0] != self.width:
                raise ValueError("XNet and VNet must have the same width")
        self._compile_aux_encoder()

        if torch.cuda.is_available():
            self._ensure_ctx()

    # ---- library context --------------------------------------------------------------------------
    def _ensure_ctx(self):
        if self._ctx is not None:
            return
        if not torch.cuda.is_available():
            raise _lib.L2HMCLibraryError("no CUDA device: the L2HMC sampling path has no CPU fallback")
        lib = _lib.load()
        cfg = _lib.Config(self.x_dim, max(self.width, 1), self.T, 1 if self.hmc else 0, self.device_index,
                          self._kernel, float(self.eps))
        ctx = C.c_void_p()
        rc = lib.l2hmc_create(C.byref(cfg), C.byref(ctx))
        if rc != 0:
            raise _lib.L2HMCError(rc, (lib.l2hmc_last_error(None) or b"?").decode())
        self._lib, self._ctx = lib, ctx
        self._push_energy()
        self._push_mask()
        if not self.hmc:
            self._push_nets()

    def _compile_aux_encoder(self):
        """The 4th Zip entry of both nets: `lambda _: 0.` or ONE softplus MLP of aux shared by XNet and VNet
        (encoder_sampler, mnist_vae.py:134-149)."""
        self._aux_encoder = None
        if self.hmc or self._net_params is None:
            return
        ex, ev = (p.get("aux_encoder") for p in self._net_params)
        if ex is None and ev is None:
            return
        if ex is None or ev is None:
            raise ValueError("XNet and VNet must both take the aux encoding or neither")
        wx, Wx, bx = compile_softplus_mlp(ex, "aux encoder")
        if ev is not ex:
            wv, Wv, bv = compile_softplus_mlp(ev, "aux encoder")
            same = wx == wv and all(np.array_equal(a, b) for a, b in zip(Wx + bx, Wv + bv))
            if not same:
                raise ValueError("XNet and VNet must share one aux encoder (mnist_vae.py:134-149 builds a single "
                                 "encoder_sampler); two different encoders are not supported")
        if wx[-1] != self.width:
            raise ValueError("aux encoder ends in %d units but the nets are %d wide" % (wx[-1], self.width))
        self._aux_encoder = (wx, Wx, bx)

    @property
    def aux_dim(self):
        """Columns of the aux rows this Dynamics consumes (0: none)."""
        if getattr(self._fn, "kind", None) == _lib.ENERGY_DECODER:
            return int(self._fn.aux_dim)
        if self._aux_encoder is not None:
            return int(self._aux_encoder[0][0])
        return 0

    @staticmethod
    def _mlp_args(widths, Ws, bs):
        n = len(Ws)
        w = (C.c_int32 * (n + 1))(*[int(v) for v in widths])
        Wp = (C.POINTER(C.c_float) * n)(*[_fptr(a) for a in Ws])
        bp = (C.POINTER(C.c_float) * n)(*[_fptr(a) for a in bs])
        return n, w, Wp, bp

    def __del__(self):
        try:
            if self._ctx is not None and self._lib is not None:
                self._lib.l2hmc_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def _chk(self, rc):
        _lib.check(self._lib, self._ctx, rc)

    def _push_energy(self):
        e = self._fn
        if e.kind == _lib.ENERGY_DECODER:
            n, w, Wp, bp = self._mlp_args(e.widths, e.Ws, e.bs)
            self._chk(self._lib.l2hmc_set_energy_decoder(self._ctx, n, w, Wp, bp))
            return
        null = C.POINTER(C.c_float)()

        def parts(e):
            return (_fptr(e.mu) if e.mu is not None else null, _fptr(e.S) if e.S is not None else null,
                    _fptr(e.logc) if e.logc is not None else null, _fptr(e.scalars) if e.scalars is not None else null,
                    0 if e.scalars is None else int(e.scalars.size))
        if e.kind == _lib.ENERGY_MIXED:   # (1 - beta) U_a + beta U_b, utils/ais.py:44-45
            da, db = (_lib.EnergyDesc(p.kind, p.n_comp, *parts(p)) for p in (e.a, e.b))
            self._chk(self._lib.l2hmc_set_energy_mixed(self._ctx, C.byref(da), C.byref(db), float(e.beta)))
            return
        mu, S, lc, sc, ns = parts(e)
        self._chk(self._lib.l2hmc_set_energy(self._ctx, e.kind, e.n_comp, mu, S, lc, sc, ns))

    def set_mix_beta(self, beta):
        """Move the weight of a mixed (annealed) energy, utils/ais.py:44-45, without re-sending its parameters."""
        if getattr(self._fn, "kind", None) != _lib.ENERGY_MIXED:
            raise TypeError("set_mix_beta applies to distributions.MixedEnergy")
        self._fn.beta = float(beta)
        if self._ctx is not None:
            self._chk(self._lib.l2hmc_set_mix_beta(self._ctx, float(beta)))

    def _push_mask(self):
        m = np.ascontiguousarray(self._mask, dtype=NP_FLOAT)
        self._chk(self._lib.l2hmc_set_masks(self._ctx, _fptr(m)))

    def _push_nets(self):
        for net_id, p in ((_lib.XNET, self._net_params[0]), (_lib.VNET, self._net_params[1])):
            st = _lib.NetParams(**{k: _fptr(p[k]) for k in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4",
                                                             "Ws", "bs", "Wt", "bt", "Wq", "bq")},
                                scale_s=_fptr(p["ls"]), scale_q=_fptr(p["lq"]))
            self._chk(self._lib.l2hmc_set_net(self._ctx, net_id, C.byref(st)))
        if self._aux_encoder is not None:
            n, w, Wp, bp = self._mlp_args(*self._aux_encoder)
            self._chk(self._lib.l2hmc_set_aux_encoder(self._ctx, n, w, Wp, bp))

    def status_flags(self, clear=False):
        """Sticky status bits of the context (l2hmc_status_flags): read from pinned host memory, no device synchronisation
        (a launch still in flight may raise a bit later)."""
        self._ensure_ctx()
        f = C.c_uint32(0)
        self._chk(self._lib.l2hmc_status_flags(self._ctx, C.byref(f), int(bool(clear))))
        return int(f.value)

    def fp16_range_exceeded(self, synchronize=True):
        """True when a launch of THIS Dynamics' tensor-core kernel met an activation outside the fp16 range while using
        the fp16 operand split.  The affected chains of that launch carry non-finite proposals (accept probability 0:
        rejected, like any non-finite p in the reference, utils/dynamics.py:309); every later launch runs the tf32 split
        (fp32 range) on its own.  ``synchronize``: wait for launches in flight first, so that the answer covers them."""
        self._ensure_ctx()
        if synchronize:
            torch.cuda.synchronize(self.device_index)
        return bool(self.status_flags() & _lib.STATUS_F16_RANGE)

    def set_likelihood_scale(self, beta):
        """Decoder energy only: U = beta * sum BCE + 0.5 |z|^2, the annealed energy between the prior and the posterior
        (utils/ais.py:44-45 with init_energy = standard normal, eval_vae.py:52-62)."""
        if getattr(self._fn, "kind", None) != _lib.ENERGY_DECODER:
            raise TypeError("set_likelihood_scale applies to the decoder energy")
        self._ensure_ctx()
        self._chk(self._lib.l2hmc_set_likelihood_scale(self._ctx, float(beta)))

    def set_energy_function(self, energy_function):
        """Replace the target of an existing Dynamics (same kind of descriptor, same dimension): the annealed energy of
        utils/ais.py:44-58 changes at every step while the leapfrog operator stays."""
        if not hasattr(energy_function, "kind") or energy_function.dim != self.x_dim:
            raise TypeError("set_energy_function needs a closed-form %d-d energy" % self.x_dim)
        self._fn = energy_function
        if self._ctx is not None:
            self._push_energy()

    def refresh(self):
        """Re-read XNet/VNet weights from the layer objects (after loading a checkpoint into them)."""
        if not self.hmc:
            self._net_params = [compile_stq_net(self.XNet, self.x_dim), compile_stq_net(self.VNet, self.x_dim)]
            self._compile_aux_encoder()
            if self._ctx is not None:
                self._push_nets()

    # ---- public attributes the reference exposes ------------------------------------------------------
    @property
    def eps(self):
        return float(np.exp(self.alpha, dtype=NP_FLOAT))

    @eps.setter
    def eps(self, value):
        self.alpha = np.log(np.float32(value)).astype(NP_FLOAT)
        if self._ctx is not None:
            self._chk(self._lib.l2hmc_set_eps(self._ctx, float(self.eps)))

    def set_alpha(self, alpha):
        """Assign the stored variable alpha = log(eps) itself (utils/dynamics.py:50-58), as an optimiser does."""
        self.alpha = np.float32(alpha)
        if self._ctx is not None:
            self._chk(self._lib.l2hmc_set_eps(self._ctx, float(self.eps)))

    @property
    def mask(self):
        return self._mask

    @mask.setter
    def mask(self, value):
        # eval_sampler.py:156 assigns dynamics.mask after construction
        value = np.asarray(value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else value, dtype=NP_FLOAT)
        if value.shape != (self.T, self.x_dim):
            raise ValueError("mask must be [T, x_dim] = [%d, %d]" % (self.T, self.x_dim))
        self._mask = value
        if self._ctx is not None:
            self._push_mask()

    @property
    def kernel_name(self):
        self._ensure_ctx()
        return self._lib.l2hmc_kernel_name(self._ctx).decode()

    @property
    def launch_count(self):
        return 0 if self._ctx is None else int(self._lib.l2hmc_launch_count(self._ctx))

    # ---- helpers ------------------------------------------------------------------------------------
    def _prep(self, t, name, cols=None):
        if not isinstance(t, torch.Tensor):
            raise TypeError("%s must be a torch.Tensor on CUDA" % name)
        if not t.is_cuda:
            raise TypeError("%s must live on the GPU (got %s)" % (name, t.device))
        if t.device.index != self.device_index:
            raise ValueError("%s is on cuda:%d but this Dynamics is bound to cuda:%d" % (name, t.device.index, self.device_index))
        t = t.detach()
        if t.dtype != TORCH_FLOAT:
            t = t.to(TORCH_FLOAT)
        t = t.contiguous()
        if cols is not None and (t.dim() != 2 or t.shape[1] != cols):
            raise ValueError("%s must have shape [N, %d], got %s" % (name, cols, tuple(t.shape)))
        return t

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device_index).cuda_stream)

    def _sync_temperature(self):
        t = float(self.temperature) if self.use_temperature else 1.0
        self._chk(self._lib.l2hmc_set_temperature(self._ctx, t))

    def _aux(self, aux, n):
        """Validated aux rows for n chains, or None when this Dynamics consumes none (then aux is ignored exactly
        as the reference's `aux=aux` pass-through to nets that drop it, SCGExperiment.ipynb:58)."""
        ad = self.aux_dim
        if ad == 0:
            return None
        if aux is None:
            raise ValueError("this target / these nets are conditioned on aux: pass aux=[N, %d]" % ad)
        aux = self._prep(aux, "aux", ad)
        if aux.shape[0] != n:
            raise ValueError("aux has %d rows for %d chains" % (aux.shape[0], n))
        return aux

    def _bind_aux(self, aux, n):
        aux = self._aux(aux, n)
        self._chk(self._lib.l2hmc_bind_aux(self._ctx, n, aux.data_ptr() if aux is not None else None))
        return aux

    def next_counter(self, n=1):
        c = self._counter
        self._counter += int(n)
        return c

    def _transition(self, x, *, v=None, dir_mode=_lib.DIR_FORWARD, direction=None, u=None, log_jac=False,
                    do_mh=False, n_transitions=1, want_v=True, counter=None, chain_offset=0, seed=None, out=None,
                    aux=None, stats=None, trace=None, chain=False, v0=None):
        """One l2hmc_transition call. Returns dict(Lx, Lv, px, x_next, accepted).  `out` may hold preallocated
        tensors of the right shapes under the same keys (steady-state loops then allocate nothing).
        stats: optional CUDA float64 [2] accumulator (+= sum of px, += number accepted, reduced in the kernel);
        trace: optional CUDA fp32 [n_transitions, N, x_dim] receiving the Metropolis output of every fused transition."""
        self._ensure_ctx()
        self._sync_temperature()
        x = self._prep(x, "x", self.x_dim)
        n = x.shape[0]
        dev = x.device
        a = _lib.TransitionArgs()
        a.n, a.chain_offset = n, int(chain_offset)
        a.x = x.data_ptr()
        keep = [x]
        aux = self._aux(aux, n)
        if aux is not None:
            keep.append(aux)
            a.aux = aux.data_ptr()
        if v is not None:
            v = self._prep(v, "init_v")
            if v.numel() != n_transitions * n * self.x_dim:
                raise ValueError("init_v must have %d x [N, x_dim] elements" % n_transitions)
            keep.append(v)
            a.v = v.data_ptr()
        if direction is not None:
            direction = direction.detach().to(device=dev, dtype=torch.uint8).contiguous()
            if direction.numel() != n_transitions * n:
                raise ValueError("direction must have N entries")
            keep.append(direction)
            a.dir = direction.data_ptr()
            dir_mode = _lib.DIR_PER_CHAIN
        if u is not None:
            u = self._prep(u, "u")
            if u.numel() != (1 if chain else n_transitions) * n:
                raise ValueError("u must have N entries")
            keep.append(u)
            a.u = u.data_ptr()
        a.dir_mode, a.log_jac, a.do_mh, a.n_transitions = int(dir_mode), int(bool(log_jac)), int(bool(do_mh)), int(n_transitions)
        a.seed = self.seed if seed is None else int(seed)
        a.counter = self.next_counter(n_transitions) if counter is None else int(counter)
        out = out if out is not None else {
            "Lx": torch.empty((n, self.x_dim), dtype=TORCH_FLOAT, device=dev),
            "Lv": torch.empty((n, self.x_dim), dtype=TORCH_FLOAT, device=dev) if want_v else None,
            "px": torch.empty((n,), dtype=TORCH_FLOAT, device=dev),
            "x_next": torch.empty((n, self.x_dim), dtype=TORCH_FLOAT, device=dev) if do_mh else None,
            "accepted": torch.empty((n,), dtype=torch.uint8, device=dev) if do_mh else None,
        }
        a.x_out = out["Lx"].data_ptr()
        a.v_out = out["Lv"].data_ptr() if want_v else None
        a.px_out = out["px"].data_ptr()
        a.x_next = out["x_next"].data_ptr() if do_mh else None
        a.accepted = out["accepted"].data_ptr() if do_mh else None
        if stats is not None:
            if not (stats.is_cuda and stats.dtype == torch.float64 and stats.numel() == 2 and stats.is_contiguous()):
                raise TypeError("stats must be a contiguous CUDA float64 tensor of 2 elements")
            a.stats = stats.data_ptr()
        if chain:
            a.chain = 1
            if v0 is not None:
                v0 = self._prep(v0, "init_v", self.x_dim)
                if v0.shape[0] != n:
                    raise ValueError("init_v must be [N, x_dim]")
                keep.append(v0)
                a.v0 = v0.data_ptr()
        if trace is not None:
            if not (trace.is_cuda and trace.dtype == TORCH_FLOAT and trace.is_contiguous() and
                    tuple(trace.shape) == (int(n_transitions), n, self.x_dim)):
                raise TypeError("trace must be a contiguous CUDA fp32 tensor [n_transitions, N, x_dim]")
            a.trace = trace.data_ptr()
        a.stream = self._stream()
        self._chk(self._lib.l2hmc_transition(self._ctx, C.byref(a)))
        return out

    def transition_host(self, x, *, v=None, direction=None, u=None, dir_mode=_lib.DIR_RANDOM, log_jac=False,
                        do_mh=True, n_transitions=1, counter=None, chain_offset=0, seed=None, out=None, aux=None, stats=None):
        """The same transition through l2hmc_transition_host: numpy in, numpy out, H2D/D2H inside.
        stats: optional numpy float64 [2] accumulator (+= sum of px, += number accepted)."""
        self._ensure_ctx()
        self._sync_temperature()
        x = np.ascontiguousarray(x, dtype=NP_FLOAT)
        n = x.shape[0]
        a = _lib.TransitionArgs()
        a.n, a.chain_offset = n, int(chain_offset)
        a.x = x.ctypes.data
        keep = [x]
        if self.aux_dim:
            if aux is None:
                raise ValueError("this target / these nets are conditioned on aux: pass aux=[N, %d]" % self.aux_dim)
            aux = np.ascontiguousarray(aux, dtype=NP_FLOAT)
            if aux.shape != (n, self.aux_dim):
                raise ValueError("aux must be [%d, %d]" % (n, self.aux_dim))
            keep.append(aux)
            a.aux = aux.ctypes.data
        if v is not None:
            v = np.ascontiguousarray(v, dtype=NP_FLOAT)
            keep.append(v)
            a.v = v.ctypes.data
        if direction is not None:
            direction = np.ascontiguousarray(direction, dtype=np.uint8)
            keep.append(direction)
            a.dir = direction.ctypes.data
            dir_mode = _lib.DIR_PER_CHAIN
        if u is not None:
            u = np.ascontiguousarray(u, dtype=NP_FLOAT)
            keep.append(u)
            a.u = u.ctypes.data
        a.dir_mode, a.log_jac, a.do_mh, a.n_transitions = int(dir_mode), int(bool(log_jac)), int(bool(do_mh)), int(n_transitions)
        a.seed = self.seed if seed is None else int(seed)
        a.counter = self.next_counter(n_transitions) if counter is None else int(counter)
        if out is None:
            out = {"Lx": np.empty((n, self.x_dim), NP_FLOAT), "Lv": np.empty((n, self.x_dim), NP_FLOAT),
                   "px": np.empty((n,), NP_FLOAT),
                   "x_next": np.empty((n, self.x_dim), NP_FLOAT) if do_mh else None,
                   "accepted": np.empty((n,), np.uint8) if do_mh else None}
        ptr = lambda k: out[k].ctypes.data if out.get(k) is not None else None  # noqa: E731
        a.x_out = ptr("Lx")
        a.v_out = ptr("Lv")
        a.px_out = ptr("px")
        a.x_next = ptr("x_next") if do_mh else None
        a.accepted = out["accepted"].ctypes.data if (do_mh and out.get("accepted") is not None) else None
        if stats is not None:
            if not (isinstance(stats, np.ndarray) and stats.dtype == np.float64 and stats.size == 2 and stats.flags.c_contiguous):
                raise TypeError("stats must be a contiguous numpy float64 array of 2 elements")
            a.stats = stats.ctypes.data
        self._chk(self._lib.l2hmc_transition_host(self._ctx, C.byref(a)))
        return out

    # ---- reference methods ----------------------------------------------------------------------------
    def _get_mask(self, step):
        m = torch.as_tensor(self._mask[int(step)])
        return m, 1. - m

    def _format_time(self, t, tile=1):
        arg = np.float32(2 * np.pi) * np.float32(t) / np.float32(self.T)
        trig_t = torch.tensor([np.cos(arg, dtype=NP_FLOAT), np.sin(arg, dtype=NP_FLOAT)], dtype=TORCH_FLOAT)
        return trig_t[None, :].repeat(tile, 1)

    def kinetic(self, v):
        self._ensure_ctx()
        v = self._prep(v, "v", self.x_dim)
        out = torch.empty((v.shape[0],), dtype=TORCH_FLOAT, device=v.device)
        self._chk(self._lib.l2hmc_kinetic(self._ctx, v.shape[0], v.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def energy(self, x, aux=None):
        self._ensure_ctx()
        self._sync_temperature()
        x = self._prep(x, "x", self.x_dim)
        aux = self._bind_aux(aux, x.shape[0])
        out = torch.empty((x.shape[0],), dtype=TORCH_FLOAT, device=x.device)
        self._chk(self._lib.l2hmc_energy(self._ctx, x.shape[0], x.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def hamiltonian(self, x, v, aux=None):
        self._ensure_ctx()
        self._sync_temperature()
        x = self._prep(x, "x", self.x_dim)
        aux = self._bind_aux(aux, x.shape[0])
        v = self._prep(v, "v", self.x_dim)
        out = torch.empty((x.shape[0],), dtype=TORCH_FLOAT, device=x.device)
        self._chk(self._lib.l2hmc_hamiltonian(self._ctx, x.shape[0], x.data_ptr(), v.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def grad_energy(self, x, aux=None):
        self._ensure_ctx()
        self._sync_temperature()
        x = self._prep(x, "x", self.x_dim)
        aux = self._bind_aux(aux, x.shape[0])
        out = torch.empty_like(x)
        self._chk(self._lib.l2hmc_grad_energy(self._ctx, x.shape[0], x.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def net_apply(self, which, a, b, step, aux=None):
        """[S, T, Q] = {X,V}Net([a, b, _format_time(step), aux]) on the GPU (diagnostic entry point)."""
        self._ensure_ctx()
        a = self._prep(a, "a", self.x_dim)
        b = self._prep(b, "b", self.x_dim)
        aux = self._bind_aux(aux, a.shape[0]) if self._aux_encoder is not None else None
        S, T, Q = (torch.empty_like(a) for _ in range(3))
        net_id = _lib.XNET if which in ("XNet", "x", 0) else _lib.VNET
        self._chk(self._lib.l2hmc_net_apply(self._ctx, net_id, a.shape[0], a.data_ptr(), b.data_ptr(), float(step),
                                            S.data_ptr(), T.data_ptr(), Q.data_ptr(), self._stream()))
        return [S, T, Q]

    def forward(self, x, init_v=None, aux=None, log_path=False, log_jac=False):
        o = self._transition(x, v=init_v, dir_mode=_lib.DIR_FORWARD, log_jac=log_jac, aux=aux)
        return o["Lx"], o["Lv"], o["px"]

    def backward(self, x, init_v=None, aux=None, log_jac=False):
        o = self._transition(x, v=init_v, dir_mode=_lib.DIR_BACKWARD, log_jac=log_jac, aux=aux)
        return o["Lx"], o["Lv"], o["px"]

    def p_accept(self, x0, v0, x1, v1, log_jac, aux=None):
        self._ensure_ctx()
        self._sync_temperature()
        x0, v0, x1, v1 = (self._prep(t, n, self.x_dim) for t, n in ((x0, "x0"), (v0, "v0"), (x1, "x1"), (v1, "v1")))
        aux = self._bind_aux(aux, x0.shape[0])
        lj = self._prep(log_jac, "log_jac")
        out = torch.empty((x0.shape[0],), dtype=TORCH_FLOAT, device=x0.device)
        self._chk(self._lib.l2hmc_p_accept(self._ctx, x0.shape[0], x0.data_ptr(), v0.data_ptr(), x1.data_ptr(),
                                           v1.data_ptr(), lj.data_ptr(), out.data_ptr(), self._stream()))
        return out

    # ---- measurement hooks ----------------------------------------------------------------------------
    def timing_enable(self, on=True):
        self._ensure_ctx()
        self._chk(self._lib.l2hmc_timing_enable(self._ctx, int(bool(on))))

    def timing_read(self):
        ms, cnt = C.c_double(), C.c_int64()
        self._chk(self._lib.l2hmc_timing_read(self._ctx, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value
