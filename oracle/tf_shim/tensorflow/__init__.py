"""Eager, torch-CPU stand-in for the TensorFlow-1 symbols brain-research/l2hmc's hot path uses.
TEST INFRASTRUCTURE ONLY (oracle/): it exists so that the UNMODIFIED reference files
(/root/reference/utils/{dynamics,sampler,layers,distributions,losses,ais,func_utils}.py) can be
imported and executed in this container, where real TensorFlow 1.x is absent, to generate golden
fixtures (tests/golden/make_ref_golden.py) that pin the oracle and the CUDA kernels to the
reference's own code.  Nothing under l2hmc_b200/ imports it.

What it is: every ``tf.<op>`` below evaluates immediately on torch CPU tensors with TF-1 semantics
(Python scalars / numpy operands take the tensor's dtype, ``tf.where`` with a rank-1 condition selects
rows, ``tf.gradients`` sums the outputs, ``tf.while_loop`` iterates in Python).  Differences from a TF1
session, all deliberate:
  * eager: ``placeholder`` values are fed before use (``shim.feed``); graph construction == evaluation;
  * randomness is injected: ``random_normal`` / ``random_uniform`` pop arrays from FIFO queues the
    caller fills (``shim.feed_random``), in the order the reference code draws them;
  * variables come from a preloaded store (``shim.preload``: the moral equivalent of restoring a
    checkpoint) or from their initializer (seeded numpy);
  * ``tf.float32`` maps to the shim's *real* dtype: torch.float32 by default, torch.float64 after
    ``shim.set_real('float64')`` -- the same reference code then yields its own fp64 ground truth
    (fp32-rounded parameters such as ``i_sigma.astype('float32')`` stay fp32-rounded);
  * fp32 matmul summation order is torch's, not Eigen's (both are IEEE fp32; the fixtures the tests
    compare against are the fp64 runs, the fp32 runs measure the reference's own rounding noise).
"""
from __future__ import annotations

import builtins
import collections
import contextlib
import math as _math
import types as _types

import numpy as _np
import torch as _torch

__version__ = "1.x-shim (torch %s)" % _torch.__version__


# ----------------------------------------------------------------------------------------------------
# dtypes
# ----------------------------------------------------------------------------------------------------
class DType(object):
    def __init__(self, name):
        self.name = name

    @property
    def torch(self):
        if self.name == "float32":
            return shim.real  # the one switch: what the reference calls float32
        return {"float64": _torch.float64, "int32": _torch.int32, "int64": _torch.int64,
                "bool": _torch.bool}[self.name]

    def __repr__(self):
        return "tf." + self.name


float32, float64, int32, int64 = DType("float32"), DType("float64"), DType("int32"), DType("int64")
bool = DType("bool")  # noqa: A001  (tf.bool)


def _tdtype(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.torch
    if isinstance(dtype, _torch.dtype):
        return dtype
    return DType(_np.dtype(dtype).name).torch


# ----------------------------------------------------------------------------------------------------
# shim state: real dtype, random queues, variable store
# ----------------------------------------------------------------------------------------------------
class _Shim(object):
    def __init__(self):
        self.real = _torch.float32
        self.normal = collections.deque()
        self.uniform = collections.deque()
        self.variables = collections.OrderedDict()
        self.preloaded = {}
        self.scope = []
        self.rng = _np.random.RandomState(0)
        self.random_log = []

    # -- configuration ------------------------------------------------------------------------------
    def set_real(self, name):
        self.real = {"float32": _torch.float32, "float64": _torch.float64}[str(name)]

    def reset(self, seed=0):
        """Forget variables, scopes and pending random draws (a fresh TF1 graph + session)."""
        self.normal.clear()
        self.uniform.clear()
        self.variables.clear()
        self.preloaded = {}
        self.scope = []
        self.rng = _np.random.RandomState(seed)
        self.random_log = []

    def preload(self, values):
        """{'XNet/embed_1/W': array, ..., 'alpha': scalar}: values get_variable hands out instead of
        running the initializer (checkpoint-restore semantics)."""
        self.preloaded.update(values)

    def feed_random(self, normal=(), uniform=()):
        """Arrays returned, in order, by the next tf.random_normal / tf.random_uniform calls."""
        self.normal.extend(normal)
        self.uniform.extend(uniform)

    def feed(self, placeholder_tensor, value):
        placeholder_tensor.t = _cv(value, dtype=placeholder_tensor._ph_dtype).t

    def input(self, value, dtype=None):
        """A fed placeholder: a leaf tf.gradients can differentiate with respect to."""
        t = _cv(_np.asarray(value), dtype=_tdtype(dtype) or self.real).t.clone()
        t.requires_grad_(True)
        return Tensor(t)

    def pending_random(self):
        return len(self.normal), len(self.uniform)


shim = _Shim()


# ----------------------------------------------------------------------------------------------------
# Tensor: a thin wrapper so that operators convert their other operand the way TF does
# ----------------------------------------------------------------------------------------------------
class Tensor(object):
    __array_priority__ = 1000  # ndarray <op> Tensor defers to the Tensor's reflected operator

    def __init__(self, t):
        self.t = t

    # -- conversion / inspection --------------------------------------------------------------------
    def numpy(self):
        return self.t.detach().cpu().numpy()

    def eval(self, *a, **k):
        return self.numpy()

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype

    def get_shape(self):
        shp = tuple(self.t.shape)
        return _types.SimpleNamespace(as_list=lambda: list(shp))

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield Tensor(self.t[i])

    def __bool__(self):
        return builtins.bool(self.t)

    def __float__(self):
        return builtins.float(self.t)

    def __int__(self):
        return builtins.int(self.t)

    def __repr__(self):
        return "tf_shim.Tensor(%r)" % (self.t,)

    def __getitem__(self, idx):
        if isinstance(idx, Tensor):
            idx = idx.t
        return Tensor(self.t[idx])

    # -- arithmetic ---------------------------------------------------------------------------------
    def _o(self, other):
        return _cv(other, like=self).t

    def __add__(self, o): return Tensor(self.t + self._o(o))
    def __radd__(self, o): return Tensor(self._o(o) + self.t)
    def __sub__(self, o): return Tensor(self.t - self._o(o))
    def __rsub__(self, o): return Tensor(self._o(o) - self.t)
    def __mul__(self, o): return Tensor(self.t * self._o(o))
    def __rmul__(self, o): return Tensor(self._o(o) * self.t)
    def __truediv__(self, o): return Tensor(self.t / self._o(o))
    def __rtruediv__(self, o): return Tensor(self._o(o) / self.t)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return Tensor(self.t ** self._o(o))
    def __neg__(self): return Tensor(-self.t)
    def __ge__(self, o): return Tensor(self.t >= self._o(o))
    def __gt__(self, o): return Tensor(self.t > self._o(o))
    def __le__(self, o): return Tensor(self.t <= self._o(o))
    def __lt__(self, o): return Tensor(self.t < self._o(o))
    __hash__ = object.__hash__


def _cv(value, like=None, dtype=None):
    """ops.convert_to_tensor: Tensors pass, lists of Tensors stack, Python / numpy values become
    constants of ``dtype`` (else of ``like``'s dtype when both are floating, else the real dtype)."""
    if isinstance(value, Tensor):
        if dtype is not None and value.t.dtype != dtype:
            return Tensor(value.t.to(dtype))
        return value
    if isinstance(value, _torch.Tensor):
        return Tensor(value if dtype is None else value.to(dtype))
    if isinstance(value, (list, tuple)) and len(value) and any(isinstance(v, Tensor) for v in value):
        ref = next(v for v in value if isinstance(v, Tensor))
        return Tensor(_torch.stack([_cv(v, like=ref).t for v in value]))
    arr = _np.asarray(value)
    explicit64 = dtype == _torch.float64 and shim.real != _torch.float64
    if dtype is None:
        if arr.dtype.kind == "f":
            dtype = like.t.dtype if (like is not None and like.t.dtype.is_floating_point) else shim.real
        elif arr.dtype.kind in "iu":
            # a Python int next to a float tensor becomes that float type (x * 2, t + 1)
            if like is not None and like.t.dtype.is_floating_point:
                dtype = like.t.dtype
            else:
                dtype = like.t.dtype if like is not None and like.t.dtype != _torch.bool else _torch.int32
        elif arr.dtype.kind == "b":
            dtype = _torch.bool
        else:
            raise TypeError("cannot convert %r to a tensor" % (value,))
    if arr.dtype.kind == "f" and dtype == _torch.float64 and shim.real == _torch.float64 and not explicit64:
        # the fp64 run is the reference's *fp32 graph* in wider arithmetic: TF1 turns every Python / numpy
        # operand into a float32 constant first (0.1 -> fp32(0.1), 2*pi -> fp32(6.2831855)), so round once
        arr = arr.astype(_np.float32)
    return Tensor(_torch.as_tensor(arr).to(dtype))


def _t(value, like=None, dtype=None):
    return _cv(value, like=like, dtype=dtype).t


def _shape(shape):
    if isinstance(shape, Tensor):
        return tuple(builtins.int(s) for s in shape.numpy().reshape(-1))
    if isinstance(shape, (builtins.int, _np.integer)):
        return (builtins.int(shape),)
    return tuple(builtins.int(s) for s in shape)


# ----------------------------------------------------------------------------------------------------
# constants, shapes, casts
# ----------------------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype=_tdtype(dtype))
    if shape is not None:
        t = t.expand(_shape(shape)).clone()
    return Tensor(t)


def convert_to_tensor(value, dtype=None, name=None):
    return _cv(value, dtype=_tdtype(dtype))


def placeholder(dtype, shape=None, name=None):
    p = Tensor(None)
    p._ph_dtype = _tdtype(dtype)
    return p


def shape(x, name=None):  # noqa: F811
    return tuple(_cv(x).t.shape)  # static ints: every use in the reference indexes or passes it on


def cast(x, dtype, name=None):
    return Tensor(_t(x).to(_tdtype(dtype)))


def zeros(shape, dtype=float32, name=None):
    return Tensor(_torch.zeros(_shape(shape), dtype=_tdtype(dtype)))


def ones(shape, dtype=float32, name=None):
    return Tensor(_torch.ones(_shape(shape), dtype=_tdtype(dtype)))


def zeros_like(x, name=None):
    return Tensor(_torch.zeros_like(_t(x)))


def ones_like(x, name=None):
    return Tensor(_torch.ones_like(_t(x)))


def linspace(start, stop, num, name=None):
    # a constant-producing op: evaluated in float32 as the fp32 graph does, then widened (see _cv)
    return Tensor(_torch.as_tensor(_np.linspace(start, stop, builtins.int(num), dtype=_np.float32)).to(shim.real))


def identity(x, name=None):
    return _cv(x)


def stop_gradient(x, name=None):
    return Tensor(_t(x).detach())


def check_numerics(x, message, name=None):
    t = _t(x)
    if not _torch.isfinite(t).all():
        raise FloatingPointError(message)
    return Tensor(t)


# ----------------------------------------------------------------------------------------------------
# elementwise / reductions / linear algebra
# ----------------------------------------------------------------------------------------------------
def _bin(fn):
    def op(a, b, name=None):
        a_is, b_is = isinstance(a, Tensor), isinstance(b, Tensor)
        if a_is or not b_is:
            a = _cv(a)
            return Tensor(fn(a.t, _t(b, like=a)))
        return Tensor(fn(_t(a, like=b), b.t))
    return op


add = _bin(_torch.add)
subtract = _bin(_torch.sub)
multiply = _bin(_torch.mul)
divide = _bin(_torch.div)
minimum = _bin(_torch.minimum)
maximum = _bin(_torch.maximum)
less = _bin(_torch.lt)
less_equal = _bin(_torch.le)
greater = _bin(_torch.gt)
greater_equal = _bin(_torch.ge)
equal = _bin(_torch.eq)


def _un(fn):
    def op(x, name=None):
        return Tensor(fn(_t(x)))
    return op


exp, log, sqrt, square, sin, cos, tanh, abs = (_un(f) for f in (  # noqa: A001
    _torch.exp, _torch.log, _torch.sqrt, _torch.square, _torch.sin, _torch.cos, _torch.tanh, _torch.abs))
sigmoid = _un(_torch.sigmoid)
negative = _un(_torch.neg)
is_finite = _un(_torch.isfinite)
is_nan = _un(_torch.isnan)
log1p = _un(_torch.log1p)


def _axis(kw, axis):
    for k in ("reduction_indices", "axis"):
        if kw.get(k) is not None:
            axis = kw[k]
    return axis


def reduce_sum(x, axis=None, keepdims=False, name=None, **kw):
    axis = _axis(kw, axis)
    t = _t(x)
    return Tensor(t.sum() if axis is None else t.sum(dim=axis, keepdim=keepdims))


def reduce_mean(x, axis=None, keepdims=False, name=None, **kw):
    axis = _axis(kw, axis)
    t = _t(x)
    return Tensor(t.mean() if axis is None else t.mean(dim=axis, keepdim=keepdims))


def reduce_max(x, axis=None, keepdims=False, name=None, **kw):
    axis = _axis(kw, axis)
    t = _t(x)
    return Tensor(t.max() if axis is None else t.max(dim=axis, keepdim=keepdims).values)


def reduce_logsumexp(x, axis=None, keepdims=False, name=None, **kw):
    axis = _axis(kw, axis)
    t = _t(x)
    if axis is None:
        return Tensor(_torch.logsumexp(t.reshape(-1), dim=0))
    return Tensor(_torch.logsumexp(t, dim=axis, keepdim=keepdims))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a = _cv(a)
    ta, tb = a.t, _t(b, like=a)
    if transpose_a:
        ta = ta.transpose(-1, -2)
    if transpose_b:
        tb = tb.transpose(-1, -2)
    return Tensor(ta @ tb)


def transpose(x, perm=None, name=None):
    t = _t(x)
    return Tensor(t.permute(*perm) if perm is not None else t.permute(*reversed(range(t.dim()))))


def diag_part(x, name=None):
    return Tensor(_torch.diagonal(_t(x)))


def where(condition, x=None, y=None, name=None):
    c = _t(condition)
    x = _cv(x)
    ty = _t(y, like=x)
    if c.dim() == 1 and x.t.dim() > 1:  # TF1: a rank-1 condition picks whole rows
        c = c.reshape((-1,) + (1,) * (x.t.dim() - 1))
    return Tensor(_torch.where(c, x.t, ty))


def gather(params, indices, name=None):
    idx = _t(indices)
    p = _t(params)
    return Tensor(p[idx.long()] if idx.dim() else p[builtins.int(idx)])


def squeeze(x, axis=None, name=None):
    t = _t(x)
    return Tensor(t.squeeze() if axis is None else t.squeeze(axis))


def expand_dims(x, axis, name=None):
    return Tensor(_t(x).unsqueeze(axis))


def tile(x, multiples, name=None):
    return Tensor(_t(x).repeat(*_shape(multiples)))


def reshape(x, shape, name=None):  # noqa: F811
    return Tensor(_t(x).reshape(_shape(shape)))


def concat(values, axis, name=None):
    ref = next(v for v in values if isinstance(v, Tensor))
    return Tensor(_torch.cat([_t(v, like=ref) for v in values], dim=axis))


def stack(values, axis=0, name=None):
    values = list(values)  # the reference passes a py2 ``map`` result (utils/ais.py:82)
    ref = next(v for v in values if isinstance(v, Tensor))
    return Tensor(_torch.stack([_t(v, like=ref) for v in values], dim=axis))


def unstack(x, axis=0, name=None):
    return [Tensor(t) for t in _torch.unbind(_t(x), dim=axis)]


def split(x, num_or_size_splits, axis=0, name=None):
    t = _t(x)
    if isinstance(num_or_size_splits, (builtins.int, _np.integer)):
        return [Tensor(c) for c in _torch.chunk(t, builtins.int(num_or_size_splits), dim=axis)]
    return [Tensor(c) for c in _torch.split(t, list(num_or_size_splits), dim=axis)]


# ----------------------------------------------------------------------------------------------------
# control flow and differentiation
# ----------------------------------------------------------------------------------------------------
def while_loop(cond, body, loop_vars, **kw):
    vars_ = list(loop_vars)
    while builtins.bool(cond(*vars_)):
        vars_ = list(body(*vars_))
    return vars_


def scan(fn, elems, initializer=None, **kw):
    """tf.scan over the leading axis of one tensor; returns the stacked accumulators."""
    acc = initializer
    outs = []
    for e in _cv(elems):
        acc = fn(acc, e)
        outs.append(acc)
    if isinstance(acc, (tuple, list)):
        return tuple(stack([o[i] for o in outs]) for i in range(len(acc)))
    return stack(outs)


def gradients(ys, xs, grad_ys=None, name=None, **kw):
    """d(sum of ys)/d(xs): a list, one entry per x (None where unconnected), differentiable again."""
    single = not isinstance(xs, (list, tuple))
    xs_l = [xs] if single else list(xs)
    ys_l = ys if isinstance(ys, (list, tuple)) else [ys]
    total = None
    for i, y in enumerate(ys_l):
        ty = _t(y)
        term = ty.sum() if grad_ys is None else (ty * _t(grad_ys[i], like=_cv(y))).sum()
        total = term if total is None else total + term
    for x in xs_l:
        if not x.t.requires_grad:
            raise ValueError("tf_shim.gradients: x is not connected to a differentiable input "
                             "(wrap inputs with tf.shim.input)")
    gs = _torch.autograd.grad(total, [x.t for x in xs_l], create_graph=True, allow_unused=True)
    return [None if g is None else Tensor(g) for g in gs]


# ----------------------------------------------------------------------------------------------------
# randomness (injected)
# ----------------------------------------------------------------------------------------------------
def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):  # noqa: F811
    shp = _shape(shape)
    if not shim.normal:
        raise RuntimeError("tf_shim.random_normal%r: no injected draw left (shim.feed_random)" % (shp,))
    a = _np.asarray(shim.normal.popleft())
    if tuple(a.shape) != shp:
        raise ValueError("injected normal draw has shape %r, the reference asked for %r" % (a.shape, shp))
    shim.random_log.append(("normal", shp))
    t = _torch.as_tensor(a).to(_tdtype(dtype)).clone()
    t.requires_grad_(True)  # the notebook differentiates through z = tf.random_normal(...)
    t = t * stddev + mean if (mean != 0.0 or stddev != 1.0) else t
    return Tensor(t)


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):  # noqa: F811
    shp = _shape(shape)
    if not shim.uniform:
        raise RuntimeError("tf_shim.random_uniform%r: no injected draw left (shim.feed_random)" % (shp,))
    a = _np.asarray(shim.uniform.popleft())
    if tuple(a.shape) != shp:
        raise ValueError("injected uniform draw has shape %r, the reference asked for %r" % (a.shape, shp))
    td = _tdtype(dtype)
    if not td.is_floating_point:
        if maxval is None:
            raise ValueError("integer random_uniform needs maxval")
        if a.min() < minval or a.max() >= maxval:
            raise ValueError("injected integer draw outside [minval, maxval)")
    shim.random_log.append(("uniform", shp))
    return Tensor(_torch.as_tensor(a).to(td))


# ----------------------------------------------------------------------------------------------------
# variables
# ----------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def variable_scope(name, reuse=None, **kw):
    shim.scope.append(name)
    try:
        yield name
    finally:
        shim.scope.pop()


name_scope = variable_scope


def constant_initializer(value=0.0, dtype=float32):
    def init(shape, dtype_=None):
        return _np.full(_shape(shape), value, dtype=_np.float64)
    return init


def zeros_initializer(dtype=float32):
    return constant_initializer(0.0)


def _variance_scaling_initializer(factor=2.0, mode="FAN_IN", uniform=False, seed=None, dtype=float32):
    """tf.contrib.layers.variance_scaling_initializer (TF 1.x contrib): with uniform=False a truncated
    normal (|z| <= 2 resampled) of stddev sqrt(1.3 * factor / n), n = fan_in / fan_out / their mean."""
    def init(shape, dtype_=None):
        shp = _shape(shape)
        fan_in = builtins.float(shp[-2]) if len(shp) > 1 else builtins.float(shp[-1])
        fan_out = builtins.float(shp[-1])
        n = {"FAN_IN": fan_in, "FAN_OUT": fan_out, "FAN_AVG": (fan_in + fan_out) / 2.0}[mode]
        if uniform:
            limit = _math.sqrt(3.0 * factor / n)
            return shim.rng.uniform(-limit, limit, size=shp)
        std = _math.sqrt(1.3 * factor / n)
        z = shim.rng.standard_normal(size=shp)
        bad = _np.abs(z) > 2.0
        while bad.any():
            z[bad] = shim.rng.standard_normal(size=builtins.int(bad.sum()))
            bad = _np.abs(z) > 2.0
        return z * std
    return init


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):  # noqa: F811
    full = "/".join(shim.scope + [name])
    if full in shim.variables:
        raise ValueError("Variable %s already exists (TF1 would need reuse=True)" % full)
    if full in shim.preloaded:
        val = _np.asarray(shim.preloaded[full])
        if shape is not None and tuple(val.shape) != _shape(shape):
            raise ValueError("preloaded %s has shape %r, the reference declares %r" % (full, val.shape, _shape(shape)))
        t = _torch.as_tensor(val).to(_tdtype(dtype) or shim.real).clone()
    elif isinstance(initializer, Tensor):
        t = initializer.t.detach().clone()
    elif callable(initializer):
        t = _torch.as_tensor(_np.asarray(initializer(shape))).to(_tdtype(dtype) or shim.real).clone()
    else:
        raise ValueError("get_variable(%s): no initializer and nothing preloaded" % full)
    t.requires_grad_(builtins.bool(trainable) and t.dtype.is_floating_point)
    v = Tensor(t)
    v.name = full + ":0"
    v.trainable = builtins.bool(trainable)
    shim.variables[full] = v
    return v


def Variable(initial_value, trainable=True, name=None, dtype=None):
    t = _t(initial_value, dtype=_tdtype(dtype)).detach().clone()
    t.requires_grad_(builtins.bool(trainable) and t.dtype.is_floating_point)
    v = Tensor(t)
    full = "/".join(shim.scope + [name or "Variable_%d" % len(shim.variables)])
    v.name = full + ":0"
    v.trainable = builtins.bool(trainable)
    shim.variables[full] = v
    return v


class GraphKeys(object):
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "trainable_variables"


def get_collection(key, scope=None):
    vs = [v for k, v in shim.variables.items() if scope is None or k.startswith(scope)]
    if key == GraphKeys.TRAINABLE_VARIABLES:
        vs = [v for v in vs if v.trainable]
    return vs


def trainable_variables():
    return get_collection(GraphKeys.TRAINABLE_VARIABLES)


def global_variables_initializer():
    return None


# ----------------------------------------------------------------------------------------------------
# tf.nn, tf.contrib
# ----------------------------------------------------------------------------------------------------
def _sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
    """max(l, 0) - l * z + log(1 + exp(-|l|))  (the formula TF documents and implements)."""
    l = _cv(logits)
    z = _t(labels, like=l)
    lt = l.t
    return Tensor(_torch.clamp(lt, min=0) - lt * z + _torch.log1p(_torch.exp(-_torch.abs(lt))))


nn = _types.SimpleNamespace(
    relu=_un(_torch.relu),
    tanh=tanh,
    sigmoid=sigmoid,
    softplus=_un(_torch.nn.functional.softplus),
    sigmoid_cross_entropy_with_logits=_sigmoid_cross_entropy_with_logits,
)

from . import contrib  # noqa: E402,F401  (tf.contrib.layers.variance_scaling_initializer)
