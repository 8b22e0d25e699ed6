"""Layer combinators used to *describe* the S/T/Q nets (mirror of the reference's utils/layers.py).

Same names, constructor arguments and call behaviour as /root/reference/utils/layers.py:
``Linear`` (:29-37), ``ConcatLinear`` (:40-58), ``Parallel`` (:60-66), ``Sequential`` (:68-79),
``ScaleTanh`` (:81-86), ``Zip`` (:88-95).  Parameters are plain fp32 torch tensors (host memory);
``relu`` / ``softplus`` stand in for ``tf.nn.relu`` / ``tf.nn.softplus``.

The objects stay callable on torch tensors (eager, any device) so a net description can be inspected
or unit-tested, but ``Dynamics`` never evaluates them on the sampling path: ``compile_stq_net``
pattern-matches the canonical ``net_factory`` structure (SCGExperiment.ipynb:51-77) and hands the
weights to the CUDA library.  A net that does not match raises -- there is no CPU fallback.
"""
from __future__ import annotations

import builtins
import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

TORCH_FLOAT = torch.float32
NP_FLOAT = np.float32

_default_generator = torch.Generator().manual_seed(0)


def manual_seed(seed: int) -> None:
    """Seed the initializer RNG (the reference leaves TF's global seed unset)."""
    _default_generator.manual_seed(int(seed))


def _variance_scaling(shape, factor: float, generator=None) -> torch.Tensor:
    """tf.contrib.layers.variance_scaling_initializer(factor, mode='FAN_IN', uniform=False):
    truncated normal (+-2 sigma) with stddev sqrt(1.3 * factor / fan_in) (utils/layers.py:32)."""
    fan_in = shape[0]
    std = math.sqrt(1.3 * factor / fan_in)
    w = torch.empty(shape, dtype=TORCH_FLOAT)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2.0 * std, b=2.0 * std,
                                generator=generator or _default_generator)
    return w


def relu(x):
    return torch.relu(x) if isinstance(x, torch.Tensor) else np.maximum(x, 0)


def softplus(x):
    return torch.nn.functional.softplus(x)


class Linear(object):
    def __init__(self, in_, out_, scope='linear', factor=1.0):
        self.scope = scope
        self.in_, self.out_ = int(in_), int(out_)
        self.W = _variance_scaling((self.in_, self.out_), factor * 2.0)
        self.b = torch.zeros(self.out_, dtype=TORCH_FLOAT)

    def __call__(self, x):
        return torch.add(torch.matmul(x, self.W.to(x.device)), self.b.to(x.device))


class ConcatLinear(object):
    def __init__(self, ins_, out_, factors=None, scope='concat_linear'):
        self.layers = []
        for i, in_ in enumerate(ins_):
            factor = 1.0 if factors is None else factors[i]
            self.layers.append(Linear(in_, out_, scope='linear_%d' % i, factor=factor))

    def __call__(self, inputs):
        output = 0.
        for i, x in enumerate(inputs):
            output += self.layers[i](x)
        return output


class Parallel(object):
    def __init__(self, layers=None):
        self.layers = [] if layers is None else layers

    def add(self, layer):
        self.layers.append(layer)

    def __call__(self, x):
        return [layer(x) for layer in self.layers]


class Sequential(object):
    def __init__(self, layers=None):
        self.layers = [] if layers is None else layers

    def add(self, layer):
        self.layers.append(layer)

    def __call__(self, x):
        y = x
        for layer in self.layers:
            y = layer(y)
        return y


class ScaleTanh(object):
    def __init__(self, in_, scope='scale_tanh'):
        self.scope = scope
        self.log_scale = torch.zeros((1, int(in_)), dtype=TORCH_FLOAT)  # variable 'scale', init 0

    @property
    def scale(self):
        return torch.exp(self.log_scale)

    def __call__(self, x):
        return self.scale.to(x.device) * torch.tanh(x)


class Zip(object):
    def __init__(self, layers=None):
        self.layers = [] if layers is None else layers

    def __call__(self, x):
        assert len(x) == len(self.layers)
        n = len(self.layers)
        return [self.layers[i](x[i]) for i in range(n)]


# --------------------------------------------------------------------------------------------------
# Net description -> packed parameters
# --------------------------------------------------------------------------------------------------
class NetStructureError(TypeError):
    pass


def _is_relu(f) -> bool:
    return f is relu or f is torch.relu or f is torch.nn.functional.relu or getattr(f, "__name__", "") == "relu"


def _is_sum(f) -> bool:
    return f is builtins.sum or f is sum


def _zero_aux(f) -> bool:
    """The 4th Zip entry of the canonical net is ``lambda _: 0.`` (SCGExperiment.ipynb:58)."""
    if isinstance(f, (Linear, Sequential, Parallel, Zip)):
        return False
    try:
        r = f(None)
    except Exception:
        return False
    return isinstance(r, (int, float)) and r == 0


def _is_softplus(f) -> bool:
    return (f is softplus or f is torch.nn.functional.softplus or getattr(f, "__name__", "") == "softplus")


def compile_softplus_mlp(seq, what="softplus MLP"):
    """Weights of Sequential([Linear, softplus, Linear, ..., Linear]) (the decoder, mnist_vae.py:104-111, and the
    nets' aux encoder, mnist_vae.py:134-140) as (widths, [W...], [b...]) fp32 arrays."""
    if not isinstance(seq, Sequential) or len(seq.layers) < 1 or len(seq.layers) % 2 != 1:
        raise NetStructureError("%s must be Sequential([Linear, softplus, ..., Linear])" % what)
    Ws, bs, widths = [], [], []
    for i, l in enumerate(seq.layers):
        if i % 2 == 0:
            if not isinstance(l, Linear):
                raise NetStructureError("%s: stage %d must be Linear" % (what, i))
            if widths and widths[-1] != l.in_:
                raise NetStructureError("%s: Linear %s takes %d inputs after a layer of %d" % (what, l.scope, l.in_, widths[-1]))
            if not widths:
                widths.append(l.in_)
            widths.append(l.out_)
            Ws.append(np.ascontiguousarray(l.W.detach().cpu().numpy().astype(np.float32)))
            bs.append(np.ascontiguousarray(l.b.detach().cpu().numpy().astype(np.float32)))
        elif not _is_softplus(l):
            raise NetStructureError("%s: stage %d must be softplus" % (what, i))
    return widths, Ws, bs


def compile_stq_net(net, x_dim: int) -> Dict[str, np.ndarray]:
    """Extract the weights of a canonical S/T/Q net as fp32 arrays keyed like the C ABI struct
    (include/l2hmc.h: l2hmc_net_params).  Accepted structure (SCGExperiment.ipynb:51-77):

        Sequential([Zip([Linear(D,H), Linear(D,H), Linear(2,H), <zero aux | softplus-MLP encoder of aux>]), sum, relu,
                    Linear(H,H), relu,
                    Parallel([Sequential([Linear(H,D), ScaleTanh(D)]), Linear(H,D),
                              Sequential([Linear(H,D), ScaleTanh(D)])])])
    """
    def bad(msg):
        raise NetStructureError("net_factory produced a net the CUDA path cannot take: " + msg)

    if not isinstance(net, Sequential) or len(net.layers) != 6:
        bad("expected Sequential of 6 stages (Zip, sum, relu, Linear, relu, Parallel)")
    z, s, r1, lin, r2, par = net.layers
    if not isinstance(z, Zip) or len(z.layers) not in (3, 4):
        bad("stage 0 must be Zip of 3 Linear layers (+ aux branch)")
    e1, e2, e3 = z.layers[:3]
    if not all(isinstance(e, Linear) for e in (e1, e2, e3)):
        bad("Zip entries 0-2 must be Linear")
    aux_enc = None
    if len(z.layers) == 4 and not _zero_aux(z.layers[3]):
        # encoder_sampler of mnist_vae.py:134-149: a Linear/softplus stack of aux, shared by XNet and VNet
        if not isinstance(z.layers[3], Sequential):
            bad("the aux branch (4th Zip entry) must be `lambda _: 0.` or Sequential([Linear, softplus, ..., Linear])")
        aux_enc = z.layers[3]
    if not _is_sum(s) or not _is_relu(r1) or not _is_relu(r2):
        bad("stages 1,2,4 must be sum, relu, relu")
    if not isinstance(lin, Linear):
        bad("stage 3 must be Linear")
    if not isinstance(par, Parallel) or len(par.layers) != 3:
        bad("stage 5 must be Parallel of 3 heads")
    hs, ht, hq = par.layers

    def head_scaled(h):
        if not (isinstance(h, Sequential) and len(h.layers) == 2 and isinstance(h.layers[0], Linear)
                and isinstance(h.layers[1], ScaleTanh)):
            bad("S and Q heads must be Sequential([Linear, ScaleTanh])")
        return h.layers[0], h.layers[1]

    ls_, ss_ = head_scaled(hs)
    lq_, sq_ = head_scaled(hq)
    if not isinstance(ht, Linear):
        bad("T head must be Linear")
    H = e1.out_
    D = int(x_dim)
    shapes = [(e1, D, H), (e2, D, H), (e3, 2, H), (lin, H, H), (ls_, H, D), (ht, H, D), (lq_, H, D)]
    for l, i, o in shapes:
        if (l.in_, l.out_) != (i, o):
            bad("Linear %s has shape (%d,%d), expected (%d,%d)" % (l.scope, l.in_, l.out_, i, o))

    def f(t):
        return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))

    return {
        "W1": f(e1.W), "b1": f(e1.b), "W2": f(e2.W), "b2": f(e2.b), "W3": f(e3.W), "b3": f(e3.b),
        "W4": f(lin.W), "b4": f(lin.b), "Ws": f(ls_.W), "bs": f(ls_.b), "Wt": f(ht.W), "bt": f(ht.b),
        "Wq": f(lq_.W), "bq": f(lq_.b),
        "ls": f(ss_.log_scale).reshape(-1), "lq": f(sq_.log_scale).reshape(-1),
        "aux_encoder": aux_enc,
    }


def load_stq_net(net, params: Dict[str, np.ndarray]) -> None:
    """Write a parameter dict (same keys as compile_stq_net) into a canonical net's layers."""
    z, _, _, lin, _, par = net.layers
    e1, e2, e3 = z.layers[:3]
    hs, ht, hq = par.layers
    pairs = [(e1, "W1", "b1"), (e2, "W2", "b2"), (e3, "W3", "b3"), (lin, "W4", "b4"),
             (hs.layers[0], "Ws", "bs"), (ht, "Wt", "bt"), (hq.layers[0], "Wq", "bq")]
    for l, w, b in pairs:
        l.W = torch.as_tensor(np.asarray(params[w], np.float32)).clone()
        l.b = torch.as_tensor(np.asarray(params[b], np.float32)).clone()
    hs.layers[1].log_scale = torch.as_tensor(np.asarray(params["ls"], np.float32)).reshape(1, -1).clone()
    hq.layers[1].log_scale = torch.as_tensor(np.asarray(params["lq"], np.float32)).reshape(1, -1).clone()
