"""tf.contrib namespace of the eager stand-in (TEST INFRASTRUCTURE ONLY; see ../__init__.py)."""
import types as _types

import tensorflow as _tf

layers = _types.SimpleNamespace(variance_scaling_initializer=_tf._variance_scaling_initializer)


class _HParams(object):  # tf.contrib.training.HParams: attribute bag (mnist_vae.py:40)
    def __init__(self, **kw):
        self.__dict__.update(kw)


training = _types.SimpleNamespace(HParams=_HParams)
