"""Stand-in for tensorflow.examples.tutorials.mnist.input_data (TEST INFRASTRUCTURE ONLY): the import in
utils/func_utils.py:28 must succeed; reading MNIST is impossible here (no network, no dataset)."""


def read_data_sets(*a, **k):
    raise RuntimeError("tf_shim: MNIST is not available in this environment")
