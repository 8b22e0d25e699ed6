"""Import the UNMODIFIED reference sources from /root/reference on top of the eager TensorFlow stand-in
(oracle/tf_shim).  TEST INFRASTRUCTURE ONLY: used by tests/golden/make_ref_golden.py (fixture
generation, run in the build container where /root/reference exists) and by the CPU tests that re-run
the reference when it is present.  Nothing here runs on the GPU box and the package never imports it.

No reference text is edited.  The only transformation is ``source.expandtabs(8)`` for
utils/sampler.py, whose body mixes tabs and spaces the way Python 2 accepted (tab = next multiple of
eight columns, utils/sampler.py:29-51) and Python 3 refuses to tokenize; expanding tabs is exactly
Python 2's reading of that file.  Modules are compiled from the files where they lie, by path.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(HERE, "tf_shim")
REF_ROOT = os.environ.get("L2HMC_REFERENCE", "/root/reference")

# load order follows the reference's own imports (utils/ais.py:27-28, utils/notebook_utils.py)
_FILES = ["layers", "distributions", "dynamics", "losses", "func_utils", "sampler", "ais"]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "dynamics.py"))


def _tf():
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    tf = importlib.import_module("tensorflow")
    if not hasattr(tf, "shim"):
        raise RuntimeError("a real tensorflow shadows oracle/tf_shim; the fixtures are defined on the stand-in")
    return tf


def _compile(path):
    with open(path, "rb") as fh:
        src = fh.read().decode("utf-8")
    if "\t" in src:
        src = src.expandtabs(8)  # Python 2's tab rule; the only transformation applied to reference text
    return compile(src, path, "exec")


class Reference(types.SimpleNamespace):
    """ref.tf (the stand-in), ref.dynamics / layers / distributions / losses / func_utils / sampler / ais
    (the reference's modules), ref.notebook_network (SCGExperiment.ipynb cell 3, executed verbatim)."""


_cached = None


def load() -> Reference:
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError("reference sources not found under %s" % REF_ROOT)
    tf = _tf()
    ref = Reference(tf=tf, root=REF_ROOT)
    # the reference uses Python-2 implicit relative imports (`from dynamics import Dynamics`,
    # utils/ais.py:27): expose the already-loaded siblings under their bare names while loading
    saved = {k: sys.modules.get(k) for k in _FILES}
    try:
        for name in _FILES:
            path = os.path.join(REF_ROOT, "utils", name + ".py")
            mod = types.ModuleType(name)
            mod.__file__ = path
            sys.modules[name] = mod
            exec(_compile(path), mod.__dict__)
            setattr(ref, name, mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ref.notebook_network = _notebook_network(ref)
    _cached = ref
    return ref


def _notebook_network(ref):
    """`network(x_dim, scope, factor)` of SCGExperiment.ipynb (the cell that defines it), source executed as is."""
    with open(os.path.join(REF_ROOT, "SCGExperiment.ipynb")) as fh:
        nb = json.load(fh)
    for cell in nb["cells"]:
        src = "".join(cell["source"])
        if cell["cell_type"] == "code" and src.lstrip().startswith("def network("):
            ns = {"tf": ref.tf}
            for k in ("Linear", "Sequential", "Zip", "Parallel", "ScaleTanh"):
                ns[k] = getattr(ref.layers, k)
            exec(compile(src, "SCGExperiment.ipynb:network", "exec"), ns)
            return ns["network"]
    raise RuntimeError("network() cell not found in SCGExperiment.ipynb")
