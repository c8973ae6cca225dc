"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.

Imports the UNMODIFIED reference modules G-Meta/learner.py, G-Meta/meta.py and
G-Meta/subgraph_data_processing.py from /root/reference with oracle/dgl_shim on
sys.path in place of the (uninstallable) dgl wheel.  Used only

  * by oracle/make_golden.py to generate tests/golden/*.npz, and
  * by tests/test_oracle_vs_reference.py to pin oracle/gmeta_oracle.py
    bit-for-bit against the reference's own arithmetic,

and only in the build container: /root/reference does not exist on the GPU box,
where `available()` returns False and the dependent tests skip.
"""
import importlib
import os
import sys
import warnings

REFERENCE_DIR = os.environ.get("GMETA_REFERENCE_DIR", "/root/reference/G-Meta")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dgl_shim")

_cache = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "meta.py"))


def load():
    """Return (learner, meta, subgraph_data_processing) reference modules."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_DIR)
    for p in (_SHIM_DIR, REFERENCE_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # `is 'Linear'` SyntaxWarnings (learner.py:83,...)
        learner = importlib.import_module("learner")
        meta = importlib.import_module("meta")
        sdp = importlib.import_module("subgraph_data_processing")
    assert os.path.abspath(learner.__file__).startswith(os.path.abspath(REFERENCE_DIR))
    assert os.path.abspath(meta.__file__).startswith(os.path.abspath(REFERENCE_DIR))
    _cache["mods"] = (learner, meta, sdp)
    return _cache["mods"]


def shim_dgl():
    if _SHIM_DIR not in sys.path:
        sys.path.insert(0, _SHIM_DIR)
    return importlib.import_module("dgl")
