"""Import the UNMODIFIED reference (`/root/reference/rectipy`) in the build container.  TEST INFRASTRUCTURE ONLY.

The reference imports three packages that are absent here (`pyrates`, `multipledispatch`, `matplotlib`).
They are only needed for YAML->vector-field generation, `Network.__getitem__` overloading and plotting, so
minimal stand-ins are placed in ``sys.modules``.  The reference's node/edge/network/
observer classes then run unchanged, driven by the hand-written vector fields from ``rectipy_oracle`` through
the public ``RateNet(rnn_func, rnn_args, var_map, param_map, ...)`` constructor (nodes.py:58) -- the same
route the reference's own test uses (rectipy_tests/test_nodes.py:32-33,53).

This module only works where ``/root/reference`` exists (the build container).  It is used by
``oracle/make_golden.py`` to mint the fixtures under ``tests/golden/``; nothing on the GPU box imports it.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RECTIPY_REFERENCE", "/root/reference")


def _install_stubs():
    if "pyrates" not in sys.modules:
        m = types.ModuleType("pyrates")

        class NodeTemplate:  # only touched inside from_pyrates (nodes.py:238)
            pass

        class CircuitTemplate:
            pass

        m.NodeTemplate, m.CircuitTemplate = NodeTemplate, CircuitTemplate
        m.clear = lambda *a, **k: None
        m.clear_frontend_caches = lambda *a, **k: None
        sys.modules["pyrates"] = m
    if "multipledispatch" not in sys.modules:
        m = types.ModuleType("multipledispatch")
        registry = {}

        def dispatch(*types_):
            def deco(fn):
                table = registry.setdefault(fn.__qualname__, [])
                table.append((types_, fn))

                def call(self, arg):
                    for tys, f in table:
                        if isinstance(arg, tys[0]):
                            return f(self, arg)
                    raise NotImplementedError(f"no dispatch for {type(arg)}")
                return call
            return deco

        m.dispatch = dispatch
        sys.modules["multipledispatch"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.Axes = object
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def import_reference():
    """Return the reference's ``rectipy`` package (unmodified source, stubbed third-party imports)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "rectipy")):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}; ref_shim only works in the build container")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import rectipy  # noqa: E402
    return rectipy
