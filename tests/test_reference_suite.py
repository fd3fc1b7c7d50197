"""Run the reference's OWN test module for the edge classes against the drop-in (only where /root/reference exists).

`rectipy_tests/test_edges.py` is executed unmodified with `rectipy` aliased to `rectipy_b200`, so the upstream assertions
(shape/transposition rules, dtype handling, trainable-parameter counting, Linear == torch.nn.Linear, error types, RLS
init/forward/update) are checked verbatim.  The node/network test modules of the reference construct their models through
PyRates objects and run on CPU, which this engine does not do; their assertions are mirrored in tests/test_host.py and
tests/test_gpu_api.py instead."""
import importlib.util
import os
import sys
import types

import pytest

REF = os.environ.get("RECTIPY_REFERENCE", "/root/reference")
TEST_FILE = os.path.join(REF, "rectipy_tests", "test_edges.py")


@pytest.mark.skipif(not os.path.exists(TEST_FILE), reason="reference tree not present (GPU box)")
def test_reference_test_edges_passes_against_drop_in():
    import rectipy_b200
    import rectipy_b200.edges
    saved = {k: sys.modules.get(k) for k in ("rectipy", "rectipy.edges")}
    alias = types.ModuleType("rectipy")
    alias.edges = rectipy_b200.edges
    sys.modules["rectipy"] = alias
    sys.modules["rectipy.edges"] = rectipy_b200.edges
    try:
        spec = importlib.util.spec_from_file_location("ref_test_edges", TEST_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ran = 0
        for name in sorted(dir(mod)):
            if name.startswith("test_"):
                getattr(mod, name)()
                ran += 1
        assert ran >= 2
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
