"""TEST INFRASTRUCTURE ONLY — import the untouched reference (build container only).

``/root/reference`` exists only in the build container; the GPU box never has it, so
nothing reachable from ``-m gpu`` tests / ``smoke()`` / ``bench.py`` imports this module.
It puts a stub for the reference's absent private ``todos`` debug package ahead of the
reference on ``sys.path`` and turns ``pdb.set_trace`` into a no-op (the fork leaves
breakpoints at cWCT.py:36,:118 and RevResNet.py:242).  No reference file is modified.
"""
import os
import pdb
import sys
import types

REF_ROOT = os.environ.get("VST_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "RevResNet.py"))


def load():
    """Returns (RevResNet, cWCT) classes of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "todos" not in sys.modules:
        stub = types.ModuleType("todos")
        stub.debug = types.SimpleNamespace(output_var=lambda *a, **k: None)
        sys.modules["todos"] = stub
    pdb.set_trace = lambda *a, **k: None
    # the reference's top-level package is called `models`; load it under a private
    # name so it cannot collide with this repo's own `models/` drop-in shim.
    import importlib.util
    out = []
    for name in ("RevResNet", "cWCT"):
        spec = importlib.util.spec_from_file_location(
            "_vst_reference_" + name, os.path.join(REF_ROOT, "models", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out.append(getattr(mod, name))
    return tuple(out)
