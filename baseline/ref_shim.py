"""Loads the UNMODIFIED reference package from baseline/_ref (installed by baseline/install_ref.sh) for the CPU arm of
bench.py and for tests.  Two host-side shims that touch no arithmetic:
  * a stub `graphviz` module - autodiff/__init__.py:7 imports the drawing code unconditionally and graphviz is not installed;
  * `Slice.__init__` tuple-ises a LIST of slices - Concat/Pad gradients index with a list (reshape.py:41-48,133-134),
    which numpy >= 1.23 rejects (old numpy treated the list as a tuple, so behaviour is preserved).
Nothing of this repo's product code is on the path of the module returned here."""
import importlib
import os
import sys
import types

_REF = None
REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "autodiff"))


def load():
    """The reference's top-level module (`import autodiff as ad`), or None when baseline/_ref is not installed."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        return None
    if "graphviz" not in sys.modules:
        g = types.ModuleType("graphviz")
        g.Digraph = type("Digraph", (), {"__init__": lambda self, *a, **k: None})
        sys.modules["graphviz"] = g
    if sys.getrecursionlimit() < 50000:
        sys.setrecursionlimit(50000)           # Node.eval recurses through the graph (node.py:42-46)
    sys.path.insert(0, REF_DIR)
    try:
        ref = importlib.import_module("autodiff")
    finally:
        sys.path.remove(REF_DIR)
    if not os.path.abspath(ref.__file__).startswith(REF_DIR):
        raise RuntimeError("`autodiff` resolved to %s, not to baseline/_ref" % ref.__file__)
    from autodiff.core import reshape as _r
    orig = _r.Slice.__init__

    def _init(self, node, slice_val, name="Slice"):
        orig(self, node, tuple(slice_val) if isinstance(slice_val, list) else slice_val, name)
    _r.Slice.__init__ = _init
    _REF = ref
    return ref
