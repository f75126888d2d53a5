import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """ADB_EW_STATS=<file>: dump which elementwise programs the GPU tests ran (tools/gen_ewise_catalog.py --merge)."""
    import os
    if os.environ.get("ADB_TAPE_SELFCHECK") == "1":
        try:
            from autodiff_b200.core import tape
            print("\nlaunch-tape self-check: %s" % (tape.stats,))
        except Exception:
            pass
    path = os.environ.get("ADB_EW_STATS")
    if path:
        try:
            import autodiff_b200 as ad
            open(path, "w").write(ad.ewise_stats())
        except Exception:
            pass
