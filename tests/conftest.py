import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = Path("/root/reference/src/smolyax").exists()
    skip_ref = pytest.mark.skip(reason="/root/reference is not present on this machine")
    for item in items:
        if item.get_closest_marker("reference") is not None and not have_ref:
            item.add_marker(skip_ref)
