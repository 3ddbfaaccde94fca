"""pytest configuration: markers, import path, build of the test-side native
helpers (the oracle's C restatement and the host build of svt_semantics.h)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        from sparsearray_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_helpers():
    """Build whatever is missing (no-ops when the prebuilt files travelled)."""
    from sparsearray_b200 import build
    if not (os.path.exists(build.LIBSVTGPU) and os.path.exists(build.LIBRGLUE)):
        build.build_all()
    from oracle import port
    port.build()
    yield
