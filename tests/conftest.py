import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _have_b200() -> bool:
    try:
        import ctypes as C

        from wekua_b200 import capi

        n = C.c_int32(0)
        return capi.lib().wk_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """a plain `pytest` on a machine without a B200 skips the -m gpu tests instead of failing every one of them"""
    if any(i.get_closest_marker("gpu") for i in items) and not _have_b200():
        skip = pytest.mark.skip(reason="no CUDA device (the -m gpu tests run under gpurun on a B200)")
        for i in items:
            if i.get_closest_marker("gpu"):
                i.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.lib()  # builds on first use
    return pyoracle
