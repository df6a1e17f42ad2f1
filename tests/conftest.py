import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not failed) on hosts without a CUDA device."""
    if any("gpu" in item.keywords for item in items) and not has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


def has_gpu() -> bool:
    try:
        from dpo_b200 import _capi
        import ctypes
        lib = _capi.load_library()
        c = ctypes.c_int(0)
        return lib.dpgo_device_count(ctypes.byref(c)) == 0 and c.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def data_dir():
    return DATA


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
