import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

GOLDEN_DIR = os.path.join(HERE, "golden")
REFERENCE_ASSETS = "/root/reference/tests/Assets"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import jpeglibrary_b200 as J
        return J._native.cuda.jb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))


def golden_bytes(name):
    return open(os.path.join(GOLDEN_DIR, name), "rb").read()
