import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as h:
        return json.load(h)


@pytest.fixture(scope="session")
def plastome():
    return _load("plastome_cfg1.json.gz")


@pytest.fixture(scope="session")
def synthetic_small():
    return _load("synthetic_small.json.gz")


@pytest.fixture(scope="session")
def subpaths_golden():
    return _load("subpaths.json.gz")


@pytest.fixture(scope="session")
def lib_built():
    """libtvs.so must exist for both CPU (symbol checks) and GPU tests; build it if absent."""
    from traversome_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import subprocess
        subprocess.check_call([os.path.join(ROOT, "traversome_b200", "csrc", "build.sh")])
    return _cabi.LIB_PATH
