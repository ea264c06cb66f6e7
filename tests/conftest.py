import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected on a box without CUDA;
    the CPU round (`-m "not gpu"`) deselects them."""
    return


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np

    class G:
        def __init__(self):
            with open(os.path.join(GOLDEN, "manifest.json")) as f:
                self.manifest = json.load(f)
            self._c = {}

        def cfg(self, name):
            return self.manifest["cases"][name]

        def __call__(self, name):
            if name not in self._c:
                self._c[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
            return self._c[name]
    return G()

sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.fixture(scope="session")
def cuda():
    """GPU tests must FAIL (not skip) without a device or without the native library."""
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    import __graft_entry__ as ge
    ge.build()
    from pymd_b200 import _lib
    _lib.load()
    return torch.device("cuda", 0)
