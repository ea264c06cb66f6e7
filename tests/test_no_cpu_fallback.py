"""There is no CPU fallback: without a CUDA device (or without the in-tree shared library) the product
path raises instead of computing anything on the host, and nothing under pymd_b200/ imports the
oracle (test infrastructure)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_package_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|pymd_oracle|oracle\._ref|oracle/_ref")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pymd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                for n, line in enumerate(open(os.path.join(dirpath, f), errors="replace"), 1):
                    assert not pat.search(line), "%s:%d mentions the oracle: %s" % (f, n, line.strip())


def test_stepping_without_a_gpu_raises():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import pymd_testlib as T
    c = T.CASES["ph2"]
    with pytest.raises(Exception) as e:
        m, j = T.device_md(c, 1)
        m.initialise()
        m.ResetHis()
        for b in m.baths:
            b.gnoi()
        m.vv()
    assert not isinstance(e.value, (AssertionError, AttributeError, TypeError))


def test_missing_library_is_reported(monkeypatch, tmp_path):
    from pymd_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libpymd_b200.so"))
    with pytest.raises((_lib.PymdError, OSError)):
        _lib.load()
