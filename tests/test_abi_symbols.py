"""The C-ABI library loads on a CPU-only box and exports every symbol include/pymd_b200.h
declares; the ctypes prototype table covers the same set (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pymd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pymd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("pymd_ctx_create", "pymd_step", "pymd_noise_generate", "pymd_powerspec", "pymd_init_state",
                 "pymd_bath_set", "pymd_philox_normal", "pymd_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(ge.LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), "libpymd_b200.so does not export %s" % name


def test_ctypes_table_matches_header():
    from pymd_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    lib = _lib.load()
    assert lib.pymd_version() >= 100
    assert lib.pymd_bath_ncp(15) == 16 and lib.pymd_ring_len(100, 0) >= 99 + 16
    assert lib.pymd_ring_len(1, 0) == 0


def test_errors_are_reported_not_swallowed():
    """Bad arguments return a negative status with a message (the reference print+exit convention
    becomes an exception in Python)."""
    import pytest
    from pymd_b200 import _lib
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.pymd_ctx_create(12, 33, 1, 1.0, 16, 0, ctypes.byref(ctx))   # ldb not a multiple of 32
    assert rc < 0 and b"multiple of 32" in lib.pymd_last_error()
    with pytest.raises(_lib.PymdError):
        _lib.call("pymd_ctx_create", 0, 32, 1, 1.0, 16, 0, ctypes.byref(ctx))
