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


def test_fft_plan_lengths():
    """Host-side planning of the column FFT: fewest sweeps over HBM (fused radix-128/64/16 passes first)."""
    from pymd_b200 import _lib
    lib = _lib.load()
    want = {2: 1, 4: 1, 8: 1, 16: 1, 32: 2, 64: 1, 128: 1, 256: 2, 1024: 2, 2048: 2, 4096: 2, 8192: 2, 16384: 2,
            32768: 3, 65536: 3, 131072: 3}
    for n, np_ in want.items():
        assert lib.pymd_fft_num_passes(n) == np_, n
    assert lib.pymd_fft_num_passes(0) == -1 and lib.pymd_fft_num_passes(96) == -1


def test_setup_kernel_size_limits():
    """Shared-memory limits of the set-up kernels (host-side queries, no GPU needed)."""
    from pymd_b200 import _lib
    lib = _lib.load()
    assert lib.pymd_harmonic_max_n() >= 96                   # config 3's nd
    assert lib.pymd_eigh_max_n(1) >= 96 and lib.pymd_eigh_max_n(0) >= 118
    assert lib.pymd_eigh_work_bytes(100, 60, 1) == 0         # fits shared memory twice: no scratch
    assert lib.pymd_eigh_work_bytes(100, 96, 1) == 100 * 96 * 96 * 16
    assert lib.pymd_eigh_work_bytes(100, 500, 1) == 0        # unsupported size: nothing to allocate
