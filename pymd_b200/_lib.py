"""ctypes binding of libpymd_b200.so (the C ABI declared in include/pymd_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails this module
raises.  Build the library with ``python __graft_entry__.py`` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYMD_B200_LIB", os.path.join(_HERE, "libpymd_b200.so"))   # override: debug builds only

_lib = None

c_dp = C.c_void_p  # device/host double* passed as integer addresses
c_ctx = C.c_void_p

# name -> (restype, argtypes); must list EVERY symbol of include/pymd_b200.h
PROTOTYPES = {
    "pymd_last_error": (C.c_char_p, []),
    "pymd_version": (C.c_int, []),
    "pymd_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(c_ctx)]),
    "pymd_ctx_destroy": (C.c_int, [c_ctx]),
    "pymd_set_dyn": (C.c_int, [c_ctx, c_dp]),
    "pymd_set_constraints": (C.c_int, [c_ctx, c_dp, C.c_int]),
    "pymd_bath_set": (C.c_int, [c_ctx, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_dp, C.c_int]),
    "pymd_bind_state": (C.c_int, [c_ctx, c_dp, c_dp]),
    "pymd_bath_ncp": (C.c_int, [C.c_int]),
    "pymd_ring_len": (C.c_int, [C.c_int, C.c_int]),
    "pymd_bind_history": (C.c_int, [c_ctx, C.c_int, c_dp, C.c_int]),
    "pymd_bind_noise": (C.c_int, [c_ctx, C.c_int, c_dp]),
    "pymd_bind_outputs": (C.c_int, [c_ctx, c_dp, c_dp, c_dp]),
    "pymd_bind_outputs_strided": (C.c_int, [c_ctx, c_dp, c_dp, c_dp, C.c_longlong]),
    "pymd_bind_noise_strided": (C.c_int, [c_ctx, C.c_int, c_dp, C.c_longlong]),
    "pymd_set_noise_wrap_row": (C.c_int, [c_ctx, C.c_int]),
    "pymd_bind_bath_outputs": (C.c_int, [c_ctx, C.c_int, c_dp, c_dp]),
    "pymd_reset_history": (C.c_int, [c_ctx, c_dp]),
    "pymd_invalidate_cache": (C.c_int, [c_ctx]),
    "pymd_history_get": (C.c_int, [c_ctx, C.c_int, c_dp, C.c_int, c_dp]),
    "pymd_history_set": (C.c_int, [c_ctx, C.c_int, c_dp, C.c_int, c_dp]),
    "pymd_step": (C.c_int, [c_ctx, C.c_longlong, C.c_int, C.c_int, c_dp]),
    "pymd_launch_count": (C.c_longlong, [c_ctx]),
    "pymd_noise_work_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "pymd_noise_generate": (C.c_int, [c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                      c_dp, c_dp, c_dp]),
    "pymd_noise_generate_chunk": (C.c_int, [c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                            c_dp, c_dp, c_dp]),
    "pymd_noise_generate_chunk_ld": (C.c_int, [c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                               c_dp, C.c_longlong, c_dp, c_dp]),
    "pymd_powerspec_ld": (C.c_int, [c_dp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_dp, c_dp,
                                    C.c_longlong, c_dp]),
    "pymd_profile_enable": (C.c_int, [c_ctx, C.c_int]),
    "pymd_profile_read": (C.c_int, [c_ctx, c_dp, c_dp]),
    "pymd_philox_normal": (C.c_int, [c_dp, C.c_longlong, C.c_int, C.c_uint64, C.c_uint32, C.c_longlong, c_dp]),
    "pymd_philox_uniform": (C.c_int, [c_dp, C.c_longlong, C.c_int, C.c_uint64, C.c_uint32, C.c_longlong, c_dp]),
    "pymd_init_state": (C.c_int, [c_ctx, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "pymd_powerspec_work_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_longlong]),
    "pymd_powerspec": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_dp, c_dp,
                                 C.c_longlong, c_dp]),
    "pymd_column_sums_work_bytes": (C.c_size_t, [C.c_int]),
    "pymd_column_sums": (C.c_int, [c_dp, C.c_longlong, C.c_int, C.c_double, c_dp, c_dp, c_dp]),
    "pymd_fft_columns": (C.c_int, [c_dp, C.c_int, C.c_longlong, C.c_int, c_dp, c_dp]),
    "pymd_fft_work_bytes": (C.c_size_t, [C.c_int, C.c_longlong]),
    "pymd_fft_num_passes": (C.c_int, [C.c_int]),
    "pymd_gemm": (C.c_int, [c_dp, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_dp]),
    "pymd_gamt": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int, C.c_double, C.c_double, c_dp, c_dp]),
    "pymd_harmonic_max_n": (C.c_int, []),
    "pymd_eigh_max_n": (C.c_int, [C.c_int]),
    "pymd_eigh_factors": (C.c_int, [c_dp, c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                    c_dp, c_dp]),
    "pymd_eigh_work_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "pymd_harmonic": (C.c_int, [c_dp, C.c_int, c_dp, c_dp, c_dp, C.c_int, C.c_double, c_dp, c_dp, c_dp]),
}


class PymdError(RuntimeError):
    pass


def load():
    """Load the shared library and declare every prototype.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PymdError("%s not found: build it with `python __graft_entry__.py` "
                        "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PymdError("libpymd_b200: %s (status %d)" % (load().pymd_last_error().decode(), rc))


def call(name, *args):
    """Call an int-returning entry point and raise PymdError on a non-zero status."""
    check(getattr(load(), name)(*args))


def ptr(t):
    """Address of a torch tensor (device or host) or numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
