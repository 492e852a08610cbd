"""ctypes binding of the C ABI in include/mxgpu.h (matrixextra_b200/csrc/libmxgpu.so).

There is deliberately no fallback: if the CUDA library is missing or a call fails, an exception is
raised (``MxgError``) — the product path never routes through a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmxgpu.so")

MXG_OK, MXG_ERR_CUDA, MXG_ERR_ARG, MXG_ERR_INDEX, MXG_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
MXG_F64, MXG_F32 = 0, 1
MXG_ROWS_CONTIGUOUS, MXG_COLS_CONTIGUOUS = 0, 1
MXG_Y_NUMERIC, MXG_Y_INTEGER, MXG_Y_LOGICAL, MXG_Y_FLOAT32, MXG_Y_BINARY = 0, 1, 2, 3, 4
MXG_KEEP_F64, MXG_KEEP_F32 = 1, 2


class MxgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"mxgpu error {code}: {message}")
        self.code = code
        self.message = message


_vp, _i32, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    "mxg_device_count": [C.POINTER(C.c_int)],
    "mxg_set_device": [_i32],
    "mxg_set_devices": [_i32],
    "mxg_get_devices": [C.POINTER(C.c_int)],
    "mxg_cache_clear": [],
    "mxg_cache_stats": [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(_sz), C.POINTER(C.c_int)],
    "mxg_csr_spmm_host": [_vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp, _sz],
    "mxg_csr_spmm_t_host": [_vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp, _sz],
    "mxg_csr_spmv_host": [_vp, _i32, _vp, _vp],
    "mxg_set_option": [C.c_char_p, C.c_long],
    "mxg_get_option": [C.c_char_p, C.POINTER(C.c_long)],
    "mxg_trim": [],
    "mxg_spmm_csr_dense": [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _sz],
    "mxg_spmv_csr": [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "mxg_spmv_csr_svec": [_i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
    "mxg_dev_spmv_svec": [_vp, _i32, _i32, _vp, _vp, _vp, _vp],
    "mxg_check_valid_csr": [_i32, _i32, _vp, _vp, _i64, C.POINTER(C.c_int)],
    "mxg_rows_sorted": [_i32, _vp, _vp, C.POINTER(C.c_int)],
    "mxg_sort_csr_indices": [_i32, _vp, _vp, _vp],
    "mxg_mul_csr_dense": [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "mxg_mul_csr_dvec": [_i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp],
    "mxg_dev_check_valid_csr": [_i32, _i32, _vp, _vp, _i64, C.POINTER(C.c_int), _vp],
    "mxg_dev_rows_sorted": [_i32, _vp, _vp, C.POINTER(C.c_int), _vp],
    "mxg_dev_sort_csr_indices": [_i32, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int), _vp],
    "mxg_dev_mul_csr_dense": [_vp, _i32, _vp, _vp, _vp],
    "mxg_dev_mul_csr_dvec": [_vp, _vp, _sz, _vp, _vp],
    "mxg_csr2csc": [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "mxg_spmm_csrT_dense": [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _sz],
    "mxg_csr_upload": [_i32, _i32, _vp, _vp, _vp, _i32, C.POINTER(_vp)],
    "mxg_csr_wrap_device": [_i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, C.POINTER(_vp)],
    "mxg_csr_free": [_vp],
    "mxg_csr_info": [_vp, C.POINTER(_i64)],
    "mxg_csr_device_arrays": [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "mxg_csr_download": [_vp, _vp, _vp, _vp],
    "mxg_dev_spmm": [_vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp, _sz, _vp],
    "mxg_dev_spmv": [_vp, _i32, _vp, _vp, _vp],
    "mxg_dev_spmm_bcast": [_vp, _i32, _i32, _i32, _i32, _vp, _sz, _i32, _vp, _sz, _vp],
    "mxg_dev_spmm_rows": [_vp, _i32, _i32, _i32, _vp, _sz, _vp, _sz, _i32, _i32, _i32, _vp],
    "mxg_dev_copy_2d": [_vp, _sz, _vp, _sz, _sz, _sz, _vp],
    "mxg_dev_spmm_push": [_vp, _i32, _i32, _i32, _i32, _vp, _sz, _i32, _vp, _sz, _vp],
    "mxg_dev_spmv_bcast": [_vp, _i32, _vp, _i32, _vp, _vp],
    "mxg_dev_spmm_mcast": [_vp, _i32, _i32, _vp, _sz, _vp, _sz, _vp],
    "mxg_dev_alloc": [_sz, C.POINTER(_vp)],
    "mxg_dev_free": [_vp],
    "mxg_ipc_export": [_vp, _vp],
    "mxg_ipc_open": [_vp, C.POINTER(_vp)],
    "mxg_ipc_close": [_vp],
    "mxg_dev_peer_barrier": [_i32, _i32, _vp, _i32, _vp],
    "mxg_dev_barrier_failed": [C.POINTER(C.c_int)],
    "mxg_dev_csr2csc": [_vp, _i32, _vp, C.POINTER(_vp)],
    "mxg_dev_transpose_dense": [_i32, _sz, _sz, _vp, _sz, _vp, _sz, _vp],
    "mxg_row_partition": [_i32, _vp, _i32, _vp],
    "mxg_dev_gather_probe": [_i32, _vp, _sz, C.c_longlong, C.c_uint64, _vp, C.POINTER(C.c_longlong), _vp],
    "mxg_host_alloc": [_sz, C.POINTER(_vp)],
    "mxg_host_free": [_vp],
    "mxg_host_pool_stats": [C.POINTER(_sz), C.POINTER(_sz), C.POINTER(C.c_int)],
    "mxg_host_narrow": [_vp, _vp, _sz],
    "mxg_host_copy_2d": [_vp, _sz, _vp, _sz, _sz, _sz, _i32],
    "mxg_host_pack_indices": [_vp, _sz, _i32, _vp, C.POINTER(_sz), C.POINTER(C.c_int)],
    "mxg_last_call_bytes": [C.POINTER(_sz), C.POINTER(_sz)],
    "mxg_host_chunk_plan": [_i32, _vp, _sz, _vp, _i32] + [C.POINTER(C.c_int)] * 4,
    "mxg_dev_tma_gather_probe": [_i32, _vp, _sz, C.c_longlong, C.c_uint64, _vp, C.POINTER(C.c_longlong), _vp],
    "mxg_dev_spmv_probe": [_vp, _i32, _vp, _vp, _vp],
    "mxg_synth_csr": [_i32, _i32, _i64, _i32, _i32, C.c_uint64, _i32, _vp, C.POINTER(_vp)],
}
_RESTYPES = {"mxg_last_error": C.c_char_p, "mxg_launch_count": C.c_ulonglong, "mxg_csr_error_string": C.c_char_p}

_lib = None


def load() -> C.CDLL:
    """Load libmxgpu.so (once).  Raises if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m matrixextra_b200.build_native` "
            "(matrixextra_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.mxg_last_error.argtypes = []
    lib.mxg_last_error.restype = C.c_char_p
    lib.mxg_launch_count.argtypes = []
    lib.mxg_launch_count.restype = C.c_ulonglong
    lib.mxg_csr_error_string.argtypes = [C.c_int]
    lib.mxg_csr_error_string.restype = C.c_char_p
    _lib = lib
    return lib


def exported_names():
    return sorted(list(_SIGNATURES) + list(_RESTYPES))


def check(rc: int) -> None:
    if rc != MXG_OK:
        msg = load().mxg_last_error()
        raise MxgError(rc, msg.decode("utf-8", "replace") if msg else "")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))


def launch_count() -> int:
    return int(load().mxg_launch_count())


def set_option(name: str, value: int) -> None:
    call("mxg_set_option", name.encode(), int(value))


def get_option(name: str) -> int:
    out = C.c_long()
    call("mxg_get_option", name.encode(), C.byref(out))
    return int(out.value)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().mxg_device_count(C.byref(n))
    return int(n.value) if rc == MXG_OK else 0
