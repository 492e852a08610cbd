"""Device-resident handles (level 2 of the C ABI): what bench.py, the sharded path and repeated
multiplies use.  torch is only plumbing here (device memory, streams); every kernel is ours."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_KEEP_F32, MXG_KEEP_F64, MXG_ROWS_CONTIGUOUS,
                   MXG_Y_FLOAT32, MXG_Y_INTEGER, MXG_Y_LOGICAL, MXG_Y_NUMERIC)


def _stream_ptr(stream=None) -> C.c_void_p:
    if stream is None:
        import torch
        stream = torch.cuda.current_stream()
    return C.c_void_p(int(getattr(stream, "cuda_stream", stream) or 0))


def _dptr(t) -> C.c_void_p:
    return C.c_void_p(int(t.data_ptr()) if t is not None else 0)


class DeviceCSR:
    """Owns (or wraps) a CSR matrix in HBM: int32 indptr/indices, values in float64 and/or float32,
    plus the row statistics and long-row piece tables the kernels consume."""

    def __init__(self, handle: int, keepalive=None):
        self._h = C.c_void_p(handle)
        self._keepalive = keepalive
        info = (C.c_int64 * 6)()
        _lib.call("mxg_csr_info", self._h, info)
        self.m, self.K, self.nnz, self.n_long, self.n_pieces, self.max_len = (int(v) for v in info)

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def upload(cls, m, K, p, j, x, keep=MXG_KEEP_F64 | MXG_KEEP_F32):
        p = np.ascontiguousarray(p, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        h = C.c_void_p()
        _lib.call("mxg_csr_upload", int(m), int(K), C.c_void_p(p.ctypes.data), C.c_void_p(j.ctypes.data),
                  C.c_void_p(x.ctypes.data), int(keep), C.byref(h))
        return cls(h.value)

    @classmethod
    def synth(cls, m, K, nnz, row_model=1, col_model=0, seed=1000, keep=MXG_KEEP_F64 | MXG_KEEP_F32, stream=None):
        h = C.c_void_p()
        _lib.call("mxg_synth_csr", int(m), int(K), int(nnz), int(row_model), int(col_model), int(seed), int(keep),
                  _stream_ptr(stream), C.byref(h))
        return cls(h.value)

    @classmethod
    def wrap(cls, m, K, p_t, j_t, x64_t=None, x32_t=None, validate=True, stream=None):
        """Wrap torch CUDA tensors (int32, int32, float64 / float32) without copying."""
        h = C.c_void_p()
        _lib.call("mxg_csr_wrap_device", int(m), int(K), _dptr(p_t), _dptr(j_t), _dptr(x64_t), _dptr(x32_t),
                  1 if validate else 0, _stream_ptr(stream), C.byref(h))
        return cls(h.value, keepalive=(p_t, j_t, x64_t, x32_t))

    def free(self):
        if self._h and self._h.value:
            _lib.call("mxg_csr_free", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- raw device arrays ------------------------------------------------------------------------
    def device_arrays(self):
        ptrs = [C.c_void_p() for _ in range(4)]
        _lib.call("mxg_csr_device_arrays", self._h, *[C.byref(q) for q in ptrs])
        return tuple(q.value for q in ptrs)

    def to_host(self):
        """Copy indptr / indices / float64 values back (tests, CPU-baseline sampling)."""
        p = np.empty(self.m + 1, dtype=np.int32)
        j = np.empty(self.nnz, dtype=np.int32)
        x = np.empty(self.nnz, dtype=np.float64)
        _lib.call("mxg_csr_download", self._h, C.c_void_p(p.ctypes.data), C.c_void_p(j.ctypes.data),
                  C.c_void_p(x.ctypes.data))
        return p, j, x

    # -- products -----------------------------------------------------------------------------------
    def spmm(self, B_t, out_t, n, dtype, out_layout=MXG_ROWS_CONTIGUOUS, b_layout=MXG_ROWS_CONTIGUOUS,
             ldb=None, ldc=None, stream=None):
        """Out = A . B on torch CUDA tensors; asynchronous on the (current) stream."""
        if ldb is None:
            ldb = n if b_layout == MXG_ROWS_CONTIGUOUS else self.K
        if ldc is None:
            ldc = n if out_layout == MXG_ROWS_CONTIGUOUS else self.m
        _lib.call("mxg_dev_spmm", self._h, int(dtype), int(out_layout), int(b_layout), int(n), _dptr(B_t), int(ldb),
                  _dptr(out_t), int(ldc), _stream_ptr(stream))

    def spmm_bcast(self, B_t, dst_ptrs, n, dtype, out_layout=MXG_ROWS_CONTIGUOUS, ldb=None, ldc=None, stream=None):
        """The same product with every finished row stored into ALL ``dst_ptrs`` (raw device addresses: the local
        result first, then the peer-mapped results of the other GPUs) — compute and all-gather in one kernel."""
        if ldb is None:
            ldb = n
        if ldc is None:
            ldc = n if out_layout == MXG_ROWS_CONTIGUOUS else self.m
        arr = (C.c_void_p * len(dst_ptrs))(*[int(q) for q in dst_ptrs])
        _lib.call("mxg_dev_spmm_bcast", self._h, int(dtype), int(out_layout), MXG_ROWS_CONTIGUOUS, int(n), _dptr(B_t),
                  int(ldb), len(dst_ptrs), arr, int(ldc), _stream_ptr(stream))

    def spmm_rows(self, B_t, out_ptr, n, dtype, out_layout, ldc, r0=0, r1=0, pieces=False, ldb=None, stream=None):
        """Rows [r0, r1) of the product without their long rows, or (``pieces``) only the long rows; ``out_ptr`` is the
        raw device address of the FULL result's origin."""
        _lib.call("mxg_dev_spmm_rows", self._h, int(dtype), int(out_layout), int(n), _dptr(B_t), int(ldb or n),
                  C.c_void_p(int(out_ptr)), int(ldc), int(r0), int(r1), 1 if pieces else 0, _stream_ptr(stream))

    def spmm_push(self, B_t, dst_ptrs, n, dtype, out_layout=MXG_ROWS_CONTIGUOUS, ldb=None, ldc=None, stream=None):
        """Product in row slices into ``dst_ptrs[0]``; every finished slice is pushed to the other destinations by the
        copy engines (NVLink peer copies) while the next slice is computed."""
        if ldb is None:
            ldb = n
        if ldc is None:
            ldc = n if out_layout == MXG_ROWS_CONTIGUOUS else self.m
        arr = (C.c_void_p * len(dst_ptrs))(*[int(q) for q in dst_ptrs])
        _lib.call("mxg_dev_spmm_push", self._h, int(dtype), int(out_layout), MXG_ROWS_CONTIGUOUS, int(n), _dptr(B_t),
                  int(ldb), len(dst_ptrs), arr, int(ldc), _stream_ptr(stream))

    def spmm_mcast(self, B_t, mc_ptr, n, dtype, ldb=None, ldc=None, stream=None):
        """Fused product + all-gather through NVLS multicast: ``mc_ptr`` is the raw multicast address of this block's
        first row (rows-contiguous); every row is stored once and replicated by the switch into all GPUs' results."""
        _lib.call("mxg_dev_spmm_mcast", self._h, int(dtype), int(n), _dptr(B_t), int(ldb or n), C.c_void_p(int(mc_ptr)),
                  int(ldc or n), _stream_ptr(stream))

    def spmv_bcast(self, y_t, dst_ptrs, ytype=MXG_Y_NUMERIC, stream=None):
        arr = (C.c_void_p * len(dst_ptrs))(*[int(q) for q in dst_ptrs])
        _lib.call("mxg_dev_spmv_bcast", self._h, int(ytype), _dptr(y_t), len(dst_ptrs), arr, _stream_ptr(stream))

    def spmv(self, y_t, out_t, ytype=MXG_Y_NUMERIC, stream=None):
        _lib.call("mxg_dev_spmv", self._h, int(ytype), _dptr(y_t), _dptr(out_t), _stream_ptr(stream))

    def spmv_svec(self, yidx_base1_t, yvals_t, out_t, ytype=MXG_Y_NUMERIC, stream=None):
        """out (float64) = A . sparse vector given by int32 1-based positions and values (device tensors)."""
        _lib.call("mxg_dev_spmv_svec", self._h, int(ytype), int(yidx_base1_t.numel()), _dptr(yidx_base1_t), _dptr(yvals_t),
                  _dptr(out_t), _stream_ptr(stream))

    def transpose(self, keep=MXG_KEEP_F64 | MXG_KEEP_F32, stream=None) -> "DeviceCSR":
        """Deep CSR -> CSC on device, returned as the CSR handle of t(A)."""
        h = C.c_void_p()
        _lib.call("mxg_dev_csr2csc", self._h, int(keep), _stream_ptr(stream), C.byref(h))
        return DeviceCSR(h.value)


def row_partition(p_host: np.ndarray, parts: int) -> np.ndarray:
    """nnz-balanced contiguous row blocks (host indptr) — the shard boundaries of SURVEY.md §8 e."""
    p = np.ascontiguousarray(p_host, dtype=np.int32)
    out = np.empty(parts + 1, dtype=np.int32)
    _lib.call("mxg_row_partition", int(p.size - 1), C.c_void_p(p.ctypes.data), int(parts), C.c_void_p(out.ctypes.data))
    return out
