#!/usr/bin/env python
"""Host probe: cost of first-touching a fresh 512 MB result buffer — page faults by copying, madvise(MADV_POPULATE_WRITE)
from N threads, then the copy.  Development aid for csrc/hoststage.cu."""
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from matrixextra_b200 import _lib  # noqa: E402

libc = C.CDLL("libc.so.6", use_errno=True)
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
MADV_POPULATE_WRITE = 23
N = 512_000_000
src = np.random.default_rng(0).integers(0, 255, N, dtype=np.uint8)


def vp(a):
    return C.c_void_p(a.ctypes.data)


def populate(buf, threads):
    a0 = (buf.ctypes.data + 4095) & ~4095
    a1 = (buf.ctypes.data + buf.nbytes) & ~4095
    step = ((a1 - a0) // threads + 4095) & ~4095
    errs = []

    def work(i):
        lo = a0 + i * step
        hi = min(a1, lo + step)
        if hi > lo and libc.madvise(lo, hi - lo, MADV_POPULATE_WRITE) != 0:
            errs.append(C.get_errno())
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return (time.perf_counter() - t0) * 1e3, errs


for threads in (1, 4, 8, 16):
    _lib.set_option("host_threads", threads)
    out = np.empty(N, dtype=np.uint8)
    t0 = time.perf_counter()
    _lib.call("mxg_host_copy_2d", vp(out), N, vp(src), N, N, 1, 0)
    t_fault_copy = (time.perf_counter() - t0) * 1e3
    del out
    out = np.empty(N, dtype=np.uint8)
    t_pop, errs = populate(out, threads)
    t0 = time.perf_counter()
    _lib.call("mxg_host_copy_2d", vp(out), N, vp(src), N, N, 1, 0)
    t_copy = (time.perf_counter() - t0) * 1e3
    del out
    print(json.dumps(dict(threads=threads, copy_into_fresh_ms=t_fault_copy, populate_ms=t_pop, copy_after_populate_ms=t_copy,
                          errno=errs[:1])), flush=True)
