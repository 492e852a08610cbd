#!/usr/bin/env python
"""Host-side roofs of the staging engine on the GPU box (no kernels): narrowing and copy bandwidth by thread
count, cost of first-touching a fresh result, effect of MADV_HUGEPAGE.  One JSON object per measurement."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from matrixextra_b200 import _lib  # noqa: E402

libc = C.CDLL("libc.so.6", use_errno=True)
MADV_HUGEPAGE = 14


def vp(a):
    return C.c_void_p(a.ctypes.data)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def best(f, k=3):
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def main():
    for f in ("/sys/kernel/mm/transparent_hugepage/enabled", "/sys/kernel/mm/transparent_hugepage/defrag"):
        try:
            emit(file=f, value=open(f).read().strip())
        except OSError as e:
            emit(file=f, error=str(e))
    try:
        emit(cpu=[l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0], cpus=os.cpu_count())
    except Exception:  # noqa: BLE001
        pass
    n = 100_000_000
    x = np.random.default_rng(0).standard_normal(n)
    dst = np.zeros(n, dtype=np.float32)
    src8 = x.view(np.uint8)
    dst8 = np.zeros(512_000_000, dtype=np.uint8)
    for t in (1, 2, 4, 8, 12, 16):
        _lib.set_option("host_threads", t)
        ms = best(lambda: _lib.call("mxg_host_narrow", vp(x), vp(dst), n))
        emit(what="narrow 100M f64->f32 (touched dst)", threads=t, ms=ms, GBps_read_plus_write=1.2e9 / ms / 1e6)
        ms = best(lambda: _lib.call("mxg_host_copy_2d", vp(dst8), 256, vp(src8), 256, 256, 2_000_000, 0))
        emit(what="copy 512 MB (touched dst)", threads=t, ms=ms, GBps_read_plus_write=1.024e9 / ms / 1e6)

        def fresh(advise):
            out = np.empty(512_000_000 + 4096, dtype=np.uint8)
            if advise:
                a = (out.ctypes.data + (1 << 21) - 1) & ~((1 << 21) - 1)
                ln = (out.ctypes.data + out.nbytes - a) & ~((1 << 21) - 1)
                rc = libc.madvise(C.c_void_p(a), C.c_size_t(ln), MADV_HUGEPAGE)
                if rc != 0:
                    emit(madvise_errno=C.get_errno())
            t0 = time.perf_counter()
            _lib.call("mxg_host_copy_2d", vp(out), 256, vp(src8), 256, 256, 2_000_000, 0)
            dt = time.perf_counter() - t0
            del out
            return dt * 1e3
        emit(what="copy 512 MB into a FRESH buffer", threads=t, ms=min(fresh(False) for _ in range(3)))
        emit(what="copy 512 MB into a FRESH buffer + MADV_HUGEPAGE", threads=t, ms=min(fresh(True) for _ in range(3)))
    _lib.set_option("host_threads", 0)


if __name__ == "__main__":
    main()
