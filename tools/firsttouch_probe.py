#!/usr/bin/env python
"""How fast can the host threads fill a FRESH pageable result (an R matrix that has just been allocated)?
Copies 512 MB into np.empty() buffers with the library's parallel host copy: untouched pages, after
madvise(MADV_HUGEPAGE), after MADV_POPULATE_WRITE, and into pages that already exist.  Development probe."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matrixextra_b200 import _lib  # noqa: E402

libc = C.CDLL("libc.so.6", use_errno=True)
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
N = 512 << 20
src = np.ones(N, dtype=np.uint8)
print(json.dumps({"thp_enabled": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
                  "thp_defrag": open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip(), "cpus": os.cpu_count()}))


def copy(dst):
    t0 = time.perf_counter()
    _lib.call("mxg_host_copy_2d", C.c_void_p(dst.ctypes.data), N, C.c_void_p(src.ctypes.data), N, N, 1, 0)
    return time.perf_counter() - t0


def huge_kb():
    tot = 0
    for ln in open("/proc/self/smaps"):
        if ln.startswith("AnonHugePages:"):
            tot += int(ln.split()[1])
    return tot


for rep in range(2):
    for mode in ("fresh", "madv_hugepage", "nohugepage", "nohugepage+8threads", "fresh+8threads", "fresh+4threads"):
        dst = np.empty(N, dtype=np.uint8)
        a = (dst.ctypes.data + (1 << 21) - 1) & ~((1 << 21) - 1)
        ln = ((dst.ctypes.data + N) & ~((1 << 21) - 1)) - a
        t_adv = 0.0
        t0 = time.perf_counter()
        rc = 0
        if mode in ("madv_hugepage", "hugepage+populate"):
            rc = libc.madvise(a, ln, 14)
        if mode.startswith("nohugepage"):
            rc = libc.madvise(a, ln, 15)
        _lib.set_option("host_threads", 8 if "8threads" in mode else (4 if "4threads" in mode else 0))
        t_adv = time.perf_counter() - t0
        h0 = huge_kb()
        t1 = copy(dst)
        t2 = copy(dst)
        print(json.dumps({"mode": mode, "madvise_rc": rc, "madvise_ms": t_adv * 1e3, "first_copy_ms": t1 * 1e3,
                          "first_copy_GBps": N / t1 / 1e9, "second_copy_ms": t2 * 1e3, "second_copy_GBps": N / t2 / 1e9,
                          "anon_huge_MB_before_copy": h0 // 1024, "anon_huge_MB_after": huge_kb() // 1024}), flush=True)
        del dst
