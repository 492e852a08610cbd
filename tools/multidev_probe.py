#!/usr/bin/env python
"""One level-1 call (cfg3: tcrossprod(float32 64 x 1M, CSR 2M x 1M)) spread over 1 .. N GPUs of the box by
mxg_set_devices(n) — wall clock per call, bytes moved, bit-equality with n = 1.  Development tool; bench.py's
`e2e.multi_device` record makes the same measurement on the driver's run."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from matrixextra_b200 import _lib, rcpp_exports as rx  # noqa: E402
from matrixextra_b200._lib import MXG_KEEP_F64  # noqa: E402
from matrixextra_b200.device import DeviceCSR  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--pageable", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    torch.cuda.set_device(0)
    A = DeviceCSR.synth(wl["m"], wl["K"], wl["nnz"], wl["row_model"], wl["col_model"], seed=wl["seed"], keep=MXG_KEEP_F64)
    p, j, x = A.to_host()
    A.free()
    n, K, m = wl["n"], wl["K"], wl["m"]
    f32 = wl["dtype"] == "f32"
    np_t = np.float32 if f32 else np.float64
    X = np.asfortranarray(np.random.default_rng(1).standard_normal((n, K)).astype(np_t))
    pin = (lambda a: a) if args.pageable else (lambda a: torch.from_numpy(a).pin_memory().numpy())
    p, j, x = pin(p), pin(j), pin(x)
    Xp = pin(np.ascontiguousarray(X.T)).T  # (n x K) F-order view
    out = None if args.pageable else torch.empty(n * m, dtype=torch.float32 if f32 else torch.float64).pin_memory().numpy().reshape((n, m), order="F")
    fn = rx.tcrossprod_dense_csr_float32 if f32 else rx.tcrossprod_dense_csr_numeric
    ndev = torch.cuda.device_count()
    want = None
    _lib.set_option("multi_pageable", 1 if args.pageable else 0)  # pageable calls stay on one device by default
    for share in (1, 0):
        _lib.set_option("multi_dense_share", share)
        for G in [g for g in (1, 2, 4, 8) if g <= ndev]:
            if G == 1 and share == 0:
                continue
            for narrow in (1, 0):
                _lib.set_option("host_narrow", narrow)
                _lib.call("mxg_set_devices", G)
                res = fn(Xp, p, j, x, 0, K, out=out)
                res = fn(Xp, p, j, x, 0, K, out=out)
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    res = fn(Xp, p, j, x, 0, K, out=out)
                dt = (time.perf_counter() - t0) / args.reps
                up, down = C.c_size_t(0), C.c_size_t(0)
                _lib.call("mxg_last_call_bytes", C.byref(up), C.byref(down))
                if want is None:
                    want = res.copy()
                print(json.dumps({"devices": G, "dense_share": share, "host_narrow": narrow, "pageable": args.pageable,
                                  "ms_per_call": dt * 1e3, "GFLOPs": 2.0 * int(p[-1]) * n / dt / 1e9, "h2d": up.value, "d2h": down.value,
                                  "bit_identical_to_one_device": bool(np.array_equal(res, want))}), flush=True)
    _lib.call("mxg_set_devices", 1)


if __name__ == "__main__":
    main()
