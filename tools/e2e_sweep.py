#!/usr/bin/env python
"""End-to-end (host buffers in, host buffer out) sweep of the streamed level-1 call on cfg3 / k64f64 / cfg2:
host narrowing on/off, host threads, ring slots, page-locked vs pageable operands.  Development tool, prints
one JSON object per measurement."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from matrixextra_b200 import _lib, rcpp_exports as rx  # noqa: E402
from matrixextra_b200._lib import MXG_KEEP_F64  # noqa: E402
from matrixextra_b200.device import DeviceCSR  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def timeit(f, k=4):
    f()
    t0 = time.perf_counter()
    for _ in range(k):
        f()
    return (time.perf_counter() - t0) / k * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="cfg3,k64f64,cfg2")
    ap.add_argument("--quick", action="store_true", help="default options only")
    ap.add_argument("--slots", action="store_true", help="only: ring slots 4 / 6 / 8 x packed ids off / on")
    ap.add_argument("--lag", action="store_true", help="only: packed ids off, then on with upload-lag thresholds 2 / 3 / 4")
    ap.add_argument("--pack", action="store_true", help="only: packed column ids on / off (option host_pack), twice each")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    emit(cpus=os.cpu_count())
    for name in args.workloads.split(","):
        wl = WORKLOADS[name]
        m, K, n = wl["m"], wl["K"], wl["n"]
        f32 = wl["dtype"] == "f32"
        A = DeviceCSR.synth(m, K, wl["nnz"], wl["row_model"], wl["col_model"], seed=wl["seed"], keep=MXG_KEEP_F64)
        p, j, x = A.to_host()
        nnz = A.nnz
        A.free()
        rng = np.random.default_rng(1)
        np_t = np.float32 if f32 else np.float64
        tt = torch.float32 if f32 else torch.float64
        if wl["op"] == "spmv":
            d = rng.standard_normal(K)
            flops = 2.0 * nnz

            def make(pp, jj, xx, dd, out):
                return lambda: rx.matmul_csr_dvec_numeric(pp, jj, xx, dd, 0, out=out)
            out_pin = torch.empty(m, dtype=torch.float64).pin_memory().numpy()
        else:
            d = np.asfortranarray(rng.standard_normal((n, K)).astype(np_t))
            flops = 2.0 * nnz * n
            if wl["op"] == "dense_tcsr":
                fn = rx.tcrossprod_dense_csr_float32 if f32 else rx.tcrossprod_dense_csr_numeric

                def make(pp, jj, xx, dd, out):
                    return lambda: fn(dd, pp, jj, xx, 0, K, out=out)
                out_pin = torch.empty(n * m, dtype=tt).pin_memory().numpy().reshape((n, m), order="F")
            else:
                fn = rx.tcrossprod_csr_dense_float32 if f32 else rx.tcrossprod_csr_dense_numeric

                def make(pp, jj, xx, dd, out):
                    return lambda: fn(pp, jj, xx, dd, 0, out=out)
                out_pin = torch.empty(n * m, dtype=tt).pin_memory().numpy().reshape((m, n), order="F")
        pinned = (pin(p), pin(j), pin(x))
        d_pin = torch.from_numpy(np.ascontiguousarray(d.T)).pin_memory().numpy().T if d.ndim == 2 else pin(d)
        cases = [("pinned", make(*pinned, d_pin, out_pin)), ("pageable_in_fresh_out", make(p, j, x, d, None)),
                 ("pinned_in_fresh_out", make(*pinned, d_pin, None))]

        def run(tag, **opts):
            for k, v in opts.items():
                _lib.set_option(k, v)
            for cname, f in cases:
                ms = timeit(f)
                emit(workload=name, case=cname, tag=tag, opts=opts, ms=ms, gflops=flops / ms / 1e6)

        if args.slots:
            for rep in range(2):
                for slots in (4, 6, 8):
                    for pack in (0, 1):
                        run("slots_x_pack", pipe_slots=slots, host_pack=pack)
            continue
        if args.lag:
            for rep in range(3):
                run("lag", host_pack=0)
                for lag in (2, 3, 4):
                    run("lag", host_pack=1, host_pack_lag=lag)
            continue
        if args.pack:
            for rep in range(2):
                run("pack_off", host_pack=0)
                run("pack_on", host_pack=1)
                run("pack_forced", host_pack=2)
            continue
        run("default")
        if args.quick:
            continue
        run("no_host_narrow", host_narrow=0)
        _lib.set_option("host_narrow", 1)
        run("no_staging", host_narrow=0, host_stage=0)
        _lib.set_option("host_narrow", 1)
        _lib.set_option("host_stage", 1)
        # rcpp_exports maps nthreads -> host_threads on every call, so sweep threads through the mirror's argument
        for t in (1, 2, 4, 8, 12):
            def with_threads(f, t=t):
                def g():
                    old = rx._threads
                    rx._threads = lambda _n: _lib.set_option("host_threads", t)
                    try:
                        return f()
                    finally:
                        rx._threads = old
                return g
            for cname, f in cases:
                ms = timeit(with_threads(f))
                emit(workload=name, case=cname, tag="threads", threads=t, ms=ms, gflops=flops / ms / 1e6)
        for slots in (3, 6, 8):
            run("slots", pipe_slots=slots)
        _lib.set_option("pipe_slots", 4)
        for div in (32, 64):
            run("chunks", pipe_chunk_nnz=max(nnz // div, 1 << 18))
        _lib.set_option("pipe_chunk_nnz", 0)


if __name__ == "__main__":
    main()
