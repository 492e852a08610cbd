#!/usr/bin/env python
"""Kernel-variant sweep on device-resident synthetic inputs (development tool, not a bench line).
Prints one JSON object per measurement to stdout."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS, _time_ms, w_alg_bytes  # noqa: E402
from matrixextra_b200 import _lib  # noqa: E402
from matrixextra_b200._lib import (MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_KEEP_F32, MXG_KEEP_F64,  # noqa: E402
                                   MXG_ROWS_CONTIGUOUS)
from matrixextra_b200.device import DeviceCSR  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--what", default="spmm32,spmm64,spmv,piece,transpose")
    args = ap.parse_args()
    what = set(args.what.split(","))
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    wl = WORKLOADS["cfg3"]
    m, K, n = wl["m"], wl["K"], 64

    def spmm_case(A, dtype, layout, n, tag, **opts):
        tdt = torch.float32 if dtype == MXG_F32 else torch.float64
        s = 4 if dtype == MXG_F32 else 8
        B = torch.randn(A.K, n, device="cuda", dtype=tdt, generator=g)
        out = torch.empty(A.m * n, device="cuda", dtype=tdt)
        for k, v in opts.items():
            _lib.set_option(k, v)
        try:
            ms = _time_ms(lambda: A.spmm(B, out, n, dtype, layout), args.steps, 3)
            emit(case=tag, n=n, dtype="f32" if dtype == MXG_F32 else "f64", layout=layout, opts=opts, ms=ms,
                 gflops=2.0 * A.nnz * n / ms / 1e6, eff_gbps=w_alg_bytes(A.m, A.K, A.nnz, n, s) / ms / 1e6)
        except Exception as e:  # noqa: BLE001
            emit(case=tag, opts=opts, error=str(e))
        for k in opts:
            _lib.set_option(k, 0)

    if what & {"spmm32", "spmm64", "piece"}:
        A = DeviceCSR.synth(m, K, wl["nnz"], 1, 1, seed=1003, keep=MXG_KEEP_F32 | MXG_KEEP_F64)
        emit(case="matrix", m=A.m, K=A.K, nnz=A.nnz, n_long=A.n_long, n_pieces=A.n_pieces, max_len=A.max_len)
        if "spmm32" in what:
            for rpw in (4, 8, 16):
                spmm_case(A, MXG_F32, MXG_ROWS_CONTIGUOUS, 64, "cfg3_f32_rm", spmm_rpw=rpw)
            spmm_case(A, MXG_F32, MXG_COLS_CONTIGUOUS, 64, "cfg3_f32_cm")
            for mb in (64, 128):
                spmm_case(A, MXG_F32, MXG_ROWS_CONTIGUOUS, 64, "cfg3_f32_rm_panels", spmm_panel_mb=mb)
            for nn in (8, 16, 32, 128, 256):
                spmm_case(A, MXG_F32, MXG_ROWS_CONTIGUOUS, nn, "f32_rm_n")
        if "spmm64" in what:
            for rpw in (4, 8, 16):
                spmm_case(A, MXG_F64, MXG_ROWS_CONTIGUOUS, 64, "k64_f64_rm", spmm_rpw=rpw)
            spmm_case(A, MXG_F64, MXG_ROWS_CONTIGUOUS, 64, "k64_f64_rm_lpr16", spmm_lpr=16)
            spmm_case(A, MXG_F64, MXG_COLS_CONTIGUOUS, 64, "k64_f64_cm")
            for nn in (8, 16, 32, 128):
                spmm_case(A, MXG_F64, MXG_ROWS_CONTIGUOUS, nn, "f64_rm_n")
        A.free()
        if "piece" in what:
            for piece in (256, 512, 2048, 4096, 65536):
                _lib.set_option("piece", piece)
                A = DeviceCSR.synth(m, K, wl["nnz"], 1, 1, seed=1003, keep=MXG_KEEP_F32)
                spmm_case(A, MXG_F32, MXG_ROWS_CONTIGUOUS, 64, f"cfg3_f32_rm_piece{piece}")
                A.free()
            _lib.set_option("piece", 1024)

    if "cpl" in what:
        # one vs two vectors per lane (half-width teams) across row widths, types and result layouts
        A = DeviceCSR.synth(m, K, wl["nnz"], 1, 1, seed=1003, keep=MXG_KEEP_F32 | MXG_KEEP_F64)
        for rep in range(2):
            for dt, nn in ((MXG_F32, 64), (MXG_F32, 128), (MXG_F32, 96), (MXG_F32, 40), (MXG_F32, 32), (MXG_F64, 32),
                           (MXG_F64, 64), (MXG_F64, 16)):
                for layout in (MXG_ROWS_CONTIGUOUS, MXG_COLS_CONTIGUOUS):
                    for cpl in (1, 2):
                        spmm_case(A, dt, layout, nn, "cpl", spmm_cpl=cpl)
        A.free()

    if "gatherroof" in what:
        # random-row gathers with nothing attached: the roof of the dense-operand gathers of K1/K2
        import ctypes as C
        sink = torch.zeros(4, device="cuda", dtype=torch.float32)
        for table_mb in (32, 256, 512, 2048, 8192):
            table = torch.randn(table_mb * (1 << 20) // 4, device="cuda", dtype=torch.float32)
            for row_bytes in (128, 256, 512):
                rows = table.numel() * 4 // row_bytes
                done = C.c_longlong(0)
                gathers = 200_000_000 * 256 // row_bytes // 2

                def probe():
                    _lib.call("mxg_dev_gather_probe", row_bytes, C.c_void_p(table.data_ptr()), rows, gathers, 12345,
                              C.c_void_p(sink.data_ptr()), C.byref(done), C.c_void_p(torch.cuda.current_stream().cuda_stream))
                ms = _time_ms(probe, args.steps, 3)
                emit(case="gather_probe", table_mb=table_mb, row_bytes=row_bytes, gathers=done.value, ms=ms,
                     gather_TBps=done.value * row_bytes / ms / 1e9, Ggathers_per_s=done.value / ms / 1e6)
            del table

    if "tmagather" in what:
        # the same random-row gathers through LDG.128 (mxg_dev_gather_probe) and through the TMA unit (tile::gather4)
        import ctypes as C
        sink = torch.zeros(4, device="cuda", dtype=torch.float32)
        for table_mb in (32, 256, 2048):
            table = torch.randn(table_mb * (1 << 20) // 4, device="cuda", dtype=torch.float32)
            for row_bytes in (128, 256, 512):
                rows = table.numel() * 4 // row_bytes
                gathers = 200_000_000 * 256 // row_bytes // 2
                for name in ("mxg_dev_gather_probe", "mxg_dev_tma_gather_probe"):
                    done = C.c_longlong(0)

                    def probe():
                        _lib.call(name, row_bytes, C.c_void_p(table.data_ptr()), rows, gathers, 12345,
                                  C.c_void_p(sink.data_ptr()), C.byref(done), C.c_void_p(torch.cuda.current_stream().cuda_stream))
                    try:
                        sink.zero_()
                        ms = _time_ms(probe, args.steps, 3)
                        emit(case="gather_ldg_vs_tma", path="TMA tile::gather4" if "tma" in name else "LDG.128", table_mb=table_mb,
                             row_bytes=row_bytes, gathers=done.value, ms=ms, gather_TBps=done.value * row_bytes / ms / 1e9,
                             timed_out=bool(sink[1].item() != 0))
                    except Exception as e:  # noqa: BLE001
                        emit(case="gather_ldg_vs_tma", path=name, table_mb=table_mb, row_bytes=row_bytes, error=str(e)[:300])
            del table

    if "l2res" in what:
        # the same product with a dense operand that fits L2: the ceiling column-panel tiling could reach
        for KK in (62_500, 125_000, 250_000, 500_000):
            A = DeviceCSR.synth(m, KK, wl["nnz"], 1, 1, seed=1003, keep=MXG_KEEP_F32 | MXG_KEEP_F64)
            spmm_case(A, MXG_F32, MXG_ROWS_CONTIGUOUS, 64, f"l2res_K{KK}_f32")
            spmm_case(A, MXG_F64, MXG_ROWS_CONTIGUOUS, 64, f"l2res_K{KK}_f64")
            spmm_case(A, MXG_F64, MXG_ROWS_CONTIGUOUS, 64, f"l2res_K{KK}_f64", spmm_lpr=16)
            A.free()

    if "spmv" in what:
        w2 = WORKLOADS["cfg2"]
        for piece in (1024, 4096):
            _lib.set_option("piece", piece)
            A = DeviceCSR.synth(w2["m"], w2["K"], w2["nnz"], 1, 0, seed=1002, keep=MXG_KEEP_F64)
            y = torch.randn(A.K, device="cuda", dtype=torch.float64, generator=g)
            o = torch.empty(A.m, device="cuda", dtype=torch.float64)
            for tex in (0, 1):
                _lib.set_option("spmv_tex", tex)
                for lpr in (8, 16, 32):
                    _lib.set_option("spmv_lpr", lpr)
                    ms = _time_ms(lambda: A.spmv(y, o), args.steps, 3)
                    emit(case="cfg2_spmv", piece=piece, lpr=lpr, tex=tex, ms=ms, gflops=2.0 * A.nnz / ms / 1e6,
                         eff_gbps=w_alg_bytes(A.m, A.K, A.nnz, 1, 8) / ms / 1e6)
            _lib.set_option("spmv_tex", 0)
            _lib.set_option("spmv_lpr", 0)
            A.free()
        _lib.set_option("piece", 1024)

    if "panels" in what:
        # column panels (k_spmm<PANELS>: one launch per L2-sized slab of the dense operand) vs none, on the cfg3 matrix
        A = DeviceCSR.synth(m, K, wl["nnz"], 1, 1, seed=1003, keep=MXG_KEEP_F32 | MXG_KEEP_F64)
        for rep in range(2):
            spmm_case(A, MXG_F32, MXG_ROWS_CONTIGUOUS, 64, "cfg3_f32_rm_nopanels")
            for mb in (40, 48, 56, 64, 86, 128):
                spmm_case(A, MXG_F32, MXG_ROWS_CONTIGUOUS, 64, "cfg3_f32_rm_panels", spmm_panel_mb=mb)
            spmm_case(A, MXG_F64, MXG_ROWS_CONTIGUOUS, 64, "k64_f64_rm_nopanels")
            for mb in (48, 56, 64, 86):
                spmm_case(A, MXG_F64, MXG_ROWS_CONTIGUOUS, 64, "k64_f64_rm_panels", spmm_panel_mb=mb)
        A.free()

    if "spmvprobe" in what:
        # K3's access pattern without the row structure (mxg_dev_spmv_probe): ids only / + 8-byte gathers of y
        # (plain loads, texture) / + values and FMA — the floor the SpMV kernel is measured against
        import ctypes as C
        w2 = WORKLOADS["cfg2"]
        A = DeviceCSR.synth(w2["m"], w2["K"], w2["nnz"], 1, 0, seed=1002, keep=MXG_KEEP_F64)
        y = torch.randn(A.K, device="cuda", dtype=torch.float64, generator=g)
        o = torch.empty(A.m, device="cuda", dtype=torch.float64)
        sink = torch.zeros(2, device="cuda", dtype=torch.float64)
        names = {0: "ids only (4 B/entry streamed)", 1: "ids + 8-byte gathers, plain loads", 2: "ids + 8-byte gathers, texture path",
                 3: "ids + values + texture gathers + FMA (12 B/entry streamed)"}
        for rep in range(2):
            for mode in (0, 1, 2, 3):
                def probe():
                    _lib.call("mxg_dev_spmv_probe", A._h, mode, C.c_void_p(y.data_ptr()), C.c_void_p(sink.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
                ms = _time_ms(probe, args.steps, 3)
                emit(case="spmv_probe", mode=mode, what=names[mode], nnz=A.nnz, ms=ms, Ggathers_per_s=A.nnz / ms / 1e6,
                     streamed_GBps=A.nnz * (12 if mode == 3 else 4) / ms / 1e6)
            ms = _time_ms(lambda: A.spmv(y, o), args.steps, 3)
            emit(case="spmv_probe", mode="k_spmv", what="the SpMV kernel itself (cfg2)", ms=ms,
                 eff_gbps=w_alg_bytes(A.m, A.K, A.nnz, 1, 8) / ms / 1e6)
        A.free()

    if "transpose" in what:
        w4 = WORKLOADS["cfg4"]
        A = DeviceCSR.synth(w4["m"], w4["K"], w4["nnz"], 1, 0, seed=1004, keep=MXG_KEEP_F64)

        def tr():
            t = A.transpose(keep=MXG_KEEP_F64)
            t.free()
        for rb in (5, 6, 7, 8, 9, 10, 7, 8, 10):
            _lib.set_option("radix_bits", rb)
            ms = _time_ms(tr, 5, 2)
            emit(case="cfg4_transpose", radix_bits=rb, ms=ms, alg_gbps=(24 * A.nnz + 4 * (A.m + A.K + 2)) / ms / 1e6)
        _lib.set_option("radix_bits", 8)
        A.free()
        w3 = WORKLOADS["cfg3"]
        A = DeviceCSR.synth(w3["m"], w3["K"], w3["nnz"], 1, 1, seed=1003, keep=MXG_KEEP_F32)
        for rb in (7, 8, 10):
            _lib.set_option("radix_bits", rb)
            ms = _time_ms(lambda: A.transpose(keep=MXG_KEEP_F32).free(), 5, 2)
            emit(case="cfg3_matrix_transpose_f32", radix_bits=rb, ms=ms)
        _lib.set_option("radix_bits", 8)
        A.free()

    # raw copy bandwidth on this box for context (same method as MEASURED_PEAKS.json)
    a = torch.empty(1 << 30, device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    ms = _time_ms(lambda: b.copy_(a), 10, 3)
    emit(case="copy_bw", gbps=2 * a.numel() * 2 / ms / 1e6)


if __name__ == "__main__":
    main()
