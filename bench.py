#!/usr/bin/env python
"""bench.py — throughput of the sparse x dense multiplication path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over the workload's synthetic
inputs (generated on device by the library's counter-based generator, SURVEY.md §8 d):

  value     whole-job GFLOP/s (2*nnz*n flops per product) with every operand resident in HBM, timed with
            CUDA events on the launching stream, max over ranks;
  e2e       the same metric through the reference-facing entry point (the Rcpp-export mirror on the
            level-1 C ABI) with HOST buffers: H2D of the CSR + dense operand and D2H of the result are
            inside the timed region;
  roofline  algorithmic bytes W_alg (every operand byte once) / step time, against the measured HBM peak;
  cpu_baseline  the reference's own src/matmul.cpp (oracle/_ref, OpenMP, all host cores) on a bounded
            row sample of the same matrix.

Multi-GPU (torchrun, one rank per GPU): weak scaling — every rank owns one row block of the workload's
size (global matrix = N blocks stacked), the dense operand is replicated, and the step ends with the
NCCL all-gather of the output row blocks (north_star subsystem 4).  `compute_only` reports the step
without the collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (m, K, nnz, row_model, col_model, n, dtype, op, seed)
    "cfg1": dict(m=10_000, K=5_000, nnz=500_000, row_model=0, col_model=0, n=32, dtype="f64", op="csr_dense", seed=1001,
                 desc="cfg1: dgRMatrix 10k x 5k (1% dense) %*% dense 5k x 32 fp64 [tcrossprod_csr_dense_numeric]"),
    "cfg2": dict(m=2_000_000, K=1_000_000, nnz=100_000_000, row_model=1, col_model=0, n=1, dtype="f64", op="spmv", seed=1002,
                 desc="cfg2: SpMV CSR 2M x 1M, 100M nnz power-law rows %*% dense vector fp64 [matmul_csr_dvec_numeric]"),
    "cfg3": dict(m=2_000_000, K=1_000_000, nnz=100_000_000, row_model=1, col_model=1, n=64, dtype="f32", op="dense_tcsr", seed=1003,
                 desc="cfg3: tcrossprod(dense 64 x 1M float32, CSR 2M x 1M, 100M nnz power-law) k=64 fp32 [tcrossprod_dense_csr_float32]"),
    "k64f64": dict(m=2_000_000, K=1_000_000, nnz=100_000_000, row_model=1, col_model=1, n=64, dtype="f64", op="csr_dense", seed=1003,
                   desc="k64f64: CSR 2M x 1M (100M nnz power-law) %*% dense 1M x 64 fp64 [tcrossprod_csr_dense_numeric]"),
    "cfg4": dict(m=5_000_000, K=500_000, nnz=100_000_000, row_model=1, col_model=0, n=64, dtype="f64", op="crossprod", seed=1004,
                 desc="cfg4: crossprod(CSR 5M x 500k, 100M nnz, dense 5M x 64) fp64: device CSR->CSC + gather product"),
    "cfg5": dict(m=50_000_000, K=10_000_000, nnz=2_000_000_000, row_model=1, col_model=0, n=128, dtype="f32", op="csr_dense", seed=1005,
                 desc="cfg5: CSR 50M x 10M, 2B nnz Zipf rows %*% dense 10M x 128 fp32 [tcrossprod_csr_dense_float32]"),
}


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch, from the committed `ncu --set full`
# captures of exactly these workloads (profiles/README.md).  Every entry is tied to the kernel it was captured from
# (mangled instantiation) and to the SHA-256 of the source file that defines it: when the source has changed since the
# capture, `roofline.traffic` is reported as null instead of a stale number (tools/ncu_summary.py prints the entry to
# paste here after a new capture).
NCU_TRAFFIC = {
    "cfg3": dict(bytes=15.616967e9 + 0.565772e9, kernel="k_spmm<float, 4, 8, 2, 4, 0, 0, 6, 0, 0>", src="spmm.cu",
                 src_sha16="67d601880a394fe5", capture="profiles/r02_v5_spmm_f32_k64_cfg3.ncu.txt"),
    "k64f64": dict(bytes=39.134754e9 + 1.028319e9, kernel="k_spmm<double, 2, 16, 2, 4, 1, 0, 5, 0, 0>", src="spmm.cu",
                   src_sha16="67d601880a394fe5", capture="profiles/r02_v5_spmm_f64_k64.ncu.txt"),
    "cfg2": dict(bytes=1.220682e9 + 0.018391e9, kernel="k_spmv<0, double, 16>", src="spmv.cu",
                 src_sha16="b7a13919e41675af", capture="profiles/r02_v5_spmv_f64_cfg2.ncu.txt"),
}


def source_sha16(name):
    import hashlib
    path = os.path.join(ROOT, "matrixextra_b200", "csrc", name)
    try:
        with open(path, "rb") as fh:
            return hashlib.sha256(fh.read()).hexdigest()[:16]
    except OSError:
        return None


def ncu_traffic(workload):
    """(bytes or None, provenance string) for roofline.traffic."""
    e = NCU_TRAFFIC.get(workload)
    if e is None:
        return None, None
    now = source_sha16(e["src"])
    if not e["src_sha16"] or now != e["src_sha16"]:
        return None, (f"{e['capture']} ({e['kernel']}) was captured from another version of {e['src']} "
                      f"(sha256 {e['src_sha16'] or 'unrecorded'} then, {now} now): no current figure")
    return e["bytes"], f"{e['capture']}: {e['kernel']}, {e['src']} sha256 {now}"


def bench_config(wl, world, nnz):
    """`config` of the JSON line — identical in the GPU arm and the reference arm (same workload, same partition)."""
    s = 4 if wl["dtype"] == "f32" else 8
    out_rows = wl["K"] if wl["op"] == "csrT_dense" else wl["m"]
    step_bytes = int(nnz) * (4 + s) + s * wl["K"] * wl["n"] + s * out_rows * wl["n"]
    return {"workload": wl["desc"], "rows_per_gpu": wl["m"], "cols": wl["K"], "nnz_per_gpu": int(nnz), "n": wl["n"],
            "parallelism": (f"row-block shards x{world}, replicated dense operand, all-gather of the output row blocks"
                            if world > 1 else "single GPU"),
            # timing rule: flush L2 between timed steps, or use inputs larger than L2 — which one this workload is
            "l2_between_steps": (f"no flush: the operands of one step ({step_bytes / 1e6:.0f} MB) exceed the 126 MB L2 several times over"
                                 if step_bytes > 2 * 126e6 else
                                 f"NOT flushed: the operands of one step ({step_bytes / 1e6:.0f} MB) fit the 126 MB L2 — a warm-cache parity config, not a bench line")}


def w_alg_bytes(m, K, nnz, n, s):
    """compulsory traffic, every operand byte once; values stored in the compute type."""
    return nnz * (4 + s) + 4 * (m + 1) + s * K * n + s * m * n


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [q.strip() for q in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# --------------------------------------------------------------------------------------------------
def cpu_run(cpu, wl, p, j, x, dense, nthreads):
    """One pass of the reference's CPU path on host arrays; returns (seconds, result or None)."""
    t0 = time.perf_counter()
    if wl["op"] == "dense_tcsr":
        fn = cpu.tcrossprod_dense_csr_float32 if wl["dtype"] == "f32" else cpu.tcrossprod_dense_csr_numeric
        res = fn(dense, p, j, x, nthreads, wl["K"])
    elif wl["op"] == "csr_dense":
        fn = cpu.tcrossprod_csr_dense_float32 if wl["dtype"] == "f32" else cpu.tcrossprod_csr_dense_numeric
        res = fn(p, j, x, dense, nthreads)
    elif wl["op"] == "spmv":
        res = cpu.matmul_csr_dvec_numeric(p, j, x, dense, nthreads)
    elif wl["op"] == "crossprod":
        # the all-MatrixExtra CPU route: stable CSR->CSC, then matmul_dense_csc on t(Y)
        from oracle.cpu_oracle import Port
        p2, i2, x2 = Port().csr2csc(p.size - 1, wl["K"], p, j, x)
        res = cpu.matmul_dense_csc_numeric(dense, p2, i2, x2, nthreads)
    else:
        raise ValueError(wl["op"])
    return time.perf_counter() - t0, res


def cpu_dense_operand(wl, rows, rng):
    np_t = np.float32 if wl["dtype"] == "f32" else np.float64
    if wl["op"] in ("dense_tcsr", "csr_dense"):
        return np.asfortranarray(rng.standard_normal((wl["n"], wl["K"])).astype(np_t))  # (n x K) column-major
    if wl["op"] == "spmv":
        return rng.standard_normal(wl["K"])
    if wl["op"] == "crossprod":
        return np.asfortranarray(rng.standard_normal((wl["n"], rows)).astype(np_t))  # t(Y): n x m
    raise ValueError(wl["op"])


def host_cores():
    """Host threads the reference may use: every CPU this process is allowed on.  Not omp_get_max_threads():
    torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would time a single-threaded reference."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def gpu_level1(rx, wl, p, j, x, dense, nth=0, out=None):
    """The same product through the reference-facing entry point (Rcpp-export mirror -> level-1 C ABI), host buffers."""
    from matrixextra_b200._lib import MXG_F32, MXG_F64
    f32 = wl["dtype"] == "f32"
    op = wl["op"]
    if op == "spmv":
        return rx.matmul_csr_dvec_numeric(p, j, x, dense, nth, out=out)
    if op == "crossprod":  # dense is t(Y) (n x m) as the CPU route takes it; the export takes Y (m x n) column-major
        return rx.crossprod_csr_dense(p, j, x, wl["K"], np.asfortranarray(dense.T), MXG_F32 if f32 else MXG_F64).T
    if op == "dense_tcsr":
        fn = rx.tcrossprod_dense_csr_float32 if f32 else rx.tcrossprod_dense_csr_numeric
        return fn(dense, p, j, x, nth, wl["K"], out=out)
    fn = rx.tcrossprod_csr_dense_float32 if f32 else rx.tcrossprod_csr_dense_numeric
    return fn(p, j, x, dense, nth, out=out)


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    d = float(np.max(np.abs(want))) if want.size else 0.0
    return float(np.max(np.abs(got - want)) / d) if d > 0 else float(np.max(np.abs(got))) if got.size else 0.0


def cpu_baseline(wl, p, j, x, dense=None, budget_s=12.0, rx=None):
    """Time the reference's CPU implementation on a bounded row sample of the SAME matrix (all rows when that fits the
    budget) and, with `rx`, check the GPU's level-1 result for exactly that input against it: returns (record, parity)."""
    from oracle.cpu_oracle import best_cpu_baseline
    cpu = best_cpu_baseline()
    cpu.copy_result = False
    nthreads = host_cores()  # passed to the reference's `num_threads(nthreads)` clause
    m = p.size - 1
    rng = np.random.default_rng(99)
    # probe on ~1/32 of the rows, then size the sample for the time budget
    rows = max(1, m // 32)

    def sample(r):
        pe = p[: r + 1]
        d = dense
        if d is None:
            d = cpu_dense_operand(wl, r, rng)
        elif wl["op"] == "crossprod":
            d = np.asfortranarray(dense[:, :r])
        return pe, j[: pe[-1]], x[: pe[-1]], d

    ps, js, xs, ds = sample(rows)
    t_probe, _ = cpu_run(cpu, wl, ps, js, xs, ds, nthreads)
    frac = min(1.0, max(1.0 / 32, (budget_s / 2) / max(t_probe, 1e-4) / 32))
    rows = max(1, int(m * frac))
    ps, js, xs, ds = sample(rows)
    best = min(cpu_run(cpu, wl, ps, js, xs, ds, nthreads)[0] for _ in range(2))
    nnz_s = int(ps[-1])
    gflops = 2.0 * nnz_s * wl["n"] / best / 1e9
    rec = {
        "value": gflops, "unit": "GFLOP/s", "cores": nthreads, "kind": cpu.kind,
        "sample": f"first {rows} of {m} rows ({nnz_s} nnz), best of 2, {best * 1e3:.1f} ms; build: {cpu.flags}; "
                  f"host: {os.cpu_count()} logical CPUs",
        "seconds": best,
    }
    parity = None
    if rx is not None:
        cpu.copy_result = True
        _, want = cpu_run(cpu, wl, ps, js, xs, ds, nthreads)
        got = gpu_level1(rx, wl, ps, js, xs, ds)
        tol = 1e-5 if wl["dtype"] == "f32" else 1e-12
        err = rel_err(got, want)
        parity = {"max_rel_err": err, "tol": tol, "ok": bool(err <= tol), "rows_checked": int(rows), "of_rows": int(m),
                  "elements_checked": int(np.asarray(want).size),
                  "against": f"the reference's src/matmul.cpp ({cpu.kind}, {cpu.flags}) on the same host arrays; "
                             "GPU side: the level-1 call that e2e times"}
    return rec, parity


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def _wall(f, k):
    t0 = time.perf_counter()
    for _ in range(k):
        r = f()
    return (time.perf_counter() - t0) / k, r


def _lib_bytes(_lib):
    import ctypes as _C
    up, down = _C.c_size_t(0), _C.c_size_t(0)
    _lib.call("mxg_last_call_bytes", _C.byref(up), _C.byref(down))
    return int(up.value), int(down.value)


def host_dma_roof(ndev, mb=256, reps=3):
    """Aggregate page-locked host <-> device rate of `ndev` GPUs copying at once (plain cudaMemcpyAsync, both directions
    together): the roof of a host-buffer call spread over those devices, whatever the library does."""
    import torch
    n = mb << 20
    bufs = []
    for g in range(ndev):
        with torch.cuda.device(g):
            bufs.append((torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
                         torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}"), torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}"),
                         torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)))
    out = {}
    for what in ("h2d", "d2h", "both"):
        best = None
        for _ in range(reps + 1):
            for g in range(ndev):
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for g, (hin, hout, din, dout, s_up, s_down) in enumerate(bufs):
                if what in ("h2d", "both"):
                    with torch.cuda.stream(s_up):
                        din.copy_(hin, non_blocking=True)
                if what in ("d2h", "both"):
                    with torch.cuda.stream(s_down):
                        hout.copy_(dout, non_blocking=True)
            for g in range(ndev):
                torch.cuda.synchronize(g)
            t = time.perf_counter() - t0
            best = t if best is None else min(best, t)
        moved = ndev * n * (2 if what == "both" else 1)
        out[what + "_GBps"] = moved / best / 1e9
    del bufs
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from matrixextra_b200 import _lib, rcpp_exports as rx
    from matrixextra_b200._lib import (MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_KEEP_F32, MXG_KEEP_F64,
                                       MXG_ROWS_CONTIGUOUS)
    from matrixextra_b200.device import DeviceCSR

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    _lib.call("mxg_set_device", local_rank)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        # collectives that overlap a product (sharded.PipelinedColumnMajorGather) must not queue behind its CTAs
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side waits that leave the other ranks' GPUs idle

    if world > 1 and args.strong_only:
        rec = strong_cfg5(args, dist, rank, world)
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({"strong_cfg5": rec}))
        return
    wl = dict(WORKLOADS[args.workload])
    if args.scale != 1.0:
        wl["m"] = max(64, int(wl["m"] * args.scale))
        wl["nnz"] = max(64, int(wl["nnz"] * args.scale))
    m, K, n = wl["m"], wl["K"], wl["n"]
    f32 = wl["dtype"] == "f32"
    s = 4 if f32 else 8
    tdt = torch.float32 if f32 else torch.float64
    mdt = MXG_F32 if f32 else MXG_F64
    keep = (MXG_KEEP_F32 if f32 else MXG_KEEP_F64)
    if wl["op"] == "spmv":
        keep = MXG_KEEP_F64

    # every rank owns one row block of the workload's size (weak scaling); seeds differ per rank
    A = DeviceCSR.synth(m, K, wl["nnz"], wl["row_model"], wl["col_model"], seed=wl["seed"] + 7919 * rank, keep=keep)
    nnz = A.nnz
    g = torch.Generator(device="cuda").manual_seed(4242)  # replicated dense operand: same on every rank
    op = wl["op"]
    At = None
    out_rows = K if op == "crossprod" else m
    # N > 1: the all-gather of north_star subsystem 4 is part of the step.  Row-major results (spmv, dense_tcsr): every
    # finished row is stored into ALL ranks' results over NVLink by the product kernel (peer stores / NVLS multicast),
    # or finished row slices are pushed by the copy engines (mxg_dev_spmm_push) — the fastest of those is the step.
    # Column-major results (csr_dense, cfg5's shape): copy-engine pushes of 2-D blocks.  crossprod keeps NCCL.
    fused = world > 1 and op in ("spmv", "dense_tcsr", "csr_dense")
    peer = None
    if op == "spmv":
        dense = torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
        shape_all, dt_all = (world * m,), torch.float64
    elif op == "crossprod":
        dense = torch.randn(m, n, device="cuda", dtype=tdt, generator=g)  # Y rows-contiguous [m][n]
        shape_all, dt_all = (world * K, n), tdt
    elif op == "csr_dense":
        dense = torch.randn(K, n, device="cuda", dtype=tdt, generator=g)
        shape_all, dt_all = (n, world * m), tdt  # ONE column-major (world*m x n) matrix, ldc = world*m
    else:
        dense = torch.randn(K, n, device="cuda", dtype=tdt, generator=g)  # (n x K) column-major R matrix
        shape_all, dt_all = (world * m, n), tdt
    if fused:
        from matrixextra_b200.sharded import PeerResult
        peer = PeerResult(int(np.prod(shape_all)) * (8 if dt_all == torch.float64 else 4), dist, rank, world)
        out_all = peer.tensor(shape_all, dt_all)
    else:
        out_all = torch.empty(shape_all, device="cuda", dtype=dt_all)
    if op == "csr_dense":
        out_local = out_all[:, rank * m:(rank + 1) * m] if world > 1 else out_all  # this rank's rows of every column
    else:
        out_local = out_all[rank * out_rows:(rank + 1) * out_rows]
    colmajor_tmp = None
    if op == "csr_dense" and world > 1:
        colmajor_tmp = torch.empty(n, m, device="cuda", dtype=tdt)  # contiguous local block for the NCCL variant
    dst = ldc_all = None
    if fused:
        if op == "spmv":
            dst = peer.dst_ptrs(rank * m * 8)
        elif op == "dense_tcsr":
            dst, ldc_all = peer.dst_ptrs(rank * m * n * s), n
        else:  # one global column-major (world*m x n) matrix: this rank owns rows [rank*m, (rank+1)*m)
            dst, ldc_all = peer.dst_ptrs(rank * m * s), world * m

    def compute(local_cm=None):
        nonlocal At
        if op == "spmv":
            A.spmv(dense, out_local)
        elif op == "dense_tcsr":
            A.spmm(dense, out_local, n, mdt, MXG_ROWS_CONTIGUOUS)
        elif op == "csr_dense":
            if local_cm is not None:
                A.spmm(dense, local_cm, n, mdt, MXG_COLS_CONTIGUOUS)
            else:
                A.spmm(dense, out_local, n, mdt, MXG_COLS_CONTIGUOUS, ldc=world * m)
        elif op == "crossprod":
            if At is not None:
                At.free()
            At = A.transpose(keep=keep)
            At.spmm(dense, out_local, n, mdt, MXG_ROWS_CONTIGUOUS)

    def step_bulk():
        _lib.set_option("spmm_bulk", 1)
        step_peer(keep_option=True)

    def step_bulk_cta():
        _lib.set_option("spmm_bulk", 2)
        step_peer(keep_option=True)

    def step_peer(keep_option=False):
        if not keep_option:
            _lib.set_option("spmm_bulk", 0)
        if op == "spmv":
            A.spmv_bcast(dense, dst)
        else:
            A.spmm_bcast(dense, dst, n, mdt, MXG_ROWS_CONTIGUOUS if op == "dense_tcsr" else MXG_COLS_CONTIGUOUS, ldc=ldc_all)
        peer.barrier()  # device-side: every rank's rows have landed when the stream gets past this

    def step_push():
        A.spmm_push(dense, dst, n, mdt, MXG_ROWS_CONTIGUOUS if op == "dense_tcsr" else MXG_COLS_CONTIGUOUS, ldc=ldc_all)
        peer.barrier()

    def step_plain():
        compute()
        if world > 1:
            dist.all_gather_into_tensor(out_all.view(-1), out_local.contiguous().view(-1))

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # candidates for the fused step; `--allgather auto` times each briefly and runs the step with the fastest
    variants, ms_probe, mres = {}, {}, None
    if fused:
        if op != "csr_dense":
            variants["peer"] = step_peer
        if op == "dense_tcsr":
            variants["bulk"] = step_bulk
            variants["bulk_cta"] = step_bulk_cta
        if op != "spmv":
            variants["push"] = step_push
        if op == "dense_tcsr" and args.allgather in ("auto", "mcast"):
            try:
                from matrixextra_b200.sharded import McastResult
                mres = McastResult(int(np.prod(shape_all)) * s, dist, rank, world)
                mc_block = mres.mc_ptr(rank * m * n * s)

                def step_mc():
                    A.spmm_mcast(dense, mc_block, n, mdt)
                    mres.barrier()
                variants["mcast"] = step_mc
            except Exception as e:  # noqa: BLE001  (no multicast support on this box)
                ms_probe["mcast_unavailable"] = str(e)[:200]
        if args.allgather != "auto" and args.allgather in variants:
            variants = {args.allgather: variants[args.allgather]}
        for name, fn in variants.items():
            for _ in range(3):
                fn()
            ms_probe[name] = timed(fn, 5) / 5
        how_fused = min(variants, key=lambda k: ms_probe[k])
        step_fn = variants[how_fused]
    else:
        how_fused, step_fn = None, step_plain
    for _ in range(max(args.warmup, 3)):
        step_fn()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_total = timed(step_fn, args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_compute = timed(compute, args.steps) if world > 1 else ms_total
    ms_nccl = None
    bit_identical = None
    if fused:
        # the unfused baseline — product into a contiguous local block, then one NCCL all-gather — and the proof that
        # the fused step left exactly those bytes on this rank (checked on every rank, AND-ed)
        step_fn()
        torch.cuda.synchronize()
        dist.barrier()
        nccl_all = torch.empty(int(np.prod(shape_all)), device="cuda", dtype=dt_all)

        def nccl_step():
            if op == "csr_dense":
                compute(colmajor_tmp)
                dist.all_gather_into_tensor(nccl_all, colmajor_tmp.view(-1))
            else:
                compute()
                dist.all_gather_into_tensor(nccl_all, out_local.contiguous().view(-1))
        nccl_step()
        torch.cuda.synchronize()
        if op == "csr_dense":  # NCCL gathered [world][n][m]; the fused step wrote [n][world*m]
            same = torch.equal(out_all.view(n, world, m), nccl_all.view(world, n, m).permute(1, 0, 2))
        else:
            same = torch.equal(out_all.reshape(-1), nccl_all)
        if how_fused == "mcast":
            same = same and torch.equal(mres.tensor(shape_all, dt_all).reshape(-1), nccl_all)
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        bit_identical = bool(flag.item())
        nccl_step()
        ms_nccl = timed(nccl_step, args.steps) / args.steps
        del nccl_all
        if peer.failed():
            raise SystemExit("peer barrier timed out")

    nnz_all = nnz
    if world > 1:
        t = torch.tensor([nnz], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        nnz_all = int(t.item())
    flops_step = 2.0 * nnz_all * n
    ms_step = ms_total / args.steps
    value = flops_step / (ms_step * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    if op == "crossprod":
        w_alg = (4 * (m + 1) + 12 * nnz) + (4 * (K + 1) + 12 * nnz) + w_alg_bytes(K, m, nnz, n, s)
    else:
        w_alg = w_alg_bytes(m, K, nnz, n, s)
    ms_kernel = ms_compute / args.steps
    achieved = w_alg / (ms_kernel * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.workload) if world == 1 else (None, None)
    fused_text = {"mcast": "fused into the product kernel (NVLS multicast stores)",
                  "bulk": "fused into the product kernel (finished rows parked in shared memory and shipped to every GPU as "
                          "cp.async.bulk copies over NVLink)",
                  "bulk_cta": "fused into the product kernel (a CTA's 32 finished rows parked in shared memory and shipped to every "
                              "GPU as one cp.async.bulk copy each over NVLink)",
                  "peer": "fused into the product kernel (NVLink peer stores)",
                  "push": "finished row slices pushed by the copy engines over NVLink while the next slice is computed",
                  None: "by NCCL"}[how_fused]

    result = {
        "metric": "CSR x dense SpMM GFLOP/s (k=%d) & effective HBM GB/s" % n if op != "spmv" else "CSR SpMV GFLOP/s & effective HBM GB/s",
        "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl["dtype"], "data": "synthetic (device-generated, Philox4x32-10; SURVEY.md 8d)",
        "config": bench_config(wl, world, nnz),
        "config_details": {"allgather": fused_text if world > 1 else None,
                   "l2_policy": ("operands (CSR %.0f MB + dense %.0f MB + out %.0f MB) " % (nnz * (4 + s) / 1e6, s * K * n / 1e6, s * out_rows * n / 1e6))
                   + ("exceed the 126 MB L2; no flush needed" if nnz * (4 + s) + s * K * n + s * out_rows * n > 2 * 126e6
                      else "FIT in the 126 MB L2: this is a warm-cache, launch-bound number (parity config, not the bench line)"),
                   "long_rows": A.n_long, "long_row_pieces": A.n_pieces, "longest_row": A.max_len},
        "effective_GBps": achieved,
        "compute_only": {"value": flops_step / (ms_kernel * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms_kernel},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": w_alg,
                     # the binding roof: HBM on the ACTUAL traffic of the launch (ncu dram bytes), several
                     # times the algorithmic bytes because the dense operand overflows L2 and its rows are re-fetched
                     "actual_traffic_GBps": traffic / ms_kernel / 1e6 if traffic else None,
                     "actual_traffic_frac_of_peak": traffic / ms_kernel / 1e6 / peak if traffic else None,
                     "gather_roofs": "random 256-byte rows, nothing attached (mxg_dev_gather_probe, profiles/"
                                     "r01_v5_gather_roof_probe.jsonl): 14.0 TB/s from L2, 9.0 TB/s on a 256 MB table, "
                                     "6.4 TB/s from DRAM",
                     "note": "W_alg = nnz*(4+s)+4(m+1)+s*K*n+s*m*n per GPU; gather-model bytes (B row per entry) = %.2f GB"
                             % ((nnz * (4 + s) + 4 * (m + 1) + s * nnz * n + s * m * n) / 1e9)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if world > 1:
        result["allgather"] = ({"how": fused_text, "variants_ms_per_step": ms_probe, "ms_per_step": ms_step,
                                "nccl_after_compute_ms_per_step": ms_nccl, "bit_identical_to_nccl": bit_identical,
                                "bytes_received_per_gpu": int((world - 1) * out_rows * n * s)}
                               if fused else {"how": "NCCL all_gather_into_tensor after the product", "ms_per_step": ms_step})

    # ---- end to end through the reference-facing entry point with host buffers ------------------------
    e2e = cpu = parity = None
    nth_all = max(1, min(16, os.cpu_count() or 4))
    if not args.skip_e2e:
        p_h, j_h, x_h = A.to_host()  # ordinary (pageable) numpy arrays: what an R session holds
        if op == "crossprod":
            d_h = np.asfortranarray(dense.cpu().numpy().T)  # t(Y): n x m column-major, the CPU route's operand
        elif op == "spmv":
            d_h = dense.cpu().numpy()
        else:
            d_h = np.ascontiguousarray(dense.cpu().numpy()).T  # (n x K) column-major view of the K x n row-major copy
    if peer is not None:
        del out_all, out_local
        peer.close(dist)
        peer = None
    if mres is not None:
        del mres
    A.free()
    del dense
    torch.cuda.empty_cache()

    if not args.skip_e2e and world == 1:
        e2e, parity, cpu = e2e_single(args, wl, rx, _lib, p_h, j_h, x_h, d_h, nnz_all, nth_all)
    elif not args.skip_e2e:
        e2e, parity = e2e_multi(args, wl, rx, _lib, dist, cpu_group, rank, world, p_h, j_h, x_h, d_h, nth_all)
    result["e2e"] = e2e
    result["cpu_baseline"] = cpu
    result["parity"] = parity

    if world > 1 and not args.skip_strong:
        result["strong_cfg5"] = strong_cfg5(args, dist, rank, world)

    if rank == 0 and world == 1 and args.others:
        result["others"] = {}
        for name in args.others.split(","):
            if name and name != args.workload:
                result["others"][name] = quick_kernel_bench(name, args)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


def _pinned(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def e2e_single(args, wl, rx, _lib, p_h, j_h, x_h, d_h, nnz_all, nth):
    """N = 1.  Headline: the call an R session makes — every operand in ordinary pageable memory, a NEW result matrix per
    call, allocated by the callee the way the Rcpp glue allocates it (Rf_allocVector3 with the library's allocator hooks,
    rglue/mxgpu_result_alloc.h: a recycled page-locked block; to R an ordinary matrix).  Sub-keys: the same call writing
    into a result on the caller's heap, with page-locked operands, with a device-resident matrix (explicit handle and
    level-1 cache), and the cpu_baseline / parity of the same product."""
    import torch
    op, n, K = wl["op"], wl["n"], wl["K"]
    m = p_h.size - 1
    f32 = wl["dtype"] == "f32"
    flops = 2.0 * nnz_all * n
    k_e2e = max(1, min(args.steps, 5))

    def rec(t, up_down=None, **kw):
        out = {"value": flops / t / 1e9, "ms_per_step": t * 1e3}
        if up_down:
            out["h2d_bytes_per_step"], out["d2h_bytes_per_step"] = up_down
        out.update(kw)
        return out

    call_pg = lambda: gpu_level1(rx, wl, p_h, j_h, x_h, d_h, nth)  # noqa: E731
    warm = [call_pg(), call_pg()]  # warm-up: allocator pools, page-locked arena, and TWO blocks in the result pool (a new
    del warm                       # result is allocated while the previous one is still referenced, in R as here)
    t_pg, res = _wall(call_pg, k_e2e)
    bytes_pg = _lib_bytes(_lib)
    res_shape, res_dtype = res.shape, res.dtype
    e2e = {"value": flops / t_pg / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": bytes_pg[0], "d2h_bytes_per_step": bytes_pg[1],
           "steps": k_e2e, "ms_per_step": t_pg * 1e3, "host_threads": nth,
           "entry_point": "level-1 C ABI (streamed row chunks) via the Rcpp-export mirror, called the way an R session "
                          "calls it: CSR and dense operand in PAGEABLE host memory (bounced through the library's page-locked "
                          "ring by its host threads), a NEW result per call allocated by the callee like the Rcpp glue does "
                          "(Rf_allocVector3 + the library's allocator hooks: a recycled page-locked block the device writes "
                          "into directly); bytes are the library's own count of the copies of the last timed call (float32 "
                          "values after host narrowing, packed column ids)"}
    del res
    if op != "crossprod":
        # the same call when the result lives on the CALLER's heap (a fresh np.empty / malloc per call: pages that do not
        # exist yet): bounced through a page-locked slot and first-touched by the host threads
        call_heap = lambda: gpu_level1(rx, wl, p_h, j_h, x_h, d_h, nth, out=np.empty(res_shape, dtype=res_dtype, order="F"))  # noqa: E731
        call_heap()
        t_heap, _ = _wall(call_heap, 3)
        e2e["caller_heap_result"] = rec(t_heap, _lib_bytes(_lib), note="everything pageable, the result a freshly allocated array "
                                        "on the caller's heap (round 1's `all_pageable`)")
    if op != "crossprod":
        # page-locked operands and a reused page-locked result (what the contract's e2e describes; R cannot do this)
        p_p, j_p, x_p = _pinned(p_h), _pinned(j_h), _pinned(x_h)
        if op == "spmv":
            d_p, o_p = _pinned(d_h), _pinned(np.empty(m))
        else:
            d_p = _pinned(d_h.T).T
            shape = (n, m) if op == "dense_tcsr" else (m, n)
            o_p = _pinned(np.empty(int(np.prod(shape)), dtype=d_h.dtype)).reshape(shape, order="F")
        call_pin = lambda: gpu_level1(rx, wl, p_p, j_p, x_p, d_p, nth, out=o_p)  # noqa: E731
        call_pin()
        t_pin, _ = _wall(call_pin, k_e2e)
        e2e["pinned"] = rec(t_pin, _lib_bytes(_lib), note="page-locked operands and a reused page-locked result buffer (DMA in place)")
        call_fresh = lambda: gpu_level1(rx, wl, p_p, j_p, x_p, d_p, nth, out=np.empty(res_shape, dtype=res_dtype, order="F"))  # noqa: E731
        call_fresh()
        t_fresh, _ = _wall(call_fresh, 3)
        e2e["fresh_pageable_result"] = rec(t_fresh, note="page-locked operands, newly allocated result on the caller's heap")
        del p_p, j_p, x_p, o_p
    if op in ("dense_tcsr", "csr_dense", "spmv"):
        # the matrix stays in HBM between calls; a product moves the dense operand up and the result down
        h = rx.as_gpu_csr(p_h, j_h, x_h, K, keep_float64=not f32 or op == "spmv", keep_float32=f32 and op != "spmv")
        if op == "spmv":
            warm = lambda d=d_h: rx.gpu_csr_dvec_numeric(h, d, nth)  # noqa: E731
        elif op == "dense_tcsr":
            fn = rx.gpu_csr_dense_tcrossprod_float32 if f32 else rx.gpu_csr_dense_tcrossprod_numeric
            warm = lambda d=d_h: fn(d, h, nth)  # noqa: E731
        else:
            fn = rx.gpu_csr_tcrossprod_dense_float32 if f32 else rx.gpu_csr_tcrossprod_dense_numeric
            warm = lambda d=d_h: fn(h, d, nth)  # noqa: E731
        warm()
        t_w, res_w = _wall(warm, k_e2e)
        e2e["warm_handle"] = rec(t_w, _lib_bytes(_lib), bit_identical_to_level1=bool(np.array_equal(res_w, call_pg())),
                                 note="explicit device-resident matrix (as_gpu_csr + gpu_csr_* exports, rglue/handle_gpu_glue.cpp); "
                                      "pageable dense operand, new glue-allocated result: only they cross PCIe")
        if op != "spmv":  # page-locked dense operand AND result: the PCIe bound of the warm path (R cannot do this)
            o_w = _pinned(np.empty(res_w.size, dtype=res_w.dtype)).reshape(res_w.shape, order="F")
            t_wp, _ = _wall(lambda: (fn(d_p, h, nth, out=o_w) if op == "dense_tcsr" else fn(h, d_p, nth, out=o_w)), k_e2e)
            e2e["warm_handle"]["pinned"] = rec(t_wp, _lib_bytes(_lib), note="page-locked dense operand and result buffer")
            # A/B on this box: the same two calls with the operand and the result crossing PCIe in one piece
            call_wp = lambda: (fn(d_p, h, nth, out=o_w) if op == "dense_tcsr" else fn(h, d_p, nth, out=o_w))  # noqa: E731
            _lib.set_option("host_colsplit", 0)
            call_wp()
            t_wp1, _ = _wall(call_wp, k_e2e)
            _lib.set_option("host_colsplit", 1)
            e2e["warm_handle"]["pinned"]["one_piece_ms_per_step"] = t_wp1 * 1e3
            e2e["warm_handle"]["pinned"]["note"] += ("; the operand and the result cross PCIe as two column halves so that the directions overlap "
                                                      "(option host_colsplit; `one_piece_ms_per_step` = the same call with it off, on this box)")
            del o_w
        rx.gpu_csr_free(h)
        del res_w
        # the same through the unchanged ten exports with the level-1 operand cache on (MATRIXEXTRA_GPU_CACHE_MB)
        _lib.set_option("cache_mb", 8192)
        call_pg()
        call_pg()
        t_c, _ = _wall(call_pg, k_e2e)
        e2e["cached_level1"] = rec(t_c, _lib_bytes(_lib), note="option cache_mb on: the device CSR (and the dense operand) of "
                                   "the previous call on the same host arrays is reused; unchanged level-1 exports")
        _lib.set_option("cache_mb", 0)
        _lib.call("mxg_cache_clear")
    cpu = parity = None
    if not args.skip_cpu:
        cpu, parity = cpu_baseline(wl, p_h, j_h, x_h, d_h, rx=rx)
    return e2e, parity, cpu


def e2e_multi(args, wl, rx, _lib, dist, cpu_group, rank, world, p_h, j_h, x_h, d_h, nth):
    """N > 1.  The reference parallelises inside ONE process (src/matmul.cpp:132-136) and an R session is one process:
    rank 0 makes ONE level-1 call on ONE workload matrix spread over all N GPUs of the box (mxg_set_devices(N): N host
    threads, N streamed pipelines, each device uploading its row block over its own PCIe link) while the other ranks
    wait on the host.  Strong scaling: the same call on one device is timed next to it."""
    op, n = wl["op"], wl["n"]
    flops = 2.0 * int(p_h[-1]) * n
    e2e = parity = None
    dist.barrier(group=cpu_group)
    if rank == 0 and op != "crossprod":
        k = max(1, min(args.steps, 5))
        out = {}
        call = lambda: gpu_level1(rx, wl, p_h, j_h, x_h, d_h, nth)  # noqa: E731
        p_p, j_p, x_p = _pinned(p_h), _pinned(j_h), _pinned(x_h)
        d_p = _pinned(d_h) if op == "spmv" else _pinned(d_h.T).T
        shape = (p_h.size - 1,) if op == "spmv" else ((n, p_h.size - 1) if op == "dense_tcsr" else (p_h.size - 1, n))
        o_p = _pinned(np.empty(int(np.prod(shape)), dtype=np.float64 if op == "spmv" else d_h.dtype)).reshape(shape, order="F")
        call_pin = lambda: gpu_level1(rx, wl, p_p, j_p, x_p, d_p, nth, out=o_p)  # noqa: E731
        ref_res = None
        for G in (1, world):
            _lib.call("mxg_set_devices", G)
            warm = [call(), call()]  # two blocks in the result pool
            del warm
            t_pg, res = _wall(call, k)
            b_pg = _lib_bytes(_lib)
            call_pin()
            t_pin, _ = _wall(call_pin, k)
            b_pin = _lib_bytes(_lib)
            if G == 1:
                ref_res = res
            out[G] = {"pageable_ms": t_pg * 1e3, "pageable_GFLOPs": flops / t_pg / 1e9, "pageable_bytes": b_pg,
                      "pinned_ms": t_pin * 1e3, "pinned_GFLOPs": flops / t_pin / 1e9, "pinned_bytes": b_pin,
                      "bit_identical_to_one_device": bool(np.array_equal(res, ref_res))}
        roof = host_dma_roof(world)
        moved = sum(out[world]["pinned_bytes"])
        e2e = {"value": out[world]["pageable_GFLOPs"], "unit": "GFLOP/s", "ms_per_step": out[world]["pageable_ms"],
               "h2d_bytes_per_step": out[world]["pageable_bytes"][0], "d2h_bytes_per_step": out[world]["pageable_bytes"][1],
               "steps": k, "scaling": "strong", "host_threads": nth, "devices": world,
               "entry_point": f"ONE level-1 call (Rcpp-export mirror) by ONE process on ONE workload matrix with mxg_set_devices({world}) "
                              "in force, called the way an R session calls it: pageable operands, a new pageable result.  Such a call is "
                              "bound by the host threads that bounce pageable memory, so the library keeps it on ONE device (option "
                              f"multi_pageable); `pinned` is the same call from page-locked arrays, which IS spread over the {world} GPUs: "
                              "nnz-balanced row blocks, one host thread + one streamed pipeline per device, the dense operand uploaded once "
                              "(a slice per device) and completed over NVLink",
               "pinned": {"value": out[world]["pinned_GFLOPs"], "ms_per_step": out[world]["pinned_ms"],
                          "h2d_bytes_per_step": out[world]["pinned_bytes"][0], "d2h_bytes_per_step": out[world]["pinned_bytes"][1]},
               "one_device_same_box": {"pageable_ms": out[1]["pageable_ms"], "pinned_ms": out[1]["pinned_ms"]},
               "speedup_vs_one_device": {"pageable": out[1]["pageable_ms"] / out[world]["pageable_ms"],
                                         "pinned": out[1]["pinned_ms"] / out[world]["pinned_ms"]},
               "bit_identical_to_one_device": out[world]["bit_identical_to_one_device"],
               "host_roof": dict(roof, note=f"page-locked host <-> device copies on all {world} GPUs at once (plain cudaMemcpyAsync, "
                                            "both directions together): what this host's memory system and PCIe tree deliver to "
                                            f"{world} devices, whatever moves the bytes",
                                 pinned_call_ms_at_roof=moved / roof["both_GBps"] / 1e6)}
        _lib.call("mxg_set_devices", 1)
        if not args.skip_cpu:
            # parity of the multi-device call against the reference on a bounded row sample of the same matrix
            _lib.call("mxg_set_devices", world)
            _lib.set_option("multi_min_nnz", 1 << 16)
            _lib.set_option("multi_pageable", 1)  # the check runs on pageable arrays: spread them all the same
            _, parity = cpu_baseline(wl, p_h, j_h, x_h, d_h, budget_s=6.0, rx=rx)
            parity["against"] += f" (spread over {world} devices)"
            _lib.call("mxg_set_devices", 1)
            _lib.set_option("multi_min_nnz", 4 << 20)
            _lib.set_option("multi_pageable", 0)
    dist.barrier(group=cpu_group)
    return e2e, parity


def strong_cfg5(args, dist, rank, world):
    """BASELINE.json configs[4]: CSR 50M x 10M, 2B nnz %*% dense 10M x 128 fp32, ROW-SHARDED over the N GPUs (strong
    scaling: rows/N and nnz/N per GPU), replicated dense operand, column-major (R layout) result all-gathered to every
    GPU.  The reference cannot run this shape at all: its `int` strides overflow (src/matmul.cpp:35-39, 156, 182)."""
    import torch
    from matrixextra_b200._lib import MXG_COLS_CONTIGUOUS, MXG_F32, MXG_KEEP_F32
    from matrixextra_b200.device import DeviceCSR
    from matrixextra_b200.sharded import PeerResult
    wl = WORKLOADS["cfg5"]
    sc = args.strong_scale
    mg, K, n = int(wl["m"] * sc) // world, wl["K"], wl["n"]
    nnz_g = int(wl["nnz"] * sc) // world
    free, _ = torch.cuda.mem_get_info()
    need = 8 * nnz_g + 4 * K * n + 3 * 4 * world * mg * n + 3 * 4 * mg * n + (8 << 30)
    if free < need:
        return {"skipped": f"needs {need / 1e9:.0f} GB per GPU, {free / 1e9:.0f} GB free"}
    A = DeviceCSR.synth(mg, K, nnz_g, wl["row_model"], wl["col_model"], seed=wl["seed"] + 7919 * rank, keep=MXG_KEEP_F32)
    g = torch.Generator(device="cuda").manual_seed(4242)
    dense = torch.randn(K, n, device="cuda", dtype=torch.float32, generator=g)
    peer = PeerResult(4 * world * mg * n, dist, rank, world)
    out_all = peer.tensor((n, world * mg), torch.float32)
    dst, ldc = peer.dst_ptrs(rank * mg * 4), world * mg

    def timed(fn, steps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        A.spmm_push(dense, dst, n, MXG_F32, MXG_COLS_CONTIGUOUS, ldc=ldc)
        peer.barrier()

    def compute():
        A.spmm(dense, out_all[:, rank * mg:(rank + 1) * mg], n, MXG_F32, MXG_COLS_CONTIGUOUS, ldc=ldc)

    from matrixextra_b200 import _lib
    local = torch.empty(n, mg, device="cuda", dtype=torch.float32)
    nccl_all = torch.empty(world * n * mg, device="cuda", dtype=torch.float32)
    nccl_out = out_all  # the re-arranged copy lands in the same result buffer (checked against the other variants' bits)

    import ctypes as _C
    cur = torch.cuda.current_stream

    def nccl_step(rearrange=False):
        # the unfused baseline: product into a contiguous local block, ONE all-gather ([G][n][m]), and — for the result R
        # needs — the local re-arrangement of the G blocks into one column-major (G*m x n) matrix (copy engines)
        A.spmm(dense, local, n, MXG_F32, MXG_COLS_CONTIGUOUS)
        dist.all_gather_into_tensor(nccl_all, local.view(-1))
        if rearrange:
            for q in range(world):
                _lib.call("mxg_dev_copy_2d", _C.c_void_p(nccl_out.data_ptr() + q * mg * 4), world * mg * 4,
                          _C.c_void_p(nccl_all.data_ptr() + q * n * mg * 4), mg * 4, mg * 4, n, _C.c_void_p(cur().cuda_stream))

    from matrixextra_b200.sharded import PipelinedColumnMajorGather
    pipe = PipelinedColumnMajorGather(A, n, MXG_F32, torch.float32, dist, rank, world, slices=8)
    step_pipe = lambda: pipe.step(dense)  # noqa: E731

    def step_pipe_after():  # the same pipeline behind the whole product instead of behind its slices
        pipe.overlap_product = False
        pipe.step(dense)
        pipe.overlap_product = True
    for _ in range(2):
        step()
        step_pipe()
        step_pipe_after()
    steps = 3
    ms_push = timed(step, steps)
    ms_pipe = timed(step_pipe, steps)
    ms_pipe_after = timed(step_pipe_after, steps)
    ms_compute = timed(compute, steps)
    step()
    step_pipe()
    nccl_step()
    torch.cuda.synchronize()
    want = nccl_all.view(world, n, mg).permute(1, 0, 2)
    same = torch.equal(out_all.view(n, world, mg), want) and torch.equal(pipe.out_all.view(n, world, mg), want)
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ms_nccl_raw = timed(nccl_step, steps)
    ms_nccl = timed(lambda: nccl_step(True), steps)
    torch.cuda.synchronize()
    same_r = torch.equal(out_all.view(n, world, mg), want)
    flag2 = torch.tensor([1 if same_r else 0], device="cuda")
    dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
    ms_step = min(ms_push, ms_pipe, ms_pipe_after, ms_nccl)
    del pipe
    t = torch.tensor([A.nnz], device="cuda", dtype=torch.int64)
    dist.all_reduce(t)
    nnz_all = int(t.item())
    failed = peer.failed()
    del out_all
    peer.close(dist)
    A.free()
    one_gpu_ms = 163.0 * sc  # `others.cfg5` of the N = 1 line re-measures it on every run
    return {"workload": wl["desc"] + f" row-sharded over {world} GPUs" + ("" if sc == 1.0 else f" (scaled x{sc})"),
            "scaling": "strong", "rows_per_gpu": mg, "nnz_total": nnz_all, "steps": steps,
            "ms_per_step": ms_step, "GFLOPs": 2.0 * nnz_all * n / ms_step / 1e6,
            "compute_only_ms": ms_compute, "nccl_after_compute_ms": ms_nccl, "nccl_after_compute_without_rearrangement_ms": ms_nccl_raw,
            "variants_ms_per_step": {"push": ms_push, "pipelined_nccl": ms_pipe, "pipelined_nccl_after_product": ms_pipe_after,
                                     "nccl_after_compute": ms_nccl},
            "how": {ms_pipe_after: "the whole product into this GPU's rows of the global column-major result, then 8 row slices "
                                   "packed (copy engine), all-gathered with NCCL and unpacked into the result (copy engines) in a "
                                   "three-stage pipeline (sharded.PipelinedColumnMajorGather, overlap_product=False)",
                    ms_pipe: "product in 8 row slices into this GPU's rows of the global column-major result; behind every slice "
                             "one stream packs it (copy engine) and all-gathers it with NCCL, another unpacks the received slices "
                             "into the result (copy engines) while the next slice is computed (sharded.PipelinedColumnMajorGather)",
                    ms_push: "product in ~16 row slices per GPU into the global column-major result; every finished slice pushed "
                             "to the other GPUs as a 2-D block (n column segments) by the copy engines over NVLink while the next "
                             "slice is computed (mxg_dev_spmm_push) + device-side flag barrier",
                    ms_nccl: "product into a contiguous local column-major block, ONE NCCL all-gather of the blocks ([G][n][m]), "
                             "then the local re-arrangement into one column-major (G*m x n) matrix on copy engines"}[ms_step],
            "bit_identical_to_nccl": bool(flag.item()) and bool(flag2.item()) and not failed,
            "bytes_received_per_gpu": int(4 * (world - 1) * mg * n),
            "ingress_GBps": 4 * (world - 1) * mg * n / ms_step / 1e6,
            "one_gpu_ms": one_gpu_ms, "one_gpu_source": "163 ms for the whole of cfg5 on one B200 (profiles/r01_v6_bench_cfg5_kernel_only.json; "
                                                       "`--workload cfg5` re-measures it)",
            "speedup_vs_one_gpu": one_gpu_ms / ms_step, "efficiency": one_gpu_ms / ms_step / world}


def quick_kernel_bench(name, args):
    """Kernel-only numbers for the other BASELINE configs (same timing rules, fewer fields)."""
    import torch

    from matrixextra_b200._lib import (MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_KEEP_F32, MXG_KEEP_F64,
                                       MXG_ROWS_CONTIGUOUS)
    from matrixextra_b200.device import DeviceCSR
    wl = WORKLOADS[name]
    m, K, n = wl["m"], wl["K"], wl["n"]
    f32 = wl["dtype"] == "f32"
    s = 4 if f32 else 8
    tdt = torch.float32 if f32 else torch.float64
    mdt = MXG_F32 if f32 else MXG_F64
    keep = MXG_KEEP_F64 if (wl["op"] == "spmv" or not f32) else MXG_KEEP_F32
    need = wl["nnz"] * (4 + s) + s * K * n + s * m * n + (8 << 30)
    if torch.cuda.mem_get_info()[0] < need:
        return {"workload": wl["desc"], "skipped": f"needs {need / 1e9:.0f} GB of device memory"}
    A = DeviceCSR.synth(m, K, wl["nnz"], wl["row_model"], wl["col_model"], seed=wl["seed"], keep=keep)
    g = torch.Generator(device="cuda").manual_seed(4242)
    op = wl["op"]
    extra = {}
    if op == "spmv":
        dense = torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
        out = torch.empty(m, device="cuda", dtype=torch.float64)
        fn = lambda: A.spmv(dense, out)  # noqa: E731
    elif op == "dense_tcsr":
        dense = torch.randn(K, n, device="cuda", dtype=tdt, generator=g)
        out = torch.empty(m, n, device="cuda", dtype=tdt)
        fn = lambda: A.spmm(dense, out, n, mdt, MXG_ROWS_CONTIGUOUS)  # noqa: E731
    elif op == "csr_dense":
        dense = torch.randn(K, n, device="cuda", dtype=tdt, generator=g)
        out = torch.empty(n, m, device="cuda", dtype=tdt)
        fn = lambda: A.spmm(dense, out, n, mdt, MXG_COLS_CONTIGUOUS)  # noqa: E731
    elif op == "crossprod":
        dense = torch.randn(m, n, device="cuda", dtype=tdt, generator=g)
        out = torch.empty(K, n, device="cuda", dtype=tdt)
        At = A.transpose(keep=keep)
        fn = lambda: At.spmm(dense, out, n, mdt, MXG_ROWS_CONTIGUOUS)  # noqa: E731

        def tr():
            t = A.transpose(keep=keep)
            t.free()
        ms_t = _time_ms(tr, 3, 2)
        extra = {"transpose_ms": ms_t, "transpose_GBps": (24 * A.nnz + 4 * (m + K + 2)) / ms_t / 1e6}
    ms = _time_ms(fn, min(args.steps, 5) if name == "cfg5" else args.steps, 3)
    nnz = A.nnz
    w = w_alg_bytes(K if op == "crossprod" else m, m if op == "crossprod" else K, nnz, n, s)
    peak, _ = measured_peaks()
    out_d = {"workload": wl["desc"], "ms_per_step": ms, "GFLOPs": 2.0 * nnz * n / ms / 1e6,
             "effective_GBps": w / ms / 1e6, "roofline_frac": w / ms / 1e6 / peak, "nnz": nnz}
    out_d.update(extra)
    A.free()
    del dense, out
    torch.cuda.empty_cache()
    return out_d


def _time_ms(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on this box's host cores
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's src/matmul.cpp (oracle/_ref, all host threads) on the FULL workload matrix — the same shape, nnz,
    row / column laws, RNG and counters as the GPU arm's device-generated matrix, produced on the host by the oracle's
    twin of the generator (oracle/mx_synth.c).  Loads nothing of the product (no libmxgpu.so, no CUDA)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_oracle import best_cpu_baseline, synth_csr_host
    wl = dict(WORKLOADS[args.workload])
    if args.scale != 1.0:
        wl["m"] = max(64, int(wl["m"] * args.scale))
        wl["nnz"] = max(64, int(wl["nnz"] * args.scale))
    cpu = best_cpu_baseline()
    cpu.copy_result = False
    nthreads = host_cores()  # passed to the reference's `num_threads(nthreads)` clause
    if wl["nnz"] > 400_000_000:
        raise SystemExit("the reference cannot run this shape (int strides, src/matmul.cpp:35-39); use cfg1..cfg4")
    p, j, x = synth_csr_host(wl["m"], wl["K"], wl["nnz"], wl["row_model"], wl["col_model"], wl["seed"])
    rng = np.random.default_rng(99)
    dense = cpu_dense_operand(wl, p.size - 1, rng)
    warm = max(1, args.warmup)
    t_one, _ = cpu_run(cpu, wl, p, j, x, dense, nthreads)
    # every requested step is timed unless that would take more than ~3 minutes on this host
    steps = max(1, min(args.steps, int(150.0 / max(t_one, 1e-3))))
    warm = max(1, min(warm, int(30.0 / max(t_one, 1e-3))))
    for _ in range(warm - 1):
        cpu_run(cpu, wl, p, j, x, dense, nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_run(cpu, wl, p, j, x, dense, nthreads)
    sec = (time.perf_counter() - t0) / steps
    nnz = int(p[-1])
    value = 2.0 * nnz * wl["n"] / sec / 1e9
    sample = (f"all {p.size - 1} rows of the workload matrix ({nnz} nnz; host twin of the device generator, same seed), "
              f"{steps} timed passes after {warm} warm-up; build: {cpu.flags}")
    print(json.dumps({
        "impl": "reference", "metric": ("CSR x dense SpMM GFLOP/s (k=%d) & effective HBM GB/s" % wl["n"]) if wl["op"] != "spmv"
        else "CSR SpMV GFLOP/s & effective HBM GB/s", "value": value,
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"],
        "data": "synthetic (host-generated by the oracle's twin of the device generator, Philox4x32-10; SURVEY.md 8d)",
        "config": bench_config(wl, args.gpus, nnz),
        "config_details": {"cpu_arm": f"one row block of the workload on the host: {nthreads} OpenMP threads, schedule(dynamic) as "
                                      "written (src/matmul.cpp:132-136)"},
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": nthreads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)  # 100 x ~3 ms: long enough for several clock samples
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink rows/nnz (debugging only; invalid as a bench value)")
    ap.add_argument("--allgather", default="auto", choices=["auto", "mcast", "peer", "bulk", "bulk_cta", "push"],
                    help="all-gather of the step at N > 1: NVLS multicast stores / peer stores from the product kernel, "
                         "copy-engine pushes of finished row slices, or the fastest of them")
    ap.add_argument("--others", default="cfg2,k64f64,cfg4,cfg5", help="extra kernel-only configs reported at N=1 ('' to skip)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-strong", action="store_true", help="N > 1: skip the row-sharded cfg5 record")
    ap.add_argument("--strong-only", action="store_true", help="N > 1: only the row-sharded cfg5 record (development)")
    ap.add_argument("--strong-scale", type=float, default=1.0, help="shrink cfg5 for the strong-scaling record (debugging)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner, ...) is sent to
    # stderr by pointing fd 1 at fd 2 for the duration of the run; print() below writes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
