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


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch, from the committed
# `ncu --set full` captures of exactly these workloads (profiles/README.md); None where no capture exists.
NCU_TRAFFIC = {
    "cfg3": (15.632157e9 + 0.570318e9, "profiles/r01_v6_spmm_f32_k64_cfg3.ncu.txt"),
    "k64f64": (39.124897e9 + 1.030659e9, "profiles/r01_v6_spmm_f64_k64.ncu.txt"),
    "cfg2": (1.220678e9 + 0.022964e9, "profiles/r01_v6_spmv_f64_cfg2.ncu.txt"),
}


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
    """One pass of the reference's CPU path on host arrays; returns seconds (best of the call itself)."""
    t0 = time.perf_counter()
    if wl["op"] == "dense_tcsr":
        fn = cpu.tcrossprod_dense_csr_float32 if wl["dtype"] == "f32" else cpu.tcrossprod_dense_csr_numeric
        fn(dense, p, j, x, nthreads, wl["K"])
    elif wl["op"] == "csr_dense":
        fn = cpu.tcrossprod_csr_dense_float32 if wl["dtype"] == "f32" else cpu.tcrossprod_csr_dense_numeric
        fn(p, j, x, dense, nthreads)
    elif wl["op"] == "spmv":
        cpu.matmul_csr_dvec_numeric(p, j, x, dense, nthreads)
    elif wl["op"] == "crossprod":
        # the all-MatrixExtra CPU route: stable CSR->CSC, then matmul_dense_csc on t(Y)
        from oracle.cpu_oracle import Port
        p2, i2, x2 = Port().csr2csc(p.size - 1, wl["K"], p, j, x)
        cpu.matmul_dense_csc_numeric(dense, p2, i2, x2, nthreads)
    else:
        raise ValueError(wl["op"])
    return time.perf_counter() - t0


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


def cpu_baseline(wl, p, j, x, budget_s=12.0):
    """Time the reference's CPU implementation on a bounded row sample of the SAME matrix."""
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
        return pe, j[: pe[-1]], x[: pe[-1]]

    ps, js, xs = sample(rows)
    dense = cpu_dense_operand(wl, rows, rng)
    t_probe = cpu_run(cpu, wl, ps, js, xs, dense, nthreads)
    frac = min(1.0, max(1.0 / 32, (budget_s / 2) / max(t_probe, 1e-4) / 32))
    rows = max(1, int(m * frac))
    ps, js, xs = sample(rows)
    dense = cpu_dense_operand(wl, rows, rng)
    best = min(cpu_run(cpu, wl, ps, js, xs, dense, nthreads) for _ in range(2))
    nnz_s = int(ps[-1])
    gflops = 2.0 * nnz_s * wl["n"] / best / 1e9
    return {
        "value": gflops, "unit": "GFLOP/s", "cores": nthreads, "kind": cpu.kind,
        "sample": f"first {rows} of {m} rows ({nnz_s} nnz), best of 2, {best * 1e3:.1f} ms; build: {cpu.flags}; "
                  f"host: {os.cpu_count()} logical CPUs",
        "seconds": best,
    }


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
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
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

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
    # N > 1: the products store every finished row into ALL ranks' results over NVLink (PeerResult), so the
    # all-gather of north_star subsystem 4 is fused into the kernel; crossprod keeps the NCCL collective
    # (column-major results, op csr_dense, keep the NCCL collective: peer stores of 128-byte column segments 200 MB
    #  apart reach only ~130 GB/s — cfg5 sharded over 8 GPUs: fused 170 ms, product + NCCL all-gather 53.6 ms,
    #  profiles/r01_v5_bench_cfg5_sharded_8gpu.json — so the fused kernel is used where it wins: row-major rows)
    fused = world > 1 and op in ("spmv", "dense_tcsr")
    peer = None
    if op == "spmv":
        dense = torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
        shape_all, dt_all = (world * m,), torch.float64
    elif op == "crossprod":
        dense = torch.randn(m, n, device="cuda", dtype=tdt, generator=g)  # Y rows-contiguous [m][n]
        shape_all, dt_all = (world * K, n), tdt
    else:
        dense = torch.randn(K, n, device="cuda", dtype=tdt, generator=g)  # (n x K) column-major R matrix
        shape_all, dt_all = (world * m, n), tdt
    if fused:
        from matrixextra_b200.sharded import PeerResult
        peer = PeerResult(int(np.prod(shape_all)) * (8 if dt_all == torch.float64 else 4), dist, rank, world)
        out_all = peer.tensor(shape_all, dt_all)
    else:
        out_all = torch.empty(shape_all, device="cuda", dtype=dt_all)
    out_local = out_all[rank * out_rows:(rank + 1) * out_rows]
    colmajor_tmp = None
    if op == "csr_dense":
        # column-major (R layout) output block m x n, ldc = m  (N = 1, and the compute-only / NCCL variants)
        colmajor_tmp = torch.empty(n, m, device="cuda", dtype=tdt)
    dst = ldc_all = None
    if fused:
        if op == "spmv":
            dst = peer.dst_ptrs(rank * m * 8)
        elif op == "dense_tcsr":
            dst, ldc_all = peer.dst_ptrs(rank * m * n * s), n
        else:  # one global column-major (world*m x n) matrix: this rank owns rows [rank*m, (rank+1)*m)
            dst, ldc_all = peer.dst_ptrs(rank * m * s), world * m

    def compute():
        nonlocal At
        if op == "spmv":
            A.spmv(dense, out_local)
        elif op == "dense_tcsr":
            A.spmm(dense, out_local, n, mdt, MXG_ROWS_CONTIGUOUS)
        elif op == "csr_dense":
            A.spmm(dense, colmajor_tmp, n, mdt, MXG_COLS_CONTIGUOUS)
        elif op == "crossprod":
            if At is not None:
                At.free()
            At = A.transpose(keep=keep)
            At.spmm(dense, out_local, n, mdt, MXG_ROWS_CONTIGUOUS)

    def step(with_gather=True):
        if fused and with_gather:
            if op == "spmv":
                A.spmv_bcast(dense, dst)
            else:
                A.spmm_bcast(dense, dst, n, mdt, MXG_ROWS_CONTIGUOUS if op == "dense_tcsr" else MXG_COLS_CONTIGUOUS, ldc=ldc_all)
            peer.barrier()  # device-side: every rank's rows have landed when the stream gets past this
            return
        compute()
        if world > 1 and with_gather:
            src = colmajor_tmp if op == "csr_dense" else out_local
            dist.all_gather_into_tensor(out_all.view(-1), src.view(-1))

    def nccl_step():
        # the unfused baseline: product into the local block, then one NCCL all-gather of the blocks
        compute()
        src = colmajor_tmp if op == "csr_dense" else out_local
        dist.all_gather_into_tensor(nccl_all.view(-1), src.contiguous().view(-1))

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

    # NVLS multicast variant of the fused all-gather (rows-contiguous results): every row is written once with
    # multimem.st and replicated by the NVSwitch; used for the timed step when the box offers it and it is faster
    step_fn, how_fused, ms_probe = step, "peer", {}
    mres = None
    if fused and op == "dense_tcsr" and args.allgather != "peer":
        try:
            from matrixextra_b200.sharded import McastResult
            mres = McastResult(int(np.prod(shape_all)) * s, dist, rank, world)
            mc_block = mres.mc_ptr(rank * m * n * s)

            def step_mc():
                A.spmm_mcast(dense, mc_block, n, mdt)
                mres.barrier()
            for _ in range(3):
                step_mc()
                step()
            ms_probe = {"mcast": timed(step_mc, 5) / 5, "peer": timed(step, 5) / 5}
            if args.allgather == "mcast" or ms_probe["mcast"] < ms_probe["peer"]:
                step_fn, how_fused = step_mc, "mcast"
        except Exception as e:  # noqa: BLE001  (no multicast support: keep the peer-store kernel)
            ms_probe = {"mcast_unavailable": str(e)[:200]}
    for _ in range(max(args.warmup, 3)):
        step_fn()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_total = timed(step_fn, args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_compute = timed(lambda: step(False), args.steps) if world > 1 else ms_total
    ms_nccl = None
    if fused:
        nccl_all = torch.empty(shape_all, device="cuda", dtype=dt_all)
        for _ in range(2):
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

    result = {
        "metric": "CSR x dense SpMM GFLOP/s (k=%d) & effective HBM GB/s" % n if op != "spmv" else "CSR SpMV GFLOP/s & effective HBM GB/s",
        "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl["dtype"], "data": "synthetic (device-generated, Philox4x32-10; SURVEY.md 8d)",
        "config": {"workload": wl["desc"], "rows_per_gpu": m, "cols": K, "nnz_per_gpu": nnz, "n": n,
                   "parallelism": (f"row-block shards x{world}, replicated dense operand, all-gather of the output row blocks "
                                   + (("fused into the product kernel (NVLS multicast stores)" if how_fused == "mcast"
                                       else "fused into the product kernel (NVLink peer stores)") if fused else "by NCCL"))
                   if world > 1 else "single GPU",
                   "l2_policy": ("operands (CSR %.0f MB + dense %.0f MB + out %.0f MB) " % (nnz * (4 + s) / 1e6, s * K * n / 1e6, s * out_rows * n / 1e6))
                   + ("exceed the 126 MB L2; no flush needed" if nnz * (4 + s) + s * K * n + s * out_rows * n > 2 * 126e6
                      else "FIT in the 126 MB L2: this is a warm-cache, launch-bound number (parity config, not the bench line)"),
                   "long_rows": A.n_long, "long_row_pieces": A.n_pieces, "longest_row": A.max_len},
        "effective_GBps": achieved,
        "compute_only": {"value": flops_step / (ms_kernel * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms_kernel},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0] if world == 1 else None,
                     "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1] if world == 1 else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": w_alg,
                     # the binding roof: HBM on the ACTUAL traffic of the launch (ncu dram bytes), which is
                     # ~10x the algorithmic bytes because the dense operand overflows L2 and its rows are re-fetched
                     "actual_traffic_GBps": (NCU_TRAFFIC[args.workload][0] / ms_kernel / 1e6
                                             if world == 1 and args.workload in NCU_TRAFFIC else None),
                     "actual_traffic_frac_of_peak": (NCU_TRAFFIC[args.workload][0] / ms_kernel / 1e6 / peak
                                                     if world == 1 and args.workload in NCU_TRAFFIC else None),
                     "gather_roofs": "random 256-byte rows, nothing attached (mxg_dev_gather_probe, profiles/"
                                     "r01_v5_gather_roof_probe.jsonl): 14.0 TB/s from L2, 9.0 TB/s on a 256 MB table, "
                                     "6.4 TB/s from DRAM",
                     "note": "W_alg = nnz*(4+s)+4(m+1)+s*K*n+s*m*n per GPU; gather-model bytes (B row per entry) = %.2f GB"
                             % ((nnz * (4 + s) + 4 * (m + 1) + s * nnz * n + s * m * n) / 1e9)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if world > 1:
        result["allgather"] = ({"how": ("fused, NVLS multicast: every finished row is written once with multimem.st by the "
                                        "product kernel (mxg_dev_spmm_mcast) and replicated by the NVSwitch into all ranks' "
                                        "results + symmetric-memory barrier" if how_fused == "mcast" else
                                        "fused: every finished row is stored into all ranks' results over NVLink by the product "
                                        "kernel (mxg_dev_spmm_bcast) + device-side flag barrier"),
                                "variants_ms_per_step": ms_probe, "ms_per_step": ms_step,
                                "nccl_after_compute_ms_per_step": ms_nccl, "bytes_received_per_gpu": int((world - 1) * out_rows * n * s)}
                               if fused else {"how": "NCCL all_gather_into_tensor after the product", "ms_per_step": ms_step})

    # ---- end to end through the reference-facing entry point with host buffers (rank-local) ----------
    e2e = None
    cpu = None
    if not args.skip_e2e:
        p_h, j_h, x_h = A.to_host()
        # host staging threads of this rank (the exports' `nthreads`): the box's cores are shared by the ranks
        nth = max(1, min(16, (os.cpu_count() or 4) // max(world, 1)))
        # With several ranks calling at once the host's memory bandwidth, not the per-GPU PCIe links, bounds the
        # calls, and narrowing on the host costs 20 instead of 12 bytes of host memory traffic per entry (measured:
        # 8 ranks 127 ms per call with host narrowing, 100 ms without; 2 ranks 31.9 vs 30.2 ms): narrow on the device.
        if world >= 2:
            _lib.set_option("host_narrow", 0)
        np_t = np.float32 if f32 else np.float64
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
        p_h, j_h, x_h = pin(p_h), pin(j_h), pin(x_h)
        def pinned_out(shape, dt):
            return torch.empty(int(np.prod(shape)), dtype=dt).pin_memory().numpy().reshape(shape, order="F")

        if op == "spmv":
            d_h = pin(dense.cpu().numpy())
            o_h = pinned_out((m,), torch.float64)
            call = lambda out=o_h: rx.matmul_csr_dvec_numeric(p_h, j_h, x_h, d_h, nth, out=out)  # noqa: E731
            h2d = p_h.nbytes + j_h.nbytes + x_h.nbytes + d_h.nbytes
            d2h = 8 * m
        elif op == "crossprod":
            d_h = pin(np.asfortranarray(dense.cpu().numpy()))  # Y (m x n) column-major
            o_h = None
            call = lambda out=None: rx.crossprod_csr_dense(p_h, j_h, x_h, K, d_h, mdt)  # noqa: E731
            h2d = p_h.nbytes + j_h.nbytes + x_h.nbytes + d_h.nbytes
            d2h = s * K * n
        else:
            d_h = torch.from_numpy(dense.cpu().numpy()).pin_memory().numpy().T  # (n x K) F-order view, pinned
            fn = {("dense_tcsr", True): rx.tcrossprod_dense_csr_float32, ("dense_tcsr", False): rx.tcrossprod_dense_csr_numeric,
                  ("csr_dense", True): rx.tcrossprod_csr_dense_float32, ("csr_dense", False): rx.tcrossprod_csr_dense_numeric}[(op, f32)]
            if op == "dense_tcsr":
                o_h = pinned_out((n, m), tdt)
                call = lambda out=o_h: fn(d_h, p_h, j_h, x_h, nth, K, out=out)  # noqa: E731
            else:
                o_h = pinned_out((m, n), tdt)
                call = lambda out=o_h: fn(p_h, j_h, x_h, d_h, nth, out=out)  # noqa: E731
            h2d = p_h.nbytes + j_h.nbytes + x_h.nbytes + d_h.nbytes
            d2h = s * m * n
        call()  # warm-up (allocator pools)
        k_e2e = max(1, min(args.steps, 5))

        def time_calls(f, k):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(k):
                r = f()
            t = (time.perf_counter() - t0) / k
            if world > 1:
                tt = torch.tensor([t], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt.item())
            return t, r

        t_e2e, res = time_calls(call, k_e2e)
        import ctypes as _C
        h2d_lib, d2h_lib = _C.c_size_t(0), _C.c_size_t(0)
        _lib.call("mxg_last_call_bytes", _C.byref(h2d_lib), _C.byref(d2h_lib))  # the last timed call, copy by copy
        # same call the way R makes it: the result is a freshly allocated (pageable, untouched) matrix
        call(out=None)  # warm-up: the page-locked arena grows by the output slots once
        t_fresh, res = time_calls(lambda: call(out=None), 3)
        # bytes that cross PCIe: a float32 product narrows the float64 values on the host (hoststage.cu)
        host_narrow = f32 and op != "crossprod" and _lib.get_option("host_narrow") != 0 and _lib.get_option("pipeline") != 0
        if host_narrow:
            h2d -= x_h.nbytes // 2
        # ... and column ids travel packed (2 / 2.5 / 3 bytes per entry) in the chunks where the link, not the host
        # threads, is the bottleneck (pipeline.cu upload_chunk): take the library's own count of what it copied
        host_pack = False
        if h2d_lib.value > 0:
            host_pack = h2d_lib.value < h2d - 1024
            h2d, d2h = h2d_lib.value, d2h_lib.value
        e2e = {"value": 2.0 * nnz_all * n / t_e2e / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": k_e2e, "ms_per_step": t_e2e * 1e3,
               "entry_point": "level-1 C ABI (streamed row chunks) via the Rcpp-export mirror; pinned host inputs, "
                              "result into a page-locked host buffer"
                              + ("; float64 values narrowed to float32 by the library's host threads before the copy "
                                 "(h2d bytes counted after narrowing)" if host_narrow else "")
                              + ("; column ids of the chunks uploaded while the link was the bottleneck packed to 2-3 "
                                 "bytes per entry by the same threads and rebuilt on the device" if host_pack else "")
                              + ("; bytes are the library's count of its copies in the last timed call" if h2d_lib.value else ""),
               "host_threads": nth,
               "fresh_pageable_result": {"value": 2.0 * nnz_all * n / t_fresh / 1e9, "ms_per_step": t_fresh * 1e3,
                                         "note": "same call returning a newly allocated pageable matrix, as the Rcpp glue does"}}
        if op not in ("crossprod",) and world == 1:
            # everything pageable, as in an R session: operands are ordinary (non page-locked) arrays, the result is new
            pg = lambda a: np.array(a, copy=True, order="K")  # noqa: E731
            p_g, j_g, x_g, d_g = pg(p_h), pg(j_h), pg(x_h), pg(d_h)
            if op == "spmv":
                call_pg = lambda: rx.matmul_csr_dvec_numeric(p_g, j_g, x_g, d_g, nth)  # noqa: E731
            elif op == "dense_tcsr":
                call_pg = lambda: fn(d_g, p_g, j_g, x_g, nth, K)  # noqa: E731
            else:
                call_pg = lambda: fn(p_g, j_g, x_g, d_g, nth)  # noqa: E731
            call_pg()
            t_pg, res = time_calls(call_pg, 3)
            e2e["all_pageable"] = {"value": 2.0 * nnz_all * n / t_pg / 1e9, "ms_per_step": t_pg * 1e3,
                                   "note": "operands AND result in pageable memory (an R session): bounced through the "
                                           "library's page-locked ring by its host threads"}
            del p_g, j_g, x_g, d_g
        del res
        if rank == 0 and world == 1 and not args.skip_cpu:
            cpu = cpu_baseline(wl, p_h, j_h, x_h)
    result["e2e"] = e2e
    result["cpu_baseline"] = cpu

    if rank == 0 and world == 1 and args.others:
        result["others"] = {}
        A.free()
        del dense, out_all
        torch.cuda.empty_cache()
        for name in args.others.split(","):
            if name and name != args.workload:
                result["others"][name] = quick_kernel_bench(name, args)

    if peer is not None:
        del out_all, out_local
        peer.close(dist)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


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
    ms = _time_ms(fn, args.steps, 3)
    nnz = A.nnz
    w = w_alg_bytes(K if op == "crossprod" else m, m if op == "crossprod" else K, nnz, n, s)
    peak, _ = measured_peaks()
    out_d = {"workload": wl["desc"], "ms_per_step": ms, "GFLOPs": 2.0 * nnz * n / ms / 1e6,
             "effective_GBps": w / ms / 1e6, "roofline_frac": w / ms / 1e6 / peak, "nnz": nnz}
    out_d.update(extra)
    A.free()
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_oracle import best_cpu_baseline
    wl = dict(WORKLOADS[args.workload])
    cpu = best_cpu_baseline()
    cpu.copy_result = False
    nthreads = host_cores()  # passed to the reference's `num_threads(nthreads)` clause
    # bounded sample of the same workload: the first rows of the same synthetic matrix when a GPU is there to
    # generate it, else a host-generated matrix with the same row-length law
    m_s = max(1, wl["m"] // 8)
    nnz_s = max(1, wl["nnz"] // 8)
    p = j = x = None
    try:
        import torch
        if torch.cuda.is_available():
            from matrixextra_b200._lib import MXG_KEEP_F64
            from matrixextra_b200.device import DeviceCSR
            A = DeviceCSR.synth(m_s, wl["K"], nnz_s, wl["row_model"], wl["col_model"], seed=wl["seed"], keep=MXG_KEEP_F64)
            p, j, x = A.to_host()
            A.free()
    except Exception:
        p = None
    if p is None:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import powerlaw_csr
        m_s = min(m_s, 100_000)
        p, j, x = powerlaw_csr(m_s, wl["K"], wl["nnz"] / wl["m"], 1, cap=min(wl["K"], 65536))
    rng = np.random.default_rng(99)
    dense = cpu_dense_operand(wl, p.size - 1, rng)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_run(cpu, wl, p, j, x, dense, nthreads)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_run(cpu, wl, p, j, x, dense, nthreads)
    sec = (time.perf_counter() - t0) / steps
    nnz = int(p[-1])
    value = 2.0 * nnz * wl["n"] / sec / 1e9
    sample = (f"rows 0..{p.size - 2} of the workload matrix generator ({nnz} nnz, 1/8 of one GPU's block), "
              f"{steps} timed passes; build: {cpu.flags}")
    print(json.dumps({
        "impl": "reference", "metric": "CSR x dense SpMM GFLOP/s (k=%d) & effective HBM GB/s" % wl["n"], "value": value,
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"],
        "data": "synthetic", "config": {"workload": wl["desc"], "sampled_rows": int(p.size - 1), "sampled_nnz": nnz},
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
    ap.add_argument("--allgather", default="auto", choices=["auto", "mcast", "peer"],
                    help="fused all-gather of row-major results at N > 1: NVLS multicast stores, peer stores, or the faster")
    ap.add_argument("--others", default="cfg2,k64f64,cfg4", help="extra kernel-only configs reported at N=1 ('' to skip)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
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
