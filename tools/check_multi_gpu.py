#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, one rank per GPU): the fused product + all-gather
(mxg_dev_spmm_bcast / mxg_dev_spmv_bcast over cudaIpc peer pointers + device-side flag barrier) must leave, on
every rank, exactly the bytes that `product into the local block` + NCCL all_gather leaves.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from matrixextra_b200 import _lib  # noqa: E402
from matrixextra_b200._lib import MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_ROWS_CONTIGUOUS  # noqa: E402
from matrixextra_b200.device import DeviceCSR  # noqa: E402
from matrixextra_b200.sharded import McastResult, PeerResult  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    _lib.call("mxg_set_device", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m, K, nnz = 300_000, 200_000, 12_000_000
    A = DeviceCSR.synth(m, K, nnz, 1, 1, seed=900 + rank)
    gen = torch.Generator(device="cuda").manual_seed(11)
    report = {}
    for name, dtype, tdt, n in (("f32_n64", MXG_F32, torch.float32, 64), ("f64_n24", MXG_F64, torch.float64, 24)):
        B = torch.randn(K, n, device="cuda", dtype=tdt, generator=gen)
        s = 4 if tdt == torch.float32 else 8
        for lname, layout in (("rows", MXG_ROWS_CONTIGUOUS), ("cols", MXG_COLS_CONTIGUOUS)):
            res = PeerResult(world * m * n * s, dist, rank, world)
            full = res.tensor((world * m * n,), tdt)
            full.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            if layout == MXG_ROWS_CONTIGUOUS:
                off, ldc = rank * m * n * s, n
            else:
                off, ldc = rank * m * s, world * m
            for _ in range(2):
                A.spmm_bcast(B, res.dst_ptrs(off), n, dtype, layout, ldc=ldc)
                res.barrier()
            torch.cuda.synchronize()
            local_out = torch.empty(m * n, device="cuda", dtype=tdt)
            A.spmm(B, local_out, n, dtype, layout)
            gathered = torch.empty(world * m * n, device="cuda", dtype=tdt)
            dist.all_gather_into_tensor(gathered, local_out)
            if layout == MXG_ROWS_CONTIGUOUS:
                ok = torch.equal(full, gathered)
            else:  # NCCL gathered [world][n][m]; fused wrote one column-major (world*m x n) matrix = [n][world*m]
                ok = torch.equal(full.view(n, world, m), gathered.view(world, n, m).permute(1, 0, 2))
            report[f"{name}_{lname}"] = bool(ok) and not res.failed()
            # the same all-gather by the copy engines (mxg_dev_spmm_push): slices pushed while the next is computed
            full.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            for _ in range(2):
                A.spmm_push(B, res.dst_ptrs(off), n, dtype, layout, ldc=ldc)
                res.barrier()
            torch.cuda.synchronize()
            if layout == MXG_ROWS_CONTIGUOUS:
                ok = torch.equal(full, gathered)
            else:
                ok = torch.equal(full.view(n, world, m), gathered.view(world, n, m).permute(1, 0, 2))
            report[f"{name}_{lname}_push"] = bool(ok) and not res.failed()
            del full
            res.close(dist)
    # column-major results gathered by the pipelined pack / NCCL all-gather / unpack behind every row slice
    from matrixextra_b200.sharded import PipelinedColumnMajorGather
    for name, dtype, tdt, n in (("f32_n64", MXG_F32, torch.float32, 64), ("f64_n24", MXG_F64, torch.float64, 24)):
        B = torch.randn(K, n, device="cuda", dtype=tdt, generator=gen)
        pipe = PipelinedColumnMajorGather(A, n, dtype, tdt, dist, rank, world, slices=5)
        pipe.out_all.fill_(float("nan"))
        for _ in range(2):
            pipe.step(B)
        torch.cuda.synchronize()
        local_out = torch.empty(m * n, device="cuda", dtype=tdt)
        A.spmm(B, local_out, n, dtype, MXG_COLS_CONTIGUOUS)
        gathered = torch.empty(world * m * n, device="cuda", dtype=tdt)
        dist.all_gather_into_tensor(gathered, local_out)
        report[f"{name}_cols_pipelined"] = bool(torch.equal(pipe.out_all.view(n, world, m), gathered.view(world, n, m).permute(1, 0, 2)))
        del pipe
    # the same through NVLS multicast (rows-contiguous results): one multimem.st per row, replicated by the switch
    try:
        for name, dtype, tdt, n in (("f32_n64", MXG_F32, torch.float32, 64), ("f64_n24", MXG_F64, torch.float64, 24)):
            s = 4 if tdt == torch.float32 else 8
            B = torch.randn(K, n, device="cuda", dtype=tdt, generator=gen)
            mres = McastResult(world * m * n * s, dist, rank, world)
            full = mres.tensor((world * m * n,), tdt)
            full.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            for _ in range(2):
                A.spmm_mcast(B, mres.mc_ptr(rank * m * n * s), n, dtype)
                mres.barrier()
            torch.cuda.synchronize()
            local_out = torch.empty(m * n, device="cuda", dtype=tdt)
            A.spmm(B, local_out, n, dtype, MXG_ROWS_CONTIGUOUS)
            gathered = torch.empty(world * m * n, device="cuda", dtype=tdt)
            dist.all_gather_into_tensor(gathered, local_out)
            report[f"{name}_rows_mcast"] = bool(torch.equal(full, gathered))
            del full, mres
    except RuntimeError as e:  # no multicast on this box
        report["mcast_unavailable"] = True
        if rank == 0:
            print("multicast unavailable:", e, file=sys.stderr)
    y = torch.randn(K, device="cuda", dtype=torch.float64, generator=gen)
    res = PeerResult(world * m * 8, dist, rank, world)
    full = res.tensor((world * m,), torch.float64)
    full.fill_(float("nan"))
    torch.cuda.synchronize()
    dist.barrier()
    A.spmv_bcast(y, res.dst_ptrs(rank * m * 8))
    res.barrier()
    torch.cuda.synchronize()
    loc = torch.empty(m, device="cuda", dtype=torch.float64)
    A.spmv(y, loc)
    gathered = torch.empty(world * m, device="cuda", dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, loc)
    report["spmv"] = bool(torch.equal(full, gathered)) and not res.failed()
    del full
    res.close(dist)
    flags = torch.tensor([int(all(report.values()))], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "all_ranks_ok": bool(flags.item()), "rank0": report}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flags.item() else 1)


if __name__ == "__main__":
    main()
