"""The fused product + all-gather (``mxg_dev_spmm_bcast`` / ``mxg_dev_spmv_bcast`` + ``mxg_dev_peer_barrier``):
1 GPU is enough to exercise it — (a) several local destinations in one process, (b) two PROCESSES sharing the
device through cudaIpc handles, each storing its row block into both processes' result buffers, exactly the
code path of ``bench.py --gpus N`` (there the peers are other GPUs and the stores cross NVLink)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multiple_local_destinations_hold_identical_bits():
    import torch
    from matrixextra_b200._lib import MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_ROWS_CONTIGUOUS
    from matrixextra_b200.device import DeviceCSR
    A = DeviceCSR.synth(30000, 8000, 900000, row_model=1, col_model=1, seed=77)
    assert A.n_long > 0
    g = torch.Generator(device="cuda").manual_seed(3)
    from matrixextra_b200 import _lib
    # n = 64 (fp32 and fp64), n = 32 fp32: rows exactly one column block wide -> the bulk-copy instantiation (finished rows
    # leave shared memory as cp.async.bulk copies, one per destination); n = 24: peer stores from registers
    for dtype, tdt, n in ((MXG_F32, torch.float32, 64), (MXG_F64, torch.float64, 64), (MXG_F32, torch.float32, 32),
                          (MXG_F64, torch.float64, 24)):
        B = torch.randn(A.K, n, device="cuda", dtype=tdt, generator=g)
        for layout in (MXG_ROWS_CONTIGUOUS, MXG_COLS_CONTIGUOUS):
            want = torch.empty(A.m * n, device="cuda", dtype=tdt)
            A.spmm(B, want, n, dtype, layout)
            for bulk in (1, 2, 0):  # warp-wide bulk copies, CTA-wide bulk copies, register stores
                _lib.set_option("spmm_bulk", bulk)
                for n_dst in (3, 8):
                    outs = [torch.full((A.m * n,), float("nan"), device="cuda", dtype=tdt) for _ in range(n_dst)]
                    A.spmm_bcast(B, [o.data_ptr() for o in outs], n, dtype, layout)
                    for o in outs:
                        assert torch.equal(o, want)
            _lib.set_option("spmm_bulk", 1)
            # the copy-engine variant: product in row slices into outs[0], finished slices pushed to the others by DMA
            outs = [torch.full((A.m * n,), float("nan"), device="cuda", dtype=tdt) for _ in range(3)]
            A.spmm_push(B, [o.data_ptr() for o in outs], n, dtype, layout)
            torch.cuda.synchronize()
            for o in outs:
                assert torch.equal(o, want)
    y = torch.randn(A.K, device="cuda", dtype=torch.float64, generator=g)
    want = torch.empty(A.m, device="cuda", dtype=torch.float64)
    A.spmv(y, want)
    outs = [torch.full((A.m,), float("nan"), device="cuda", dtype=torch.float64) for _ in range(4)]
    A.spmv_bcast(y, [o.data_ptr() for o in outs])
    for o in outs:
        assert torch.equal(o, want)
    A.free()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from matrixextra_b200._lib import MXG_COLS_CONTIGUOUS, MXG_F32, MXG_ROWS_CONTIGUOUS
    from matrixextra_b200.device import DeviceCSR
    from matrixextra_b200.sharded import PeerResult
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # host plumbing only: handle exchange
    try:
        m, K, n = 20000, 6000, 64
        blocks = [DeviceCSR.synth(m, K, 500000, 1, 1, seed=500 + g, keep=2) for g in range(world)]  # every rank can rebuild every block
        gen = torch.Generator(device="cuda").manual_seed(9)
        B = torch.randn(K, n, device="cuda", dtype=torch.float32, generator=gen)
        ok = True
        for layout in (MXG_ROWS_CONTIGUOUS, MXG_COLS_CONTIGUOUS):
            res = PeerResult(world * m * n * 4, dist, rank, world)
            full = res.tensor((world * m * n,), torch.float32)
            full.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            if layout == MXG_ROWS_CONTIGUOUS:
                off, ldc = rank * m * n * 4, n
            else:
                off, ldc = rank * m * 4, world * m  # column-major global result (world*m x n)
            for _ in range(3):  # several epochs of the flag protocol
                blocks[rank].spmm_bcast(B, res.dst_ptrs(off), n, MXG_F32, layout, ldc=ldc)
                res.barrier()
            torch.cuda.synchronize()
            ok = ok and not res.failed()
            # reference: all blocks computed locally with the plain kernel
            for g in range(world):
                want = torch.empty(m * n, device="cuda", dtype=torch.float32)
                blocks[g].spmm(B, want, n, MXG_F32, layout)
                if layout == MXG_ROWS_CONTIGUOUS:
                    got = full[g * m * n:(g + 1) * m * n]
                else:
                    got = full.view(n, world * m)[:, g * m:(g + 1) * m].reshape(-1)
                ok = ok and bool(torch.equal(got, want))
            res.close(dist)
        q.put((rank, ok))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_processes_fill_each_others_result_through_ipc():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=240) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(ok for _, ok in results), results
