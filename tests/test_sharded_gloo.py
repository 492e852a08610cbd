"""world_size-2 gloo test (CPU) of the multi-GPU plan: nnz-balanced row blocks, replicated dense operand,
one all-gather of the output row blocks.  The per-rank product is injected (the CPU oracle here — test
infrastructure; the CUDA kernels in production), so this covers exactly the host-side sharding logic."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_file, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from helpers import powerlaw_csr
    from matrixextra_b200.sharded import plan_row_blocks, sharded_spmm
    from oracle.cpu_oracle import Port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_file)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = Port()
    p, j, x = powerlaw_csr(501, 300, 12, seed=3, cap=250)
    rng = np.random.default_rng(5)
    B = torch.from_numpy(rng.standard_normal((300, 16)))  # K x n rows-contiguous, replicated

    def compute(shard, B_rows):
        Xr = np.asfortranarray(B_rows.numpy().T)  # (n x K) column-major, as R passes it
        out = port.tcrossprod_dense_csr_numeric(Xr, shard.p, shard.j, shard.x)  # (n x rows) F-order
        return torch.from_numpy(np.ascontiguousarray(out.T))

    full = sharded_spmm(p, j, x, B, compute, dist, rank, world)
    want = port.tcrossprod_dense_csr_numeric(np.asfortranarray(B.numpy().T), p, j, x).T
    bounds = plan_row_blocks(p, world)
    ok = bool(np.array_equal(full.numpy(), want)) and full.shape == (501, 16)
    q.put((rank, ok, bounds.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_row_sharded_product_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=150) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(ok for _, ok, _ in results)
    b0 = results[0][2]
    assert b0 == results[1][2] and b0[0] == 0 and b0[-1] == 501 and 0 < b0[1] < 501
