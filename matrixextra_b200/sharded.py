"""Row-block sharding of the CSR across the GPUs of one box (SURVEY.md §8 e).

Every output row depends on one CSR row and the whole dense operand, so the rows are split into
contiguous, nnz-balanced blocks (binary search of g*nnz/G in indptr: ``mxg_row_partition``), the dense
operand is replicated, every rank multiplies its block with the single-GPU kernels and the output row
blocks are exchanged with ONE collective: an all-gather (NCCL over NVLink on the GPU box).

One process per GPU (``torch.distributed``); this module only holds the host-side plan and the
collective call — the compute callable is injected, so the plan is testable on CPU with gloo.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Sequence

import numpy as np

from .device import row_partition


@dataclass
class RowShard:
    rank: int
    row_start: int
    row_end: int
    p: np.ndarray  # rebased indptr of the block (starts at 0)
    j: np.ndarray
    x: np.ndarray

    @property
    def rows(self) -> int:
        return self.row_end - self.row_start

    @property
    def nnz(self) -> int:
        return int(self.p[-1])


def plan_row_blocks(p: np.ndarray, parts: int) -> np.ndarray:
    """Row boundaries r_0=0 <= ... <= r_parts=m balancing the stored entries per block."""
    return row_partition(np.ascontiguousarray(p, dtype=np.int32), parts)


def take_shard(p: np.ndarray, j: np.ndarray, x: np.ndarray, bounds: Sequence[int], rank: int) -> RowShard:
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    e0, e1 = int(p[r0]), int(p[r1])
    return RowShard(rank, r0, r1, (p[r0:r1 + 1] - p[r0]).astype(np.int32), np.ascontiguousarray(j[e0:e1]),
                    np.ascontiguousarray(x[e0:e1]))


def allgather_row_blocks(local_block, bounds: Sequence[int], n_cols: int, dist, rank: int, world: int):
    """All-gather of row-major output blocks with (possibly) unequal row counts.

    Blocks are padded to the largest block so that a single equal-count all-gather can be used (the
    collective NCCL runs fastest); the padding rows are dropped when the full matrix is assembled.
    ``local_block`` is a torch tensor [rows_local, n_cols] on the collective's device."""
    import torch
    rows = [int(bounds[g + 1]) - int(bounds[g]) for g in range(world)]
    max_rows = max(rows) if rows else 0
    pad = torch.zeros((max_rows, n_cols), dtype=local_block.dtype, device=local_block.device)
    pad[: rows[rank]] = local_block
    gathered = torch.empty((world * max_rows, n_cols), dtype=local_block.dtype, device=local_block.device)
    dist.all_gather_into_tensor(gathered, pad)
    if all(r == max_rows for r in rows):
        return gathered
    return torch.cat([gathered[g * max_rows: g * max_rows + rows[g]] for g in range(world)], dim=0)


def sharded_spmm(p, j, x, B_rows, compute: Callable, dist, rank: int, world: int):
    """Out(m x n, row-major) = A . B with A row-sharded over ``world`` ranks.

    ``compute(shard: RowShard, B_rows) -> torch tensor [shard.rows, n]`` runs the single-device product
    (the CUDA kernels in production; injected so the plan can be exercised without a GPU).
    Every rank returns the full result."""
    bounds = plan_row_blocks(p, world)
    shard = take_shard(p, j, x, bounds, rank)
    local = compute(shard, B_rows)
    return allgather_row_blocks(local, bounds, local.shape[1], dist, rank, world)


def cuda_compute(dtype, stream=None) -> Callable:
    """The production compute callable: upload the shard once, multiply on the current device."""
    import torch

    from ._lib import MXG_F32, MXG_KEEP_F32, MXG_KEEP_F64, MXG_ROWS_CONTIGUOUS
    from .device import DeviceCSR

    def run(shard: RowShard, B_rows):
        f32 = dtype == MXG_F32
        K, n = B_rows.shape
        A = DeviceCSR.upload(shard.rows, K, shard.p, shard.j, shard.x, MXG_KEEP_F32 if f32 else MXG_KEEP_F64)
        out = torch.empty((shard.rows, n), device=B_rows.device, dtype=B_rows.dtype)
        if shard.rows:
            A.spmm(B_rows, out, n, dtype, MXG_ROWS_CONTIGUOUS, stream=stream)
        torch.cuda.current_stream().synchronize()
        A.free()
        return out

    return run
