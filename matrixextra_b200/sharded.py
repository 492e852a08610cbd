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


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerResult:
    """The full result of a row-sharded product, present on every GPU of the box, filled by all ranks at once.

    Every rank allocates one buffer (``mxg_dev_alloc``: plain cudaMalloc, exportable), publishes its cudaIpc
    handle through the process group (host-side plumbing only) and maps the other ranks' buffers.  A rank's
    product kernel then stores each finished output row into all ``world`` buffers (local store + NVLink peer
    stores, ``DeviceCSR.spmm_bcast``), and ``barrier()`` — a tiny device-side flag exchange on the same stream —
    tells every rank when all blocks have landed.  No separate all-gather pass exists.

    Buffer layout: 256 bytes of flags (``world`` ints used), then the payload."""

    HEADER = 256

    def __init__(self, payload_bytes: int, dist, rank: int, world: int):
        import ctypes as C

        import torch

        from . import _lib
        self._lib, self._C, self.rank, self.world = _lib, C, rank, world
        self.payload_bytes = int(payload_bytes)
        base = C.c_void_p()
        _lib.call("mxg_dev_alloc", self.HEADER + self.payload_bytes, C.byref(base))
        self.base = int(base.value)
        self.ptrs: List[int] = [0] * world
        self.ptrs[rank] = self.base
        self._opened: List[int] = []
        torch.as_tensor(_RawCuda(self.base, (self.HEADER // 4,), "<i4"), device="cuda").zero_()
        torch.cuda.synchronize()
        if world > 1:
            handle = (C.c_ubyte * 64)()
            _lib.call("mxg_ipc_export", C.c_void_p(self.base), handle)
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle))
            for g in range(world):
                if g == rank:
                    continue
                buf = (C.c_ubyte * 64).from_buffer_copy(handles[g])
                q = C.c_void_p()
                _lib.call("mxg_ipc_open", buf, C.byref(q))
                self.ptrs[g] = int(q.value)
                self._opened.append(int(q.value))
            dist.barrier()
        self.epoch = 0

    def payload_ptr(self, g: int) -> int:
        return self.ptrs[g] + self.HEADER

    def dst_ptrs(self, byte_offset: int) -> List[int]:
        """Where this rank's block starts inside every rank's result, local buffer first."""
        order = [self.rank] + [g for g in range(self.world) if g != self.rank]
        return [self.payload_ptr(g) + int(byte_offset) for g in order]

    def tensor(self, shape, dtype):
        """torch view of the local payload."""
        import torch
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]
        return torch.as_tensor(_RawCuda(self.payload_ptr(self.rank), shape, typestr), device="cuda")

    def barrier(self, stream=None):
        """Stream-ordered: returns immediately; work queued behind it sees every rank's rows."""
        from .device import _stream_ptr
        C = self._C
        self.epoch += 1
        arr = (C.c_void_p * self.world)(*self.ptrs)
        self._lib.call("mxg_dev_peer_barrier", self.rank, self.world, arr, self.epoch, _stream_ptr(stream))

    def failed(self) -> bool:
        f = self._C.c_int(0)
        self._lib.call("mxg_dev_barrier_failed", self._C.byref(f))
        return bool(f.value)

    def close(self, dist=None):
        import torch
        torch.cuda.synchronize()
        if dist is not None and self.world > 1:
            dist.barrier()
        for q in self._opened:
            self._lib.call("mxg_ipc_close", self._C.c_void_p(q))
        self._opened = []
        if dist is not None and self.world > 1:
            dist.barrier()
        if self.base:
            self._lib.call("mxg_dev_free", self._C.c_void_p(self.base))
            self.base = 0


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


class McastResult:
    """The full result of a row-sharded product in NVLS MULTICAST memory (torch symmetric memory does the plumbing:
    one allocation per rank, exchanged handles, a multicast mapping over all of them).  A rank's product kernel
    writes each finished row ONCE to the multicast address (``DeviceCSR.spmm_mcast``) and the NVSwitch replicates it
    into every GPU's buffer, the local one included; ``barrier()`` is symmetric memory's stream-ordered device
    barrier.  Raises if the box has no multicast support (then ``PeerResult`` is the way)."""

    def __init__(self, payload_bytes: int, dist, rank: int, world: int):
        import torch
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = rank, world
        self.payload_bytes = int(payload_bytes)
        self.buf = symm.empty(self.payload_bytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        self.mc_base = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if self.mc_base == 0:
            raise RuntimeError("symmetric memory gave no multicast pointer (no NVLS on this box)")

    def tensor(self, shape, dtype):
        """torch view of the LOCAL copy of the result."""
        return self.buf.view(dtype)[: int(__import__("numpy").prod(shape))].view(shape)

    def mc_ptr(self, byte_offset: int) -> int:
        return self.mc_base + int(byte_offset)

    def barrier(self):
        self.handle.barrier()


class PipelinedColumnMajorGather:
    """Row-sharded ``A %*% B`` with a COLUMN-MAJOR (R layout) result gathered to every GPU (BASELINE cfg5's shape).

    A rank's block of a column-major (G*m x n) matrix is n strided column segments, which neither peer stores (128-byte
    segments) nor copy-engine pushes (2-D peer copies) move at NVLink speed on 8 GPUs; NCCL's all-gather does, on
    contiguous buffers.  So the product is issued in ``slices`` row slices (long rows first, ``DeviceCSR.spmm_rows``)
    straight into the rank's rows of the global result, and behind every slice
      * the COMM stream packs the slice into a contiguous (n x rows) buffer (one 2-D copy on a copy engine) and
        all-gathers it with NCCL,
      * the UNPACK stream moves the G - 1 received slices into their rows of the global result (2-D copies on copy
        engines; the rank's own slice is already in place),
    while the next slice is being computed: the step costs about max(product, all-gather) instead of their sum."""

    def __init__(self, A, n, dtype, torch_dtype, dist, rank, world, slices=8, overlap_product=True):
        """overlap_product=False: the whole product first, then the slices are packed / gathered / unpacked in a pipeline of
        their own (the exchange no longer competes with the product for HBM and SMs, but nothing hides the product)."""
        import torch
        self.A, self.n, self.dtype, self.dist, self.rank, self.world = A, n, dtype, dist, rank, world
        self.overlap_product = overlap_product
        m = A.m
        self.bounds = [m * k // slices for k in range(slices + 1)]
        rows_max = max(self.bounds[k + 1] - self.bounds[k] for k in range(slices))
        self.out_all = torch.empty((n, world * m), device="cuda", dtype=torch_dtype)  # column-major (world*m x n), ldc = world*m
        self.pack = [torch.empty(n * rows_max, device="cuda", dtype=torch_dtype) for _ in range(2)]
        self.stage = [torch.empty(world * n * rows_max, device="cuda", dtype=torch_dtype) for _ in range(2)]
        self.comm, self.unpack = torch.cuda.Stream(), torch.cuda.Stream()
        self.ev_done = [torch.cuda.Event() for _ in range(slices)]
        self.ev_gathered = [torch.cuda.Event() for _ in range(slices)]
        self.ev_unpacked = [torch.cuda.Event() for _ in range(slices)]

    def step(self, B_t):
        import ctypes as C

        import torch

        from . import _lib
        from ._lib import MXG_COLS_CONTIGUOUS
        A, n, G, m = self.A, self.n, self.world, self.A.m
        ldc = G * m
        es = self.out_all.element_size()
        base = self.out_all.data_ptr()
        origin = base + self.rank * m * es  # this rank's first row inside column 0
        main = torch.cuda.current_stream()

        def copy2d(dst, dpitch, src, spitch, width, height, stream):
            _lib.call("mxg_dev_copy_2d", C.c_void_p(dst), dpitch, C.c_void_p(src), spitch, width, height, C.c_void_p(stream.cuda_stream))

        if not self.overlap_product:
            _lib.call("mxg_dev_spmm", A._h, int(self.dtype), MXG_COLS_CONTIGUOUS, 0, int(n), C.c_void_p(B_t.data_ptr()), int(n),
                      C.c_void_p(origin), int(ldc), C.c_void_p(main.cuda_stream))
        elif A.n_pieces > 0:
            A.spmm_rows(B_t, origin, n, self.dtype, MXG_COLS_CONTIGUOUS, ldc, pieces=True)
        self.comm.wait_stream(main)
        self.unpack.wait_stream(main)
        S = len(self.bounds) - 1
        for k in range(S):
            r0, r1 = self.bounds[k], self.bounds[k + 1]
            rows = r1 - r0
            if self.overlap_product:
                A.spmm_rows(B_t, origin, n, self.dtype, MXG_COLS_CONTIGUOUS, ldc, r0, r1)
            self.ev_done[k].record(main)
            b = k % 2
            send = self.pack[b][: n * rows]
            recv = self.stage[b][: G * n * rows]
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ev_done[k])
                if k >= 2:
                    self.comm.wait_event(self.ev_unpacked[k - 2])  # the staging buffer's previous slice has left it
                copy2d(send.data_ptr(), rows * es, origin + r0 * es, ldc * es, rows * es, n, self.comm)
                self.dist.all_gather_into_tensor(recv, send)
                self.ev_gathered[k].record(self.comm)
            self.unpack.wait_event(self.ev_gathered[k])
            for g in range(G):
                if g != self.rank:
                    copy2d(base + (g * m + r0) * es, ldc * es, recv.data_ptr() + g * n * rows * es, rows * es, rows * es, n, self.unpack)
            self.ev_unpacked[k].record(self.unpack)
        main.wait_stream(self.comm)
        main.wait_stream(self.unpack)
