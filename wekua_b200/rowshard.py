"""Row-sharded multi-GPU GEMM (BASELINE config 5).

The reference's multi-device mode is "one CommandQueue per device" (src/core/command_queue.zig:160-181): no op ever
splits work.  Here the split is explicit: C[M,N] = alpha*op(A)*op(B) + beta*C is partitioned into `world` contiguous
row blocks of op(A) and C, B is replicated, every rank (one process per GPU) computes its block with no
communication, and the blocks are reassembled on every GPU by the GEMM epilogue itself: each finished 16-byte piece
of C is stored into the rank's own buffer AND into every peer's buffer through CUDA-IPC mapped pointers, so the NVLink
transfer overlaps the tensor-core work tile by tile (wk_gemm_rowshard_allgather).  The host logic here -- the
partition, the view of A a rank owns, the handle exchange -- is pure Python and covered by world_size-2 gloo tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .tensor import Tensor, _scalar


def shard_rows(m: int, world: int, rank: int) -> tuple[int, int]:
    """(row0, rows) of rank's contiguous block of the M rows: M // world rows each, the last rank takes the rest"""
    if world < 1 or not 0 <= rank < world or m < world:
        raise capi.InvalidValue(f"InvalidValue: cannot shard {m} rows over {world} ranks (rank {rank})")
    base = m // world
    row0 = rank * base
    return row0, (m - row0 if rank == world - 1 else base)


def a_block(op_a: int, row0: int, rows: int, k: int, lda: int) -> tuple[int, tuple[int, int]]:
    """(element offset into the stored A, stored shape) of the part of A that produces C rows [row0, row0+rows).
    op_a = N: A is [M,K] -> rows row0.. of it (a row block).  op_a = T: A is [K,M] -> columns row0.. (a column block
    with the same pitch: a strided view, no copy)."""
    if op_a == 0:
        return row0 * lda, (rows, k)
    return row0, (k, rows)


def exchange(obj, world: int, all_gather_object=None) -> list:
    """every rank's `obj`, in rank order.  `all_gather_object(out_list, obj)` is torch.distributed's by default."""
    out = [None] * world
    if world == 1:
        out[0] = obj
        return out
    if all_gather_object is None:
        import torch.distributed as dist

        all_gather_object = dist.all_gather_object
    all_gather_object(out, obj)
    return out


def share_span(rows_total: int, row_pitch: int, itemsize: int, world: int, rank: int) -> tuple[int, int, int, int]:
    """(row0, rows, byte offset, bytes) of the rank's share of a replicated row-major operand: contiguous row blocks
    that tile the stored buffer exactly (same partition law as shard_rows)"""
    row0, rows = shard_rows(rows_total, world, rank)
    return row0, rows, row0 * row_pitch * itemsize, rows * row_pitch * itemsize


def map_on_peers(pipeline, t: Tensor, rank: int, world: int, all_gather_object=None):
    """(device pointer of every rank's copy of `t` as seen from this rank, the list of mappings to close): the ranks
    exchange CUDA-IPC handles of their own buffer and open the peers'"""
    lib = capi.lib()
    opened = []
    ptrs = (C.c_void_p * world)()
    if world == 1:
        ptrs[0] = t.buffer
        return ptrs, opened
    handle = (C.c_ubyte * 64)()
    capi.check(lib.wk_ipc_get_handle(pipeline.q, t.ptr, handle))
    handles = exchange(bytes(handle), world, all_gather_object)
    for r, h in enumerate(handles):
        if r == rank:
            ptrs[r] = t.buffer
            continue
        buf = (C.c_ubyte * 64).from_buffer_copy(h)
        p = C.c_void_p()
        capi.check(lib.wk_ipc_open_handle(pipeline.q, buf, C.byref(p)))
        ptrs[r] = p.value
        opened.append(p)
    return ptrs, opened


class ReplicatedOperand:
    """B of the row-sharded product: every rank holds all of it in HBM, but when it comes from HOST memory (identical
    in every process) each rank uploads only its 1/world share of the stored rows over PCIe and pushes that share into
    the peers' buffers over NVLink (copy engines, CUDA-IPC mapped pointers) -- world x less host traffic than every
    rank uploading the whole matrix.  The caller orders the steps: `upload_share` on an upload queue, `push_share`
    on a queue that waits for it, then a host-side barrier after the push queue has drained (every share has landed
    everywhere) before the product reads B, and another one before the next upload overwrites it.

    col_panels > 1 keeps B in HBM as that many dense column panels [rows, cols / col_panels], one after the other
    (`panels[j]` are the tensors to multiply by), so that panel j can be uploaded, pushed, fenced and multiplied while
    the later panels are still on their way; the share of a panel is again the rank's block of its rows."""

    def __init__(self, context, pipeline, b: Tensor, rank: int, world: int, all_gather_object=None, col_panels: int = 1):
        self.context, self.b, self.rank, self.world, self.col_panels = context, b, rank, world, col_panels
        rows_total, cols = b.shape
        es = b.dtype.itemsize
        if col_panels < 1 or cols % col_panels or (col_panels > 1 and (b.row_pitch != cols or (cols // col_panels) % 2)):
            raise capi.InvalidValue("InvalidValue: column panels need a dense B and an even number of columns per panel")
        nc = cols // col_panels
        pitch = b.row_pitch if col_panels == 1 else nc
        self.row0, self.rows, off, self.share_bytes = share_span(rows_total, pitch, es, world, rank)
        pipeline.wait_and_cleanup()
        self.peer_ptrs, self._opened = map_on_peers(pipeline, b, rank, world, all_gather_object)
        panel_bytes = rows_total * pitch * es
        self.offsets = [j * panel_bytes + off for j in range(col_panels)]
        self.panels = [b] if col_panels == 1 else [
            Tensor.wrap(context, pipeline, (rows_total, nc), b.dtype, b.buffer + j * panel_bytes) for j in range(col_panels)]
        self.shares = [Tensor.wrap(context, pipeline, (self.rows, nc), b.dtype, b.buffer + self.offsets[j], row_pitch=pitch)
                       for j in range(col_panels)]
        self.offset_bytes, self.share = self.offsets[0], self.shares[0]

    def upload_share(self, pipeline, host_rows: np.ndarray, panel: int = 0) -> None:
        """host_rows: this rank's rows [row0, row0 + rows) of column panel `panel` of B, dense"""
        from .tensor import memory

        if host_rows.shape != tuple(self.shares[panel].shape):
            raise capi.InvalidValue("InvalidValue: host share does not match the rank's rows of B")
        memory.read_from_buffer(pipeline, self.shares[panel], host_rows)

    def push_share(self, pipeline, panel: int = 0) -> None:
        """copy this rank's share of a panel into every peer's B at the same offset (NVLink, one copy per peer, rotated
        so that the ranks do not all target the same peer first)"""
        off = self.offsets[panel]
        for i in range(1, self.world):
            r = (self.rank + i) % self.world
            capi.check(capi.lib().wk_d2d(pipeline.q, self.peer_ptrs[r] + off, self.b.buffer + off, self.share_bytes))

    def release(self, pipeline) -> None:
        pipeline.wait_and_cleanup()
        for p in self._opened:
            capi.check(capi.lib().wk_ipc_close_handle(pipeline.q, p))
        self._opened = []


class RowShardedC:
    """The full-size result matrix of one rank plus the peers' copies, mapped through CUDA IPC."""

    def __init__(self, context, pipeline, m: int, n: int, dtype, rank: int, world: int, all_gather_object=None):
        self.context, self.rank, self.world = context, rank, world
        self.m, self.n = m, n
        self.row0, self.rows = shard_rows(m, world, rank)
        self.c = Tensor.alloc(context, pipeline, (m, n), dtype)
        pipeline.wait_and_cleanup()
        self.peer_ptrs, self._opened = map_on_peers(pipeline, self.c, rank, world, all_gather_object)

    def block(self, pipeline) -> Tensor:
        """view of this rank's row block of C"""
        off = self.row0 * self.c.row_pitch * self.c.dtype.itemsize
        return Tensor.wrap(self.context, pipeline, (self.rows, self.n), self.c.dtype, self.c.buffer + off)

    def gemm(self, pipeline, alpha, a_blk: Tensor, op_a: int, b: Tensor, op_b: int, beta) -> None:
        """this rank's rows of C = alpha*op(A)*op(B) + beta*C, written to every rank's C.  `a_blk` is the rank's block
        of A (see a_block).  Results are complete on all ranks after every rank's queue has finished + a barrier."""
        k = a_blk.shape[1 - op_a]
        rows = a_blk.shape[op_a]
        if rows != self.rows or b.shape[op_b] != k or b.shape[1 - op_b] != self.n or a_blk.dtype != self.c.dtype:
            raise capi.InvalidValue("InvalidValue: block shapes do not match the sharded C")
        _, pal = _scalar(self.c.dtype, alpha)
        _, pbe = _scalar(self.c.dtype, beta)
        capi.check(capi.lib().wk_gemm_rowshard_allgather(
            pipeline.q, self.c.type_index, op_a, op_b, self.rows, self.n, k, pal, a_blk.ptr, a_blk.row_pitch, b.ptr,
            b.row_pitch, pbe, self.row0, self.peer_ptrs, self.world, self.rank, self.c.row_pitch))

    def release(self, pipeline) -> None:
        pipeline.wait_and_cleanup()
        for p in self._opened:
            capi.check(capi.lib().wk_ipc_close_handle(pipeline.q, p))
        self._opened = []
        self.c.release(pipeline)


def reference_product(a: np.ndarray, op_a: int, b: np.ndarray, op_b: int, world: int) -> np.ndarray:
    """numpy model of the sharded product (used by the CPU tests of the partition logic): concatenation of the
    per-rank blocks computed from the a_block views"""
    A = a.T if op_a else a
    m, k = A.shape
    lda = a.shape[1]
    flat = a.reshape(-1)
    blocks = []
    for r in range(world):
        row0, rows = shard_rows(m, world, r)
        off, shape = a_block(op_a, row0, rows, k, lda)
        view = np.lib.stride_tricks.as_strided(flat[off:], shape=shape, strides=(lda * a.itemsize, a.itemsize))
        blocks.append((view.T if op_a else view) @ (b.T if op_b else b))
    return np.concatenate(blocks, axis=0)
