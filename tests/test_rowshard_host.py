"""Host logic of the row-sharded multi-GPU GEMM (wekua_b200/rowshard.py) on CPU: the partition, the A-block views for
both transpose modes, and the rank-ordered handle exchange over a world_size-2 gloo group (no GPU, no compute calls
into the CUDA library)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rowshard():
    from wekua_b200 import rowshard

    return rowshard


@pytest.mark.parametrize("m,world", [(8, 1), (8, 2), (10, 4), (32768, 8), (1000, 3), (7, 7)])
def test_shard_rows_tiles_the_rows_exactly(m, world):
    rs = _rowshard()
    nxt = 0
    for r in range(world):
        row0, rows = rs.shard_rows(m, world, r)
        assert row0 == nxt and rows >= 1
        nxt = row0 + rows
    assert nxt == m


def test_shard_rows_rejects_bad_arguments():
    rs = _rowshard()
    from wekua_b200 import capi

    for args in ((4, 0, 0), (4, 2, 2), (4, 2, -1), (1, 2, 0)):
        with pytest.raises(capi.InvalidValue):
            rs.shard_rows(*args)


@pytest.mark.parametrize("op_a", [0, 1])
@pytest.mark.parametrize("op_b", [0, 1])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_blockwise_product_equals_full_product(op_a, op_b, world):
    rs = _rowshard()
    rng = np.random.default_rng(5)
    m, n, k = 24, 10, 6
    a = rng.integers(-5, 5, (k, m) if op_a else (m, k)).astype(np.float64)
    b = rng.integers(-5, 5, (n, k) if op_b else (k, n)).astype(np.float64)
    full = (a.T if op_a else a) @ (b.T if op_b else b)
    np.testing.assert_array_equal(rs.reference_product(a, op_a, b, op_b, world), full)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from wekua_b200 import rowshard as rs

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # the 64-byte IPC handles travel as bytes objects, in rank order, to every rank
        handle = bytes([rank]) * 64
        got = rs.exchange(handle, world)
        ok = got == [bytes([r]) * 64 for r in range(world)]
        # every rank derives the same partition; the union of the blocks is the whole matrix
        m = 1000
        blocks = rs.exchange(rs.shard_rows(m, world, rank), world)
        ok &= blocks[0][0] == 0 and all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        ok &= blocks[-1][0] + blocks[-1][1] == m
        # each rank multiplies its own block; gathering the blocks reproduces the full product
        rng = np.random.default_rng(11)
        a = rng.integers(-3, 3, (m, 16)).astype(np.float64)
        b = rng.integers(-3, 3, (16, 8)).astype(np.float64)
        row0, rows = rs.shard_rows(m, world, rank)
        parts = rs.exchange(a[row0:row0 + rows] @ b, world)
        ok &= np.array_equal(np.concatenate(parts, axis=0), a @ b)
        # replicated operand from host memory: every rank contributes only its byte span of B; the spans tile the stored
        # buffer (padded pitch included) and writing every rank's span at its offset reassembles B on every rank
        pitch, es = 10, 8
        stored = np.zeros((16, pitch))
        stored[:, :8] = b
        flat = stored.reshape(-1).view(np.uint8)
        row0_b, rows_b, off, nbytes = rs.share_span(16, pitch, es, world, rank)
        spans = rs.exchange((off, nbytes, bytes(flat[off:off + nbytes])), world)
        mine = np.zeros_like(flat)
        for o_, n_, data in spans:
            mine[o_:o_ + n_] = np.frombuffer(data, dtype=np.uint8)
        ok &= spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        ok &= spans[-1][0] + spans[-1][1] == flat.size and np.array_equal(mine, flat)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_exchange_and_partition_over_gloo_world2():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert results == [(0, True), (1, True)]
