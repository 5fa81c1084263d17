"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): the row-sharded GEMM with the all-gather fused into the epilogue
(wk_gemm_rowshard_allgather) across 2 processes, checked against the single-GPU product and the oracle's fp64 ideal."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import ctypes as C

    from wekua_b200 import capi

    n = C.c_int32(0)
    capi.lib().wk_device_count(C.byref(n))
    return n.value


def _worker(rank, world, port, dtype_name, op_a, op_b, q, mnk=((640, 512, 200),)):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import wekua_b200 as wk
    from wekua_b200 import rowshard as rs

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        dtype = np.dtype(dtype_name)
        ctx = wk.Context.init([rank])
        pipe = wk.Pipeline.init(ctx.command_queues[0])
        ok_all, worst = True, 0.0
        for (M, N, K) in mnk:
            rng = np.random.default_rng(3)  # same data on every rank
            ad = rng.uniform(-1, 1, (K, M) if op_a else (M, K)).astype(dtype)
            bd = rng.uniform(-1, 1, (N, K) if op_b else (K, N)).astype(dtype)
            cd = rng.uniform(-1, 1, (M, N)).astype(dtype)
            a_full = wk.Tensor.alloc(ctx, pipe, ad.shape, dtype)
            b = wk.Tensor.alloc(ctx, pipe, bd.shape, dtype)
            wk.tensor.memory.read_from_buffer(pipe, a_full, ad)
            wk.tensor.memory.read_from_buffer(pipe, b, bd)
            shard = rs.RowShardedC(ctx, pipe, M, N, dtype, rank, world)
            wk.tensor.memory.read_from_buffer(pipe, shard.c, cd)
            pipe.wait_and_cleanup()
            dist.barrier()
            off, shape = rs.a_block(op_a, shard.row0, shard.rows, K, a_full.row_pitch)
            a_blk = wk.Tensor.wrap(ctx, pipe, shape, dtype, a_full.buffer + off * dtype.itemsize, row_pitch=a_full.row_pitch)
            shard.gemm(pipe, 0.75, a_blk, op_a, b, op_b, 0.5)
            pipe.wait_and_cleanup()
            dist.barrier()  # every rank's stores into every C have landed
            got = wk.tensor.memory.to_numpy(pipe, shard.c).astype(np.float64)
            A = (ad.T if op_a else ad).astype(np.float64)
            B = (bd.T if op_b else bd).astype(np.float64)
            ideal = 0.75 * (A @ B) + 0.5 * cd
            bound = (8 * K + 16) * np.finfo(dtype).eps * (0.75 * np.abs(A) @ np.abs(B) + 0.5 * np.abs(cd))
            ok_all &= bool(np.all(np.abs(got - ideal) <= bound))
            worst = max(worst, float(np.abs(got - ideal).max()))
            dist.barrier()
            shard.release(pipe)
            for t_ in (a_full, b):
                t_.release(pipe)
        q.put((rank, ok_all, worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name", ["float32", "float64"])
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (1, 1)])
def test_rowsharded_gemm_fused_allgather_2gpu(dtype_name, op_a, op_b):
    # ragged tiles; partial 32-column chunks + split-K; a transposed A block the tensor-core path cannot address (f32)
    mnk = ((640, 512, 200), (600, 500, 2048), (1100, 1284, 96))
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dtype_name, op_a, op_b, q, mnk)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in results), results


def _worker_replicated(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import wekua_b200 as wk
    from wekua_b200 import rowshard as rs

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ctx = wk.Context.init([rank, rank])
        p_up, p_push = (wk.Pipeline.init(cq) for cq in ctx.command_queues)
        ok = True
        for shape, dtype in (((301, 130), np.float32), ((64, 7), np.float64)):  # ragged shares, a padded pitch (7 -> 8)
            full = np.random.default_rng(5).uniform(-1, 1, shape).astype(dtype)  # the same matrix in every process
            b = wk.Tensor.alloc(ctx, p_up, shape, dtype)
            rep = rs.ReplicatedOperand(ctx, p_up, b, rank, world)
            dist.barrier()
            rep.upload_share(p_up, np.ascontiguousarray(full[rep.row0:rep.row0 + rep.rows]))
            p_push.wait_for(p_up.record_event())
            rep.push_share(p_push)
            p_push.wait_and_cleanup()
            dist.barrier()  # every share has landed everywhere
            got = wk.tensor.memory.to_numpy(p_up, b)
            ok &= bool(np.array_equal(got, full))
            dist.barrier()
            rep.release(p_up)
            b.release(p_up)
        # column panels: B kept as dense [rows, cols / 2] panels one after the other; panel by panel upload + push
        full = np.random.default_rng(6).uniform(-1, 1, (66, 12)).astype(np.float32)
        b = wk.Tensor.alloc(ctx, p_up, full.shape, np.float32)
        rep = rs.ReplicatedOperand(ctx, p_up, b, rank, world, col_panels=2)
        dist.barrier()
        for j in range(2):
            rep.upload_share(p_up, np.ascontiguousarray(full[rep.row0:rep.row0 + rep.rows, j * 6:(j + 1) * 6]), j)
            p_push.wait_for(p_up.record_event())
            rep.push_share(p_push, j)
        p_push.wait_and_cleanup()
        dist.barrier()
        for j in range(2):
            ok &= bool(np.array_equal(wk.tensor.memory.to_numpy(p_up, rep.panels[j]), full[:, j * 6:(j + 1) * 6]))
        dist.barrier()
        rep.release(p_up)
        b.release(p_up)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_replicated_operand_shared_upload_2gpu():
    """B of the row-sharded product from HOST memory: each rank uploads its share of the rows and pushes it to the peer over
    NVLink; afterwards every rank holds the whole matrix, bit for bit"""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_replicated, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert results == [(0, True), (1, True)], results
