"""The reference's src/core tests (context constructors incl. DevicesArrayEmpty context.zig:203-243, device-info
sanity command_queue.zig:236-423, Pipeline ordering pipeline.zig:75-321) against the CUDA-backed Context / CommandQueue /
Pipeline.  The constructor-error case needs no device; the rest is -m gpu."""
import numpy as np
import pytest

import wekua_b200 as wk


def test_context_init_with_empty_device_list_is_an_error():
    """context.zig:203-214: an empty device array -> error.DevicesArrayEmpty, before anything touches the driver"""
    with pytest.raises(wk.capi.DevicesArrayEmpty):
        wk.Context.init([])
    with pytest.raises(wk.capi.DevicesArrayEmpty):
        wk.Context.init_from_device_type("accelerator")


@pytest.mark.gpu
def test_context_constructors_and_queue_info():
    """context.zig:216-243 + command_queue.zig:236-423: every constructor yields >= 1 queue whose capability record is sane"""
    for ctx in (wk.Context.init([0]), wk.Context.init_from_device_type("all"), wk.Context.init_from_device_type("gpu"),
                wk.Context.init_from_best_device()):
        assert len(ctx.command_queues) >= 1
        for i, q in enumerate(ctx.command_queues):
            assert q.wekua_id == i and q.context is ctx
            assert q.compute_units > 0 and q.max_work_group_size == 1024
            assert q.local_mem_type == "local" and q.local_mem_size >= 48 * 1024
            assert q.cache_line_size == 128 and q.global_mem_size > (1 << 30)
            assert len(q.vector_widths) == 10 and all(1 <= w <= 16 for w in q.vector_widths)  # the <= 16 clamp
            assert q.compute_capability[0] == 10 and "B200" in q.device_name
            for t in wk.core.SUPPORTED_TYPES:
                assert q.is_type_supported(t)
            assert not q.is_type_supported(np.float16)
        ctx.deinit()
        assert ctx.command_queues == []
    two = wk.Context.init([0, 0])  # two queues on one device (what bench.py's upload / compute / download overlap uses)
    assert [q.wekua_id for q in two.command_queues] == [0, 1]
    two.deinit()


@pytest.mark.gpu
def test_pipeline_orders_work_across_queues_through_events():
    """pipeline.zig:35-61: work enqueued on queue B after wait_for(event of queue A) sees A's results; wait_and_cleanup
    drains the queue and drops the recorded events"""
    ctx = wk.Context.init([0, 0])
    pa, pb = (wk.Pipeline.init(q) for q in ctx.command_queues)
    n = 1 << 22
    x = wk.Tensor.alloc(ctx, pa, (n,), np.float32)
    y = wk.Tensor.alloc(ctx, pa, (n,), np.float32)
    pa.wait_and_cleanup()
    for _ in range(20):  # a chain long enough that an unordered reader would see a partial sum
        wk.tensor.fill.constant(pa, x, 1)
        for _ in range(8):
            wk.blas.axpy(pa, x, 1.0, x)  # x doubles: 1 -> 256
        ev = pa.record_event()
        pb.wait_for(ev)
        wk.blas.axpy(pb, x, 1.0, y)      # y += 256, on the OTHER queue
        pa.wait_for(pb.record_event())   # the next round's fill must not overtake the reader
    assert len(pa._events) == 20 and len(pb._events) == 20
    pb.wait_and_cleanup()
    pa.wait_and_cleanup()
    assert pa._events == [] and pb._events == []
    got = wk.tensor.memory.to_numpy(pa, y)
    assert np.all(got == 20 * 256.0)
    for t in (x, y):
        t.release(pa)
    ctx.deinit()
