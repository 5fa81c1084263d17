"""one Linear layer step (forward + backward, linear.zig:480-678) at GEMM-bound size: unfused launches vs fused epilogues"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402

ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])
B, n_in, n_out = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 8192, 8192)))
dtype = np.float32


def ev():
    e = C.c_void_p()
    wk.capi.check(wk.capi.lib().wk_event_record(pipe.q, C.byref(e)))
    return e


for fused in (False, True):
    lin = wk.nn.Linear.init(ctx, pipe, n_in, n_out, wk.nn.Sigmoid.init(), dtype=dtype, seed=42, fused=fused)
    x = wk.Tensor.alloc(ctx, pipe, (B, n_in), dtype)
    xs = wk.Tensor.alloc(ctx, pipe, (B, n_in), dtype)
    wk.tensor.random.uniform(pipe, x, 7, -1, 1)
    cache = lin.prepare_cache(pipe, B)
    for _ in range(3):
        lin.forward(pipe, x, cache)
        lin.backward(pipe, cache, x, xs)
    pipe.wait_and_cleanup()
    reps = 10
    l0 = wk.capi.launch_count()
    e0 = ev()
    for _ in range(reps):
        lin.forward(pipe, x, cache)
        lin.backward(pipe, cache, x, xs)
    e1 = ev()
    wk.capi.lib().wk_event_wait(e1)
    ms = C.c_float()
    wk.capi.lib().wk_event_elapsed_ms(e0, e1, C.byref(ms))
    flops = 3 * 2.0 * B * n_in * n_out
    print(f"Linear step B={B} {n_in}->{n_out} fused={fused}: {ms.value / reps:.3f} ms, {flops * reps / (ms.value * 1e-3) / 1e12:.1f} TFLOP/s, "
          f"{(wk.capi.launch_count() - l0) // reps} launches/step", flush=True)
    for t in (x, xs):
        t.release(pipe)
    lin.release_cache(pipe, cache)
    lin.deinit(pipe)
