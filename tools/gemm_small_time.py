"""kernel time of small GEMMs without host launch overhead: REPS gemm calls captured into one CUDA graph, the graph timed with
events.   python tools/gemm_small_time.py f32|f64 M N K [M N K ...]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402

dtype = np.float32 if sys.argv[1] == "f32" else np.float64
ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])
L = wk.capi.lib()
REPS = 20


def ev():
    e = C.c_void_p()
    wk.capi.check(L.wk_event_record(pipe.q, C.byref(e)))
    return e


dims = [int(v) for v in sys.argv[2:]]
for i in range(0, len(dims), 3):
    M, N, K = dims[i:i + 3]
    a, b, c = (wk.Tensor.alloc(ctx, pipe, s, dtype) for s in ((M, K), (K, N), (M, N)))
    wk.tensor.random.uniform(pipe, a, 42, -1, 1)
    wk.tensor.random.uniform(pipe, b, 43, -1, 1)
    for _ in range(2):  # warm-up: workspaces are allocated outside the capture
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    pipe.wait_and_cleanup()
    pipe.begin_capture()
    for _ in range(REPS):
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    g = pipe.end_capture()
    g.launch(pipe)
    pipe.wait_and_cleanup()
    e0 = ev()
    for _ in range(5):
        g.launch(pipe)
    e1 = ev()
    L.wk_event_wait(e1)
    ms = C.c_float()
    L.wk_event_elapsed_ms(e0, e1, C.byref(ms))
    us = ms.value * 1e3 / (5 * REPS)
    print(f"{sys.argv[1]} {M}x{N}x{K}: {us:.1f} us/gemm  {2.0 * M * N * K / (us * 1e-6) / 1e12:.1f} TFLOP/s", flush=True)
    g.release()
    for t in (a, b, c):
        t.release(pipe)
