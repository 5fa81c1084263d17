"""time blas.gemm at given sizes: python tools/gemm_time.py f32|f64|c64|c128 N [N ...]   (prints TFLOP/s, CUDA events;
complex products are counted as 8*N^3 real flops = 4 real multiply-adds per complex one)"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402

dtype = {"f32": np.float32, "f64": np.float64, "c64": np.complex64, "c128": np.complex128}[sys.argv[1]]
FLOPS = 8 if sys.argv[1].startswith("c") else 2
ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])


def ev():
    e = C.c_void_p()
    wk.capi.check(wk.capi.lib().wk_event_record(pipe.q, C.byref(e)))
    return e


for n in [int(x) for x in sys.argv[2:]]:
    a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
    wk.tensor.random.uniform(pipe, a, 42, -1, 1)
    wk.tensor.random.uniform(pipe, b, 43, -1, 1)
    for _ in range(3):
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    pipe.wait_and_cleanup()
    reps = 10 if n <= 8192 else (5 if n <= 16384 else 3)
    e0 = ev()
    for _ in range(reps):
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    e1 = ev()
    wk.capi.lib().wk_event_wait(e1)
    ms = C.c_float()
    wk.capi.lib().wk_event_elapsed_ms(e0, e1, C.byref(ms))
    print(f"N={n}: {FLOPS * n ** 3 * reps / (ms.value * 1e-3) / 1e12:.1f} TFLOP/s ({ms.value / reps:.2f} ms)", flush=True)
    for t in (a, b, c):
        t.release(pipe)
