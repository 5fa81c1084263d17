"""Quick GPU check of the tensor-core GEMM paths against numpy fp64 (run under gpurun with a timeout)."""
import sys
import time
import ctypes as C

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402

dtype = np.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else np.float64
ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])
rng = np.random.default_rng(0)
eps = np.finfo(dtype).eps
wk.capi.lib().wk_gemm_set_path(2)  # tensor-core path or fail
worst = 0.0
shapes = [(128, 256, 32), (128, 256, 64), (256, 512, 128), (200, 300, 100), (64, 128, 64), (1000, 1000, 1000), (129, 260, 36),
          (4, 8, 4), (2048, 1024, 512)]
for (M, N, K) in shapes:
    for op_a in (0, 1):
        for op_b in (0, 1):
            for alpha, beta in ((None, None), (1.25, None), (0.75, 0.5)):
                a_shape = (K, M) if op_a else (M, K)
                b_shape = (N, K) if op_b else (K, N)
                ad = rng.uniform(-1, 1, a_shape).astype(dtype)
                bd = rng.uniform(-1, 1, b_shape).astype(dtype)
                cd = rng.uniform(-1, 1, (M, N)).astype(dtype)
                a, b, c = (wk.Tensor.alloc(ctx, pipe, s, dtype) for s in (a_shape, b_shape, (M, N)))
                for t, d in ((a, ad), (b, bd), (c, cd)):
                    wk.tensor.memory.read_from_buffer(pipe, t, d)
                try:
                    wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
                except wk.capi.WekuaError as e:
                    print(f"{M}x{N}x{K} op{op_a}{op_b}: not eligible ({e})")
                    for t in (a, b, c):
                        t.release(pipe)
                    break
                got = wk.tensor.memory.to_numpy(pipe, c).astype(np.float64)
                A = (ad.T if op_a else ad).astype(np.float64)
                B = (bd.T if op_b else bd).astype(np.float64)
                ideal = A @ B
                bound = np.abs(A) @ np.abs(B)
                if alpha is not None:
                    ideal *= float(dtype(alpha)); bound *= abs(alpha)
                if beta is not None:
                    ideal += float(dtype(beta)) * cd; bound += abs(beta) * np.abs(cd)
                err = np.abs(got - ideal)
                ratio = float((err / ((K + 16) * eps * bound + 1e-300)).max())
                rel = float(err.max() / np.abs(ideal).max())
                worst = max(worst, ratio)
                flag = "OK " if ratio <= 1 else "BAD"
                if ratio > 1 or (alpha is None and op_a == 0 and op_b == 0):
                    print(f"{flag} {M}x{N}x{K} op{op_a}{op_b} a={alpha} b={beta}: err/bound {ratio:.3g} maxrel {rel:.3g}")
                for t in (a, b, c):
                    t.release(pipe)
print("worst err/bound", worst)

# exactness on the reference's A*I pattern (integers up to 8192 must survive the hi/lo split)
M, K = 64, 128
ad = (np.arange(M * K) + 1).astype(dtype).reshape(M, K)
a, i, c = wk.Tensor.alloc(ctx, pipe, (M, K), dtype), wk.Tensor.alloc(ctx, pipe, (K, K), dtype), wk.Tensor.alloc(ctx, pipe, (M, K), dtype)
wk.tensor.memory.read_from_buffer(pipe, a, ad)
wk.tensor.identity(pipe, i)
wk.blas.gemm(pipe, None, a, 0, i, 0, None, c)
print("A*I exact:", np.array_equal(wk.tensor.memory.to_numpy(pipe, c), ad))

# timing
def ev():
    e = C.c_void_p(); wk.capi.check(wk.capi.lib().wk_event_record(pipe.q, C.byref(e))); return e
for n in ((8192,) if "quick" in sys.argv else (4096, 8192, 16384)):
    for op_a, op_b in ((0, 0), (0, 1), (1, 0), (1, 1)):
        a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
        wk.tensor.random.uniform(pipe, a, 42, -1, 1); wk.tensor.random.uniform(pipe, b, 43, -1, 1)
        for _ in range(2):
            wk.blas.gemm(pipe, None, a, op_a, b, op_b, None, c)
        pipe.wait_and_cleanup()
        reps = 5 if n < 16384 else 3
        e0 = ev()
        for _ in range(reps):
            wk.blas.gemm(pipe, None, a, op_a, b, op_b, None, c)
        e1 = ev()
        wk.capi.lib().wk_event_wait(e1)
        ms = C.c_float(); wk.capi.lib().wk_event_elapsed_ms(e0, e1, C.byref(ms))
        print(f"N={n} op{op_a}{op_b}: {2*n**3*reps/(ms.value*1e-3)/1e12:.1f} TFLOP/s ({ms.value/reps:.2f} ms)")
        for t in (a, b, c):
            t.release(pipe)
        if n == 16384 and (op_a, op_b) == (0, 1):
            break
