"""BASELINE config 4: f32 & f64 square GEMM sweep on one B200 -- N x {NN,NT,TN,TT} x {(null,null),(1.25,null),(0.75,0.5)},
inputs U[-1,1) from the reference PRNG (seeds 42/43/44), 3 warm-ups, timed with CUDA events on the queue's stream, L2 not
flushed (operands of N >= 4096 exceed it anyway).  Writes one JSON object per line and a markdown table.

    python tools/gemm_sweep.py [out_prefix] [N ...]       default N = 1024 2048 4096 8192 16384
"""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402

out_prefix = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep"
sizes = [int(x) for x in sys.argv[2:]] or [1024, 2048, 4096, 8192, 16384]
ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])
lib = wk.capi.lib()


def ev():
    e = C.c_void_p()
    wk.capi.check(lib.wk_event_record(pipe.q, C.byref(e)))
    return e


OPS = (("NN", 0, 0), ("NT", 0, 1), ("TN", 1, 0), ("TT", 1, 1))
SCALARS = ((None, None), (1.25, None), (0.75, 0.5))
rows = []
for name, dtype in (("f32", np.float32), ("f64", np.float64)):
    for n in sizes:
        a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
        wk.tensor.random.uniform(pipe, a, 42, -1, 1)
        wk.tensor.random.uniform(pipe, b, 43, -1, 1)
        wk.tensor.random.uniform(pipe, c, 44, -1, 1)
        sec_est = 2 * n ** 3 / (250e12 if name == "f32" else 33e12)
        reps = int(max(5, min(200, 0.4 / sec_est)))
        for opname, oa, ob in OPS:
            for alpha, beta in SCALARS:
                for _ in range(3):
                    wk.blas.gemm(pipe, alpha, a, oa, b, ob, beta, c)
                pipe.wait_and_cleanup()
                e0 = ev()
                for _ in range(reps):
                    wk.blas.gemm(pipe, alpha, a, oa, b, ob, beta, c)
                e1 = ev()
                lib.wk_event_wait(e1)
                ms = C.c_float()
                lib.wk_event_elapsed_ms(e0, e1, C.byref(ms))
                tf = 2 * n ** 3 * reps / (ms.value * 1e-3) / 1e12
                row = {"dtype": name, "N": n, "op": opname, "alpha": alpha, "beta": beta, "reps": reps,
                       "ms": ms.value / reps, "tflops": tf}
                rows.append(row)
                print(json.dumps(row), flush=True)
                if beta is not None:  # beta*C feeds back: keep C bounded for the next variant
                    wk.tensor.random.uniform(pipe, c, 44, -1, 1)
        for t in (a, b, c):
            t.release(pipe)

with open(out_prefix + ".jsonl", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
with open(out_prefix + ".md", "w") as f:
    f.write("| dtype | N | op | (alpha, beta) = (null,null) | (1.25,null) | (0.75,0.5) |\n|---|---|---|---|---|---|\n")
    for name in ("f32", "f64"):
        for n in sizes:
            for opname, _, _ in OPS:
                cells = [r for r in rows if r["dtype"] == name and r["N"] == n and r["op"] == opname]
                f.write(f"| {name} | {n} | {opname} | " + " | ".join(f"{r['tflops']:.1f} TF/s ({r['ms']:.3f} ms)" for r in cells) + " |\n")
ctx.deinit()
