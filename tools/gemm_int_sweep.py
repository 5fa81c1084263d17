"""Integer GEMM sweep (8 of the 10 real dtypes of src/blas/gemm.zig:834-874; kernels gemm_nxn_gpu.cl:82-319 in the reference):
dtype x N x {NN,NT,TN,TT}, inputs = the reference PRNG over the whole integer range (seeds 42/43), CUDA events on the queue's
stream.  Tera-ops/s = 2 N^3 / t.  Also measures the float problems the tensor-core loaders cannot address (f32 with
cols % 4 == 2: row pitch not a multiple of 16 bytes) against their aligned neighbours.

    python tools/gemm_int_sweep.py [out_prefix] [N ...]        default N = 4096 8192
"""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import wekua_b200 as wk  # noqa: E402

out_prefix = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep_int"
sizes = [int(x) for x in sys.argv[2:]] or [4096, 8192]
ctx = wk.Context.init([0])
pipe = wk.Pipeline.init(ctx.command_queues[0])
lib = wk.capi.lib()


def ev():
    e = C.c_void_p()
    wk.capi.check(lib.wk_event_record(pipe.q, C.byref(e)))
    return e


def time_gemm(a, oa, b, ob, c, alpha=None, beta=None, reps=5, warm=2):
    for _ in range(warm):
        wk.blas.gemm(pipe, alpha, a, oa, b, ob, beta, c)
    pipe.wait_and_cleanup()
    e0 = ev()
    for _ in range(reps):
        wk.blas.gemm(pipe, alpha, a, oa, b, ob, beta, c)
    e1 = ev()
    lib.wk_event_wait(e1)
    ms = C.c_float()
    lib.wk_event_elapsed_ms(e0, e1, C.byref(ms))
    for e in (e0, e1):
        lib.wk_event_release(e)
    return ms.value / reps


OPS = (("NN", 0, 0), ("NT", 0, 1), ("TN", 1, 0), ("TT", 1, 1))
rows = []
for dtype in (np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64):
    name = np.dtype(dtype).name
    for n in sizes:
        a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
        wk.tensor.random.uniform(pipe, a, 42)
        wk.tensor.random.uniform(pipe, b, 43)
        for opname, oa, ob in OPS:
            ms = time_gemm(a, oa, b, ob, c, reps=3 if n >= 8192 else 6)
            row = {"dtype": name, "N": n, "op": opname, "ms": ms, "tops": 2 * n ** 3 / (ms * 1e-3) / 1e12}
            rows.append(row)
            print(json.dumps(row), flush=True)
        for t in (a, b, c):
            t.release(pipe)

# the unaligned-pitch cliff: f32 N = 4098 (pitch 4098 floats = 16392 B, not a multiple of 16) vs 4096 / 4100
cliff = []
for dtype, n in ((np.float32, 4096), (np.float32, 4098), (np.float32, 4100), (np.float64, 4098)):
    a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
    wk.tensor.random.uniform(pipe, a, 42, -1, 1)
    wk.tensor.random.uniform(pipe, b, 43, -1, 1)
    for opname, oa, ob in (OPS[0], OPS[1]):
        ms = time_gemm(a, oa, b, ob, c, reps=5)
        row = {"dtype": np.dtype(dtype).name, "N": n, "row_pitch": a.row_pitch, "op": opname, "ms": ms,
               "tflops": 2 * n ** 3 / (ms * 1e-3) / 1e12}
        cliff.append(row)
        print(json.dumps(row), flush=True)
    for t in (a, b, c):
        t.release(pipe)

with open(out_prefix + ".jsonl", "w") as f:
    for r in rows + cliff:
        f.write(json.dumps(r) + "\n")
with open(out_prefix + ".md", "w") as f:
    f.write("| dtype | N | NN | NT | TN | TT |\n|---|---|---|---|---|---|\n")
    for dtype in dict.fromkeys(r["dtype"] for r in rows):
        for n in sizes:
            cells = [r for r in rows if r["dtype"] == dtype and r["N"] == n]
            f.write(f"| {dtype} | {n} | " + " | ".join(f"{r['tops']:.1f} Top/s ({r['ms']:.2f} ms)" for r in cells) + " |\n")
    f.write("\n| dtype | N | row pitch (elements) | NN | NT |\n|---|---|---|---|---|\n")
    for key in dict.fromkeys((r["dtype"], r["N"], r["row_pitch"]) for r in cliff):
        cells = [r for r in cliff if (r["dtype"], r["N"], r["row_pitch"]) == key]
        f.write(f"| {key[0]} | {key[1]} | {key[2]} | " + " | ".join(f"{r['tflops']:.1f} TF/s ({r['ms']:.2f} ms)" for r in cells) + " |\n")
ctx.deinit()
