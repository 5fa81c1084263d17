"""HBM roofline of every streaming kernel of the path (SURVEY 8(a) rows a5-a13, 8(d) algorithmic bytes):
   python tools/stream_sweep.py OUT_PREFIX [log2_n]
Times each op through the C ABI on device-resident buffers (CUDA events on the queue's stream, 3 warm-ups, buffers
larger than L2 so nothing is re-read from cache) and writes OUT_PREFIX.jsonl / OUT_PREFIX.md with GB/s =
algorithmic bytes / time and the fraction of the measured copy bandwidth (MEASURED_PEAKS.json)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wekua_b200 as wk  # noqa: E402
from wekua_b200 import capi  # noqa: E402



def sweep(log2n=27, only=None, ctx=None, pipe=None, reps=None, warm=None, verbose=True):
    """rows (dicts) for every streaming op (or those named in `only`, comma-separated) at n = 2^log2n elements"""
    own_ctx = ctx is None
    n = 1 << log2n
    peak = 6553.3
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)

    L = capi.lib()
    if own_ctx:
        ctx = wk.Context.init([0])
        pipe = wk.Pipeline.init(ctx.command_queues[0])
    q = pipe.q


    def ev():
        e = C.c_void_p()
        capi.check(L.wk_event_record(q, C.byref(e)))
        return e


    reps_d = reps if reps is not None else int(os.environ.get('WK_SWEEP_REPS', 20))
    warm_d = warm if warm is not None else int(os.environ.get('WK_SWEEP_WARM', 3))

    def timed(fn, reps=reps_d, warm=warm_d, prep=None):
        """prep (re-initialise the in/out operand) runs untimed before every rep; each rep is then bracketed on its own"""
        for _ in range(warm):
            if prep:
                prep()
            fn()
        capi.check(L.wk_queue_finish(q))
        if prep is None:
            e0 = ev()
            for _ in range(reps):
                fn()
            e1 = ev()
            pairs = [(e0, e1)]
        else:
            pairs = []
            for _ in range(reps):
                prep()
                e0 = ev()
                fn()
                pairs.append((e0, ev()))
        capi.check(L.wk_queue_finish(q))
        total = 0.0
        for e0, e1 in pairs:
            ms = C.c_float()
            L.wk_event_elapsed_ms(e0, e1, C.byref(ms))
            total += ms.value
            L.wk_event_release(e0)
            L.wk_event_release(e1)
        return total / reps


    def sc(dtype, v):
        a = np.array([v], dtype=dtype)
        return a, a.ctypes.data_as(C.c_void_p)


    rows_out = []
    UNARY = {"sin", "cos", "tan", "sinh", "cosh", "tanh", "sigmoid"}
    for dtype, tid in ((np.float32, 8), (np.float64, 9)):
        s = np.dtype(dtype).itemsize
        name = "f32" if tid == 8 else "f64"
        bufs = [wk.Tensor.alloc(ctx, pipe, (2, n // 2), dtype) for _ in range(4)]  # dense [2, n/2]: no padding
        for i, t in enumerate(bufs):
            wk.tensor.random.uniform(pipe, t, 42 + i, 0.1, 0.9)
        x, y, z, w = (t.ptr for t in bufs)
        R, Cc = 2, n // 2
        keep = []

        def S(v):
            a, p = sc(dtype, v)
            keep.append(a)
            return p

        alpha, lr, beta, gamma, b1, b2, eps = S(0.999), S(1e-3), S(0.9), S(0.9), S(0.9), S(0.999), S(1e-8)
        one_over = S(1.0)
        side = 1 << (log2n // 2)
        tr_rows, tr_cols = side, n // side
        host = np.zeros(2, dtype=dtype)
        hp = host.ctypes.data_as(C.c_void_p)
        bias_cols = 4096
        ops = [
            ("axpy", 3, lambda: L.wk_axpy(q, tid, 1, R, Cc, alpha, x, Cc, n, y, Cc, n)),
            ("axpy(alpha=null)", 3, lambda: L.wk_axpy(q, tid, 1, R, Cc, None, x, Cc, n, y, Cc, n)),
            ("scal", 2, lambda: L.wk_scal(q, tid, 1, R, Cc, alpha, x, Cc, n)),
            ("hadamard (math.dot)", 3, lambda: L.wk_hadamard(q, tid, 1, R, Cc, x, Cc, n, y, Cc, n)),
            ("sum", 1, lambda: L.wk_sum(q, tid, 1, R, Cc, n, x, hp)),  # blocking, like math.sum: one host round trip per call
            ("sum (device scalar)", 1, lambda: L.wk_sum_async(q, tid, 1, R, Cc, n, x, w)),  # same kernel, scalar left in HBM
            ("dot_reduce", 2, lambda: L.wk_dot_reduce(q, tid, 1, R, Cc, x, Cc, n, y, Cc, n, hp)),
            ("sin", 2, lambda: L.wk_unary(q, tid, 0, z, n)),
            ("cos", 2, lambda: L.wk_unary(q, tid, 1, z, n)),
            ("tan", 2, lambda: L.wk_unary(q, tid, 2, z, n)),
            ("sinh", 2, lambda: L.wk_unary(q, tid, 3, z, n)),
            ("cosh", 2, lambda: L.wk_unary(q, tid, 4, z, n)),
            ("tanh", 2, lambda: L.wk_unary(q, tid, 5, z, n)),
            ("sigmoid", 2, lambda: L.wk_unary(q, tid, 6, z, n)),
            ("sigmoid_dev", 2, lambda: L.wk_sigmoid_dev(q, tid, x, z, n)),
            ("tanh_dev", 2, lambda: L.wk_tanh_dev(q, tid, x, z, n)),
            ("act_backward(sigmoid)", 3, lambda: L.wk_act_backward(q, tid, 1, x, None, z, n)),
            ("bias_add", 2, lambda: L.wk_bias_add(q, tid, z, y, bias_cols, n)),
            ("bias_step", 1, lambda: L.wk_bias_step(q, tid, x, z, bias_cols, n // bias_cols, bias_cols)),
            ("mse (err only)", 3, lambda: L.wk_mse(q, tid, x, y, z, None, n)),
            ("mse (+dev)", 4, lambda: L.wk_mse(q, tid, x, y, z, w, n)),
            ("gdm", 5, lambda: L.wk_gdm(q, tid, x, y, z, lr, beta, n)),
            ("adagrad", 5, lambda: L.wk_adagrad(q, tid, x, y, z, lr, n)),
            ("rmsprop", 5, lambda: L.wk_rmsprop(q, tid, x, y, z, lr, gamma, n)),
            ("adam", 7, lambda: L.wk_adam(q, tid, x, y, z, w, lr, b1, b2, eps, 3, n)),
            ("fill", 1, lambda: L.wk_fill(q, tid, 1, R, Cc, z, Cc, n, alpha)),
            ("uniform", 1, lambda: L.wk_uniform(q, tid, 1, R, Cc, z, Cc, n, 42, None, None)),
            ("transpose2d", 2, lambda: L.wk_transpose2d(q, tid, tr_rows, tr_cols, x, tr_cols, z, tr_rows)),
            ("memory.copy (d2d)", 2, lambda: L.wk_d2d(q, z, x, n * s)),
        ]
        for op, k, fn in ops:
            if only and not any(o == op for o in only.split(',')):
                continue
            for i, t in enumerate(bufs):  # fresh U[0.1, 0.9) operands: repeated in-place ops must not drift into inf/denormals
                wk.tensor.random.uniform(pipe, t, 42 + i, 0.1, 0.9)

            def call(fn=fn):
                capi.check(fn())
            # expanding in-place maps (tan, sinh, cosh ...) leave U[0.1, 0.9) after a few applications: re-draw z before each
            prep = (lambda: wk.tensor.random.uniform(pipe, bufs[2], 44, 0.1, 0.9)) if op in UNARY else None
            ms = timed(call, prep=prep)
            gbs = k * n * s / (ms * 1e-3) / 1e9
            row = {"op": op, "dtype": name, "n": n, "alg_bytes": k * n * s, "bytes_per_elem": f"{k}*s", "ms": ms, "gbs": gbs,
                   "frac_of_copy_peak": gbs / peak}
            rows_out.append(row)
            if verbose:
                print(f"{name} {op:24s} {ms:8.4f} ms  {gbs:8.1f} GB/s  {gbs / peak:.3f}", flush=True)
        for t in bufs:
            t.release(pipe)

    if own_ctx:
        ctx.deinit()
    return rows_out, peak


def main():
    out_prefix = sys.argv[1]
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 27
    rows_out, peak = sweep(log2n, os.environ.get("WK_SWEEP_ONLY"))
    with open(out_prefix + ".jsonl", "w") as f:
        for r in rows_out:
            f.write(json.dumps(r) + "\n")
    with open(out_prefix + ".md", "w") as f:
        f.write(f"Streaming kernels, n = 2^{log2n} elements per operand, device-resident, CUDA events, 20 reps after 3 warm-ups; "
                f"peak = measured copy bandwidth {peak:.1f} GB/s.\n\n")
        f.write("| op | algorithmic bytes | f32 ms | f32 GB/s | f32 frac | f64 ms | f64 GB/s | f64 frac |\n|---|---|---|---|---|---|---|---|\n")
        by = {}
        for r in rows_out:
            by.setdefault(r["op"], {})[r["dtype"]] = r
        for op, d in by.items():
            a, b = d.get("f32"), d.get("f64")
            f.write(f"| {op} | {a['bytes_per_elem']}*N | {a['ms']:.4f} | {a['gbs']:.0f} | {a['frac_of_copy_peak']:.3f} | "
                    f"{b['ms']:.4f} | {b['gbs']:.0f} | {b['frac_of_copy_peak']:.3f} |\n")


if __name__ == "__main__":
    main()
