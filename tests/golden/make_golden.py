"""Generates the committed golden fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference (kython28/wekua) cannot be built or run in this container (no zig, no OpenCL device, zig-opencl v0.8.1 not
vendored -- SURVEY 8c), so its outputs cannot be captured directly.  Two kinds of fixtures are committed instead:

  ref_kat.npz         the reference's OWN known-answer tests materialised as arrays: operands and the result the
                      reference asserts, from the closed forms in src/blas/test_helpers.zig:46-74,149-164,231-246
                      (gemm A*I / I*B with alpha/beta), src/blas/axpy.zig:178-502 (basic / alpha = null / alpha = -1), src/math/
                      trig.zig:129-187 (sin).  Nothing of this file depends on
                      the oracle: it is what a run of the reference's test-suite checks.
  oracle_vectors.npz  seeded random problems pushed through the CPU restatement (oracle/), which is itself pinned on
                      ref_kat (tests/test_oracle_kat*.py).  These freeze the restatement's outputs -- integer GEMM for
                      every dtype / transpose pair, float GEMM, the streaming kernels, the nn kernels, the PRNG stream of
                      uniform.cl -- so that (a) the oracle cannot drift silently and (b) the CUDA path is compared with
                      committed bytes, not only with a checker built in the same run.

The oracle device is the record our CUDA CommandQueue reports (vector width 1): layouts in the fixtures are the ones the
CUDA tensors use.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402
from tests import ref_cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REAL = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]


def rand(rng, dtype, shape, lo=-1.0, hi=1.0):
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return rng.uniform(lo, hi, size=shape).astype(dtype)
    info = np.iinfo(dt)
    return rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)


def ref_kat():
    """the reference's closed-form test cases as arrays"""
    out = {}
    for dtype in (np.float32, np.float64):
        name = np.dtype(dtype).name
        for i, (kind, m, x, op_a, op_b, packed, alpha, beta) in enumerate(ref_cases.GEMM_UNPACKED + ref_cases.GEMM_PACKED):
            a_shape, b_shape, c_shape = ref_cases.gemm_case_shapes(kind, m, x, op_a, op_b)
            exp = ref_cases.gemm_case_expected(kind, m, x, op_a, op_b, alpha, beta, np, dtype)
            # operands per test_helpers.zig:107-147: the data matrix holds 1..n row-major in ITS stored shape, the other
            # operand is the identity, C starts as all ones
            data_shape = a_shape if kind == "AI" else b_shape
            data = (np.arange(int(np.prod(data_shape))) + 1).reshape(data_shape).astype(dtype)
            ident = np.eye((b_shape if kind == "AI" else a_shape)[0], dtype=dtype)
            key = f"gemm/{name}/{i}"
            out[key + "/a"] = data if kind == "AI" else ident
            out[key + "/b"] = ident if kind == "AI" else data
            # C is filled with ones only when beta is given (test_helpers.zig:128-130), else left zero from alloc
            out[key + "/c0"] = (np.ones if beta is not None else np.zeros)(c_shape, dtype=dtype)
            out[key + "/meta"] = np.array([op_a, op_b, -1 if alpha is None else alpha, -1 if beta is None else beta, int(packed)])
            out[key + "/expected"] = np.asarray(exp, dtype=dtype)
    # axpy, every real dtype (src/blas/axpy.zig:178-283 basic, :285-385 alpha = null, :387-502 alpha = -1 signed only)
    for dtype in REAL:
        name = np.dtype(dtype).name
        i5, i4 = np.arange(5), np.arange(4)
        cases = [("basic", i5 + 1, (i5 + 1) * 10, 2, (i5 + 1) * 12), ("null", i4 + 1, i4 + 5, None, 2 * i4 + 6)]
        if np.dtype(dtype).kind != "u":
            cases.append(("minus_one", i4 + 1, np.full(4, 10), -1, 9 - i4))
        for cname, x, y, alpha, exp in cases:
            key = f"axpy/{name}/{cname}"
            out[key + "/x"], out[key + "/y0"], out[key + "/expected"] = x.astype(dtype), y.astype(dtype), exp.astype(dtype)
            out[key + "/alpha"] = np.array([np.nan if alpha is None else alpha])
    # sin (src/math/trig.zig:129-187): sin(0, pi/6, pi/2, pi) = (0, 0.5, 1, 0) within 1e-5 absolute
    out["trig/sin/x"] = np.array([0.0, np.pi / 6, np.pi / 2, np.pi])
    out["trig/sin/expected"] = np.array([0.0, 0.5, 1.0, 0.0])
    return out


def oracle_vectors():
    dev = o.device("b200")
    rng = np.random.default_rng(20261017)
    out = {}

    def T(dtype, shape, data=None):
        t = o.OTensor(dev, dtype, shape)
        if data is not None:
            t.read_from(np.ascontiguousarray(data, dtype=dtype))
        return t

    # GEMM: every real dtype x 4 transpose pairs x 3 (alpha, beta) variants on one odd shape, small-magnitude floats
    M, N, K = 37, 29, 41
    for dtype in REAL:
        name = np.dtype(dtype).name
        for op_a in (0, 1):
            for op_b in (0, 1):
                a = rand(rng, dtype, (K, M) if op_a else (M, K))
                b = rand(rng, dtype, (N, K) if op_b else (K, N))
                c0 = rand(rng, dtype, (M, N))
                for vi, (alpha, beta) in enumerate(((None, None), (3 if np.dtype(dtype).kind != "f" else 1.25, None),
                                                    (2 if np.dtype(dtype).kind != "f" else 0.75, 5 if np.dtype(dtype).kind != "f" else 0.5))):
                    oa, ob, oc = T(dtype, a.shape, a), T(dtype, b.shape, b), T(dtype, (M, N), c0)
                    o.gemm(alpha, oa, op_a, ob, op_b, beta, oc)
                    key = f"gemm/{name}/{op_a}{op_b}/{vi}"
                    if vi == 0:
                        out[f"gemm/{name}/{op_a}{op_b}/a"], out[f"gemm/{name}/{op_a}{op_b}/b"] = a, b
                        out[f"gemm/{name}/{op_a}{op_b}/c0"] = c0
                    out[key + "/scalars"] = np.array([np.nan if alpha is None else alpha, np.nan if beta is None else beta])
                    out[key + "/c"] = oc.to_host()
    # streaming kernels on a padded shape (5 x 7 -> pitch 8, 6 padded rows)
    shape = (5, 7)
    for dtype in REAL:
        name = np.dtype(dtype).name
        x, y = rand(rng, dtype, shape), rand(rng, dtype, shape)
        ox, oy = T(dtype, shape, x), T(dtype, shape, y)
        o.axpy(ox, 3, oy)
        out[f"axpy/{name}/x"], out[f"axpy/{name}/y0"], out[f"axpy/{name}/y"] = x, y, oy.to_host()
        ox, oy = T(dtype, shape, x), T(dtype, shape, y)
        o.hadamard(ox, oy)
        out[f"hadamard/{name}/x"] = ox.to_host()
        out[f"sum/{name}"] = np.array([o.tsum(T(dtype, shape, x))])
        tr = T(dtype, shape[::-1])
        o.transpose(tr, T(dtype, shape, x), 0, 1)
        out[f"transpose/{name}"] = tr.to_host()
        u = T(dtype, (4, 6))
        u.uniform(42)
        out[f"uniform/{name}/seed42"] = u.to_host()
        if np.dtype(dtype).kind == "f":
            u = T(dtype, (4, 6))
            u.uniform(43, -1, 1)
            out[f"uniform/{name}/seed43_pm1"] = u.to_host()
    # nn kernels and optimizers (f32 / f64): IEEE-only arithmetic => the CUDA path must reproduce these bits
    for dtype in (np.float32, np.float64):
        name = np.dtype(dtype).name
        shape = (6, 10)
        outp = rand(rng, dtype, shape, 0.05, 0.95)
        exp = rand(rng, dtype, shape, 0, 1)
        oo, oe, oerr, odev = T(dtype, shape, outp), T(dtype, shape, exp), T(dtype, shape), T(dtype, shape)
        o.mse(oo, oe, oerr, odev)
        out[f"nn/{name}/output"], out[f"nn/{name}/expected"] = outp, exp
        out[f"nn/{name}/mse_err"], out[f"nn/{name}/mse_dev"] = oerr.to_host(), odev.to_host()
        od = T(dtype, shape)
        o.sigmoid_dev(oo, od)
        out[f"nn/{name}/sigmoid_dev"] = od.to_host()
        o.tanh_dev(oo, od)
        out[f"nn/{name}/tanh_dev"] = od.to_host()
        b = rand(rng, dtype, (shape[1],))
        ob, oo2 = T(dtype, (shape[1],), b), T(dtype, shape, outp)
        o.bias(oo2, ob)
        out[f"nn/{name}/bias"], out[f"nn/{name}/bias_added"] = b, oo2.to_host()
        obg = T(dtype, (shape[1],))
        o.bias_step(oo, obg)
        out[f"nn/{name}/bias_step"] = obg.to_host()
        sg = T(dtype, shape, rand(rng, dtype, shape, -4, 4))
        out[f"nn/{name}/sigmoid_in"] = sg.to_host()
        o.unary(sg, "sigmoid")
        out[f"nn/{name}/sigmoid_out"] = sg.to_host()
        x0, g, h0 = rand(rng, dtype, shape), rand(rng, dtype, shape), rand(rng, dtype, shape, 0, 1)
        out[f"opt/{name}/x0"], out[f"opt/{name}/g"], out[f"opt/{name}/h0"] = x0, g, h0
        for opt in ("gdm", "adagrad", "rmsprop"):
            ox, og, oh = T(dtype, shape, x0), T(dtype, shape, g), T(dtype, shape, h0)
            for _ in range(3):
                if opt == "gdm":
                    o.gdm(ox, og, oh, dtype(0.01), dtype(0.9))
                elif opt == "adagrad":
                    o.adagrad(ox, og, oh, dtype(0.01))
                else:
                    o.rmsprop(ox, og, oh, dtype(0.01), dtype(0.9))
            out[f"opt/{name}/{opt}/x"], out[f"opt/{name}/{opt}/h"] = ox.to_host(), oh.to_host()
    return out


if __name__ == "__main__":
    kat = ref_kat()
    np.savez_compressed(os.path.join(HERE, "ref_kat.npz"), **kat)
    vec = oracle_vectors()
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **vec)
    for f in ("ref_kat.npz", "oracle_vectors.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print(len(kat), "+", len(vec), "arrays")
