"""Pins the complex half (type ids 10-19, src/core/types.zig:74-83) of the CPU restatement against the reference's
own known-answer tests.  The reference runs its axpy / dot / sum / fill / identity tests over ALL 20 SUPPORTED_TYPES
and its GEMM / mean / trig tests over the complex FLOAT types, always with imag = 0 inputs; those are transcribed
here.  The tests with genuinely complex inputs at the end are NOT reference tests (numpy complex128 / exact integer
arithmetic is the yardstick there)."""
import math

import numpy as np
import pytest

from tests import ref_cases as rc

REALS = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
FLOATS = [np.float32, np.float64]
SIGNED = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]
DEVICES = [("cpu", 16), ("cpu", 8), ("gpu", 1), ("b200", 1)]


def _dev(oracle, spec):
    return oracle.device(spec[0], spec[1])


def _re_im(a):
    return a["re"], a["im"]


def _assert_real(arr, expected, base):
    re, im = _re_im(arr)
    np.testing.assert_array_equal(re, np.asarray(expected).astype(base))
    assert not im.any()


def _run_gemm_case(oracle, dev, base, case):
    kind, m, x, op_a, op_b, packed, alpha, beta = case
    dt = oracle.cx(base)
    a_shape, b_shape, c_shape = rc.gemm_case_shapes(kind, m, x, op_a, op_b)
    a, b, c = (oracle.OTensor(dev, dt, s) for s in (a_shape, b_shape, c_shape))
    if kind == "AI":
        a.read_from(np.arange(a_shape[0] * a_shape[1]) + 1)  # makeDataValue: {i+1, 0}
        b.identity()
    else:
        a.identity()
        b.read_from(np.arange(b_shape[0] * b_shape[1]) + 1)
    if beta is not None:
        c.fill((1, 0))  # fill.one
    oracle.gemm(None if alpha is None else (alpha, 0), a, op_a, b, op_b, None if beta is None else (beta, 0), c, packed=packed)
    exp = rc.gemm_case_expected(kind, m, x, op_a, op_b, alpha, beta, np, np.dtype(base).type)
    _assert_real(c.to_host(), exp, base)  # expectEqualValue: exact on both components


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("base", FLOATS)
def test_gemm_all_algorithms_complex(oracle, devspec, base):
    """gemm.zig:1008 'gemm cpu - all algorithms, complex' / :1188 'gemm gpu - all algorithms, complex'"""
    dev = _dev(oracle, devspec)
    for case in (rc.GEMM_COMPLEX_GPU if devspec[0] != "cpu" else rc.GEMM_COMPLEX_CPU):
        _run_gemm_case(oracle, dev, base, case)


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("base", FLOATS)
def test_gemm_all_algorithms_with_packing_complex(oracle, devspec, base):
    """gemm.zig:1096 / :1272 'gemm cpu|gpu - all algorithms with packing, complex'"""
    dev = _dev(oracle, devspec)
    for case in rc.GEMM_COMPLEX_PACKED:
        _run_gemm_case(oracle, dev, base, case)


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("base", REALS)
def test_axpy_basic(oracle, devspec, base):
    """axpy.zig:178 'axpy - basic operation y = alpha*x + y for 1D tensor', complex branch: alpha = {2, 0}"""
    dev, dt = _dev(oracle, devspec), oracle.cx(base)
    x = oracle.OTensor(dev, dt, (5,)).read_from(np.arange(1, 6))
    y = oracle.OTensor(dev, dt, (5,)).read_from(np.arange(1, 6) * 10)
    oracle.axpy(x, (2, 0), y)
    _assert_real(y.to_host(), np.arange(1, 6) * 12, base)


@pytest.mark.parametrize("base", REALS)
def test_axpy_alpha_null(oracle, base):
    """axpy.zig:285 'axpy - with alpha = null (direct sum) for all types'"""
    dev, dt = oracle.device("cpu", 16), oracle.cx(base)
    x = oracle.OTensor(dev, dt, (4,)).read_from(np.arange(1, 5))
    y = oracle.OTensor(dev, dt, (4,)).read_from(np.arange(5, 9))
    oracle.axpy(x, None, y)
    _assert_real(y.to_host(), np.arange(1, 5) + np.arange(5, 9), base)


@pytest.mark.parametrize("base", SIGNED)
def test_axpy_alpha_minus_one(oracle, base):
    """axpy.zig:387 'axpy - with alpha = -1 (subtraction) for signed types': alpha = {-1, 0}"""
    dev, dt = oracle.device("cpu", 16), oracle.cx(base)
    x = oracle.OTensor(dev, dt, (5,)).read_from(np.arange(1, 6))
    y = oracle.OTensor(dev, dt, (5,)).read_from(np.arange(1, 6) * 10)
    oracle.axpy(x, (-1, 0), y)
    _assert_real(y.to_host(), np.arange(1, 6) * 9, base)


@pytest.mark.parametrize("base", REALS)
@pytest.mark.parametrize("shape", [(2, 3), (2, 2, 2)])
def test_axpy_nd(oracle, base, shape):
    """axpy.zig:504 '2D tensor' (alpha = {3,0}, x=i, y=3i) and :619 '3D tensor'"""
    dev, dt = oracle.device("gpu"), oracle.cx(base)
    n = int(np.prod(shape))
    x = oracle.OTensor(dev, dt, shape).read_from(np.arange(n))
    y = oracle.OTensor(dev, dt, shape).read_from(np.arange(n) * 3)
    oracle.axpy(x, (3, 0), y)
    _assert_real(y.to_host().reshape(-1), np.arange(n) * 6, base)


@pytest.mark.parametrize("base", REALS)
def test_axpy_zero_alpha(oracle, base):
    """axpy.zig:817 'axpy - zero alpha'"""
    dev, dt = oracle.device("cpu", 8), oracle.cx(base)
    x = oracle.OTensor(dev, dt, (4,)).read_from(np.arange(1, 5))
    y = oracle.OTensor(dev, dt, (4,)).read_from(np.arange(1, 5) * 10)
    oracle.axpy(x, (0, 0), y)
    _assert_real(y.to_host(), np.arange(1, 5) * 10, base)


@pytest.mark.parametrize("base", REALS)
def test_hadamard_sum(oracle, base):
    """basic.zig:258 'dot - element-wise multiplication' and :322 'sum - basic sum operation', complex branches"""
    dev, dt = oracle.device("cpu", 16), oracle.cx(base)
    x = oracle.OTensor(dev, dt, (4,)).read_from([1, 2, 3, 4])
    y = oracle.OTensor(dev, dt, (4,)).read_from([2, 3, 4, 5])
    oracle.hadamard(x, y)
    _assert_real(x.to_host(), [2, 6, 12, 20], base)
    s = oracle.tsum(oracle.OTensor(dev, dt, (5,)).read_from([1, 2, 3, 4, 5]))
    assert s["re"] == 15 and s["im"] == 0


@pytest.mark.parametrize("base", FLOATS)
def test_mean(oracle, base):
    """basic.zig:368 'mean - basic mean operation for float types', complex branch: mean({2,4,6,8}) = {5, 0}"""
    dev, dt = oracle.device("gpu"), oracle.cx(base)
    m = oracle.mean(oracle.OTensor(dev, dt, (4,)).read_from([2, 4, 6, 8]))
    assert abs(float(m["re"]) - 5.0) < 1e-5 and abs(float(m["im"])) < 1e-5


@pytest.mark.parametrize("base", FLOATS)
@pytest.mark.parametrize("op", ["sin", "cos", "tan", "sinh", "cosh", "tanh"])
def test_trig(oracle, base, op):
    """trig.zig:129-446, complex branches: real inputs with imag = 0, only `.real` is asserted (abs 1e-5)"""
    dev, dt = oracle.device("cpu", 16), oracle.cx(base)
    pts = {"tan": [0.0, math.pi / 6, math.pi / 4]}.get(op, [0.0, math.pi / 6, math.pi / 2, math.pi])
    if op in ("sinh", "cosh", "tanh"):
        pts = [0.0, 0.5, 1.0, 2.0]
    x = oracle.OTensor(dev, dt, (len(pts),)).read_from(pts)
    oracle.unary(x, op)
    np.testing.assert_allclose(x.to_host()["re"], [getattr(math, op)(p) for p in pts], atol=1e-5, rtol=0)


@pytest.mark.parametrize("base", REALS)
def test_fill_identity_transpose(oracle, base):
    """fill.zig / identity.zig / transpose.zig tests over SUPPORTED_TYPES, complex branches (identity = {1, 0})"""
    dev, dt = oracle.device("cpu", 16), oracle.cx(base)
    t = oracle.OTensor(dev, dt, (3, 5)).fill((7, 2))
    r = t.to_host()
    assert (r["re"] == 7).all() and (r["im"] == 2).all()
    assert t.buf["re"].sum() == 7 * 15  # padding untouched
    i3 = oracle.OTensor(dev, dt, (4, 4, 4)).identity().to_host()
    exp = np.zeros((4, 4, 4))
    for d in range(4):
        exp[d, d, d] = 1
    _assert_real(i3, exp, base)
    src = oracle.OTensor(dev, dt, (3, 5)).read_from(oracle.cx_pairs(np.arange(15), np.arange(15) + 100, base))
    dst = oracle.OTensor(dev, dt, (5, 3))
    oracle.transpose(dst, src, 0, 1)
    r = dst.to_host()
    np.testing.assert_array_equal(r["re"], np.arange(15).reshape(3, 5).T.astype(base))
    np.testing.assert_array_equal(r["im"], (np.arange(15) + 100).reshape(3, 5).T.astype(base))


@pytest.mark.parametrize("base", REALS)
def test_uniform_properties(oracle, base):
    """random/uniform.zig tests, complex branch: both components drawn (hash of 2i and 2i+1, uniform.cl:80-93)"""
    dev, dt = oracle.device("cpu", 16), oracle.cx(base)
    a = oracle.OTensor(dev, dt, (64, 100)).uniform(42).to_host()
    b = oracle.OTensor(dev, dt, (64, 100)).uniform(42).to_host()
    np.testing.assert_array_equal(a, b)
    assert (a["re"] != a["im"]).mean() > 0.5
    # the real tensor of the same seed hashes index i; the complex one hashes 2i for its real part
    real = oracle.OTensor(dev, base, (1, 200)).uniform(42).to_host().reshape(-1)
    cplx = oracle.OTensor(dev, dt, (1, 100)).uniform(42).to_host().reshape(-1)
    np.testing.assert_array_equal(cplx["re"], real[0::2])
    np.testing.assert_array_equal(cplx["im"], real[1::2])
    lo, hi = (-5, 5) if np.dtype(base).kind != "u" else (10, 100)
    r = oracle.OTensor(dev, dt, (64, 100)).uniform(42, lo, hi).to_host()
    for comp in ("re", "im"):
        assert r[comp].min() >= lo and r[comp].max() <= hi


# ----------------------------------------------------------------------------------- NOT reference tests
@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("base", FLOATS)
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_random_complex_vs_numpy(oracle, devspec, base, op_a, op_b):
    """the 3-multiplication product (wekua_cl_lib.cl:648-653) is a complex product: random complex operands against
    numpy complex128 within a K-scaled bound, unpacked and packed"""
    dev, dt = _dev(oracle, devspec), oracle.cx(base)
    rng = np.random.default_rng(5)
    M, N, K = 12, 8, 20
    A = rng.uniform(-1, 1, (K, M) if op_a else (M, K)) + 1j * rng.uniform(-1, 1, (K, M) if op_a else (M, K))
    B = rng.uniform(-1, 1, (N, K) if op_b else (K, N)) + 1j * rng.uniform(-1, 1, (N, K) if op_b else (K, N))
    C0 = rng.uniform(-1, 1, (M, N)) + 1j * rng.uniform(-1, 1, (M, N))
    alpha, beta = 0.75 - 0.5j, 0.25 + 1.5j
    ideal = alpha * ((A.T if op_a else A) @ (B.T if op_b else B)) + beta * C0
    for packed in (False, True):
        a, b, c = (oracle.OTensor(dev, dt, X.shape).read_from(X) for X in (A, B, C0))
        oracle.gemm(alpha, a, op_a, b, op_b, beta, c, packed=packed)
        r = c.to_host()
        got = r["re"].astype(np.float64) + 1j * r["im"].astype(np.float64)
        assert np.abs(got - ideal).max() <= 16 * (K + 4) * np.finfo(base).eps * 4


@pytest.mark.parametrize("base", [np.int8, np.uint16, np.int32, np.uint64])
def test_gemm_complex_integers_exact(oracle, base):
    """complex integer GEMM is exact Gaussian-integer arithmetic mod 2^bits whatever the tile order (the reference
    compiles these variants but never runs them)"""
    rng = np.random.default_rng(9)
    M, N, K = 6, 10, 14
    bits = np.dtype(base).itemsize * 8
    Ar, Ai, Br, Bi = (rng.integers(-50, 50, s).astype(object) for s in ((M, K), (M, K), (K, N), (K, N)))
    re = Ar.dot(Br) - Ai.dot(Bi)
    im = Ar.dot(Bi) + Ai.dot(Br)

    def wrap(x):
        def w(v):
            v = int(v) % (1 << bits)
            return v - (1 << bits) if np.dtype(base).kind == "i" and v >= (1 << (bits - 1)) else v

        return np.array([np.dtype(base).type(w(v)) for v in np.asarray(x, dtype=object).reshape(-1)], dtype=base).reshape(np.shape(x))

    for devspec in DEVICES:
        dev, dt = _dev(oracle, devspec), oracle.cx(base)
        a = oracle.OTensor(dev, dt, (M, K)).read_from(oracle.cx_pairs(wrap(Ar), wrap(Ai), base))
        b = oracle.OTensor(dev, dt, (K, N)).read_from(oracle.cx_pairs(wrap(Br), wrap(Bi), base))
        for packed in (False, True):
            c = oracle.OTensor(dev, dt, (M, N))
            oracle.gemm(None, a, 0, b, 0, None, c, packed=packed)
            r = c.to_host()
            np.testing.assert_array_equal(r["re"], wrap(re))
            np.testing.assert_array_equal(r["im"], wrap(im))


def test_axpy_complex_subtract_quirk(oracle):
    """alpha = {-1, -1} is classified as 'subtract' by isSubstracting (axpy.zig:75-76,84), so the kernel
    computes y - x instead of y + (-1-1i)*x.  The restatement reproduces it."""
    dev, dt = oracle.device("cpu", 16), oracle.cx(np.float32)
    x = oracle.OTensor(dev, dt, (3,)).read_from(np.array([1 + 2j, 3 - 1j, -2 + 0.5j]))
    y = oracle.OTensor(dev, dt, (3,)).read_from(np.array([10 + 10j, 20 + 20j, 30 + 30j]))
    oracle.axpy(x, (-1, -1), y)
    r = y.to_host()
    np.testing.assert_array_equal(r["re"] + 1j * r["im"], np.array([9 + 8j, 17 + 21j, 32 + 29.5j]))
