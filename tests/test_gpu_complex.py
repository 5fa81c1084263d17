"""-m gpu: Complex(T) (type ids 10-19, src/core/types.zig:74-83) on the CUDA path, through the C ABI.

First half: the reference's own complex test cases (it runs axpy / dot / sum / fill / identity over all 20
SUPPORTED_TYPES and GEMM / mean / trig over Complex(f32/f64), always with imag = 0 inputs).  Second half: genuinely
complex random inputs against the oracle -- bit-exact for complex integers (Gaussian integers mod 2^bits) and for
the float streaming kernels (fixed IEEE op sequence, COMPLEX_MUL's 3-multiplication order kept), K-scaled bound
against a complex128 ideal for float GEMM.
"""
import math

import numpy as np
import pytest

from tests import gpu_helpers as gh
from tests import ref_cases as rc

pytestmark = pytest.mark.gpu

REALS = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
FLOATS = [np.float32, np.float64]
SIGNED = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]


def _cx(base):
    return gh.wk().core.Complex(base)


def _tensor(shape, base, data=None):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    t = wk.Tensor.alloc(ctx, pipe, shape, _cx(base))
    if data is not None:
        wk.tensor.memory.read_from_buffer(pipe, t, data)
    return t


def _assert_real(arr, expected, base):
    np.testing.assert_array_equal(arr["re"], np.asarray(expected).astype(base))
    assert not arr["im"].any()


def _pairs(re, im, base):
    out = np.zeros(np.shape(re), dtype=_cx(base))
    out["re"], out["im"] = re, im
    return out


def _rand_cx(rng, base, shape):
    return _pairs(gh.rand_data(rng, base, shape), gh.rand_data(rng, base, shape), base)


def _c128(a):
    return a["re"].astype(np.float64) + 1j * a["im"].astype(np.float64)


# ------------------------------------------------------------------------------ the reference's complex tests
def _run_ref_gemm_case(base, case):
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    kind, m, x, op_a, op_b, packed, alpha, beta = case
    a_shape, b_shape, c_shape = rc.gemm_case_shapes(kind, m, x, op_a, op_b)
    a, b, c = (_tensor(s, base) for s in (a_shape, b_shape, c_shape))
    if kind == "AI":
        wk.tensor.memory.read_from_buffer(pipe, a, np.arange(a_shape[0] * a_shape[1]) + 1)  # makeDataValue: {i+1, 0}
        wk.tensor.identity(pipe, b)
    else:
        wk.tensor.identity(pipe, a)
        wk.tensor.memory.read_from_buffer(pipe, b, np.arange(b_shape[0] * b_shape[1]) + 1)
    if beta is not None:
        wk.tensor.fill.one(pipe, c)
    pt = wk.blas.PackedTensors.init(pipe, c, x if kind == "AI" else m, True) if packed else None
    wk.blas.gemm(pipe, None if alpha is None else (alpha, 0), a, op_a, b, op_b, None if beta is None else (beta, 0), c, pt)
    exp = rc.gemm_case_expected(kind, m, x, op_a, op_b, alpha, beta, np, np.dtype(base).type)
    _assert_real(gh.to_np(c), exp, base)  # expectEqualValue: exact on both components
    for t in (a, b, c):
        t.release(pipe)


@pytest.mark.parametrize("base", FLOATS)
@pytest.mark.parametrize("path", [0, 1])
def test_gemm_all_algorithms_complex(base, path):
    """gemm.zig:1008 / :1188 'gemm cpu|gpu - all algorithms, complex'; path 1 forces the SIMT back-end"""
    gh.wk().capi.lib().wk_gemm_set_path(path)
    try:
        for case in rc.GEMM_COMPLEX_CPU:
            _run_ref_gemm_case(base, case)
    finally:
        gh.wk().capi.lib().wk_gemm_set_path(0)


@pytest.mark.parametrize("base", FLOATS)
def test_gemm_all_algorithms_with_packing_complex(base):
    """gemm.zig:1096 / :1272"""
    for case in rc.GEMM_COMPLEX_PACKED:
        _run_ref_gemm_case(base, case)


@pytest.mark.parametrize("base", REALS)
def test_axpy_reference_cases(base):
    """axpy.zig:178 (alpha {2,0}), :285 (alpha null), :504 / :619 (2-D, 3-D, alpha {3,0}), :817 (alpha {0,0})"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    x, y = _tensor((5,), base, np.arange(1, 6)), _tensor((5,), base, np.arange(1, 6) * 10)
    wk.blas.axpy(pipe, x, (2, 0), y)
    _assert_real(gh.to_np(y), np.arange(1, 6) * 12, base)
    x, y = _tensor((4,), base, np.arange(1, 5)), _tensor((4,), base, np.arange(5, 9))
    wk.blas.axpy(pipe, x, None, y)
    _assert_real(gh.to_np(y), np.arange(1, 5) + np.arange(5, 9), base)
    for shape in [(2, 3), (2, 2, 2)]:
        n = int(np.prod(shape))
        x, y = _tensor(shape, base, np.arange(n)), _tensor(shape, base, np.arange(n) * 3)
        wk.blas.axpy(pipe, x, (3, 0), y)
        _assert_real(gh.to_np(y).reshape(-1), np.arange(n) * 6, base)
    x, y = _tensor((4,), base, np.arange(1, 5)), _tensor((4,), base, np.arange(1, 5) * 10)
    wk.blas.axpy(pipe, x, (0, 0), y)
    _assert_real(gh.to_np(y), np.arange(1, 5) * 10, base)


@pytest.mark.parametrize("base", SIGNED)
def test_axpy_alpha_minus_one(base):
    """axpy.zig:387 'axpy - with alpha = -1 (subtraction) for signed types': alpha = {-1, 0}"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    x, y = _tensor((5,), base, np.arange(1, 6)), _tensor((5,), base, np.arange(1, 6) * 10)
    wk.blas.axpy(pipe, x, (-1, 0), y)
    _assert_real(gh.to_np(y), np.arange(1, 6) * 9, base)


def test_axpy_complex_subtract_quirk():
    """SURVEY Q3 (axpy.zig:75-76,84): alpha = {-1,-1} is classified as 'subtract'; reproduced, like the oracle"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    x = _tensor((3,), np.float32, np.array([1 + 2j, 3 - 1j, -2 + 0.5j]))
    y = _tensor((3,), np.float32, np.array([10 + 10j, 20 + 20j, 30 + 30j]))
    wk.blas.axpy(pipe, x, (-1, -1), y)
    np.testing.assert_array_equal(_c128(gh.to_np(y)), np.array([9 + 8j, 17 + 21j, 32 + 29.5j]))


@pytest.mark.parametrize("base", REALS)
def test_hadamard_sum(base):
    """basic.zig:258 'dot - element-wise multiplication', :322 'sum - basic sum operation', complex branches"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    x, y = _tensor((4,), base, [1, 2, 3, 4]), _tensor((4,), base, [2, 3, 4, 5])
    wk.math.dot(pipe, x, y)
    _assert_real(gh.to_np(x), [2, 6, 12, 20], base)
    s = wk.math.sum(pipe, _tensor((5,), base, [1, 2, 3, 4, 5]))
    assert s["re"] == 15 and s["im"] == 0


@pytest.mark.parametrize("base", FLOATS)
def test_mean(base):
    """basic.zig:368, complex branch: mean({2,4,6,8}) = {5, 0}"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    m = wk.math.mean(pipe, _tensor((4,), base, [2, 4, 6, 8]))
    assert abs(float(m["re"]) - 5.0) < 1e-5 and abs(float(m["im"])) < 1e-5


@pytest.mark.parametrize("base", FLOATS)
@pytest.mark.parametrize("op", ["sin", "cos", "tan", "sinh", "cosh", "tanh"])
def test_trig(base, op):
    """trig.zig:129-446, complex branches: real inputs with imag = 0, `.real` asserted to 1e-5"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    pts = {"tan": [0.0, math.pi / 6, math.pi / 4]}.get(op, [0.0, math.pi / 6, math.pi / 2, math.pi])
    if op in ("sinh", "cosh", "tanh"):
        pts = [0.0, 0.5, 1.0, 2.0]
    x = _tensor((len(pts),), base, pts)
    getattr(wk.math, op)(pipe, x)
    np.testing.assert_allclose(gh.to_np(x)["re"], [getattr(math, op)(p) for p in pts], atol=1e-5, rtol=0)


@pytest.mark.parametrize("base", REALS)
def test_fill_identity_transpose(base):
    """fill.zig / identity.zig / transpose.zig tests over SUPPORTED_TYPES, complex branches (identity = {1, 0})"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    t = _tensor((3, 5), base)
    wk.tensor.fill.constant(pipe, t, (7, 2))
    r = gh.to_np(t)
    assert (r["re"] == 7).all() and (r["im"] == 2).all()
    assert gh.padded(t)["re"].astype(np.int64).sum() == 7 * 15  # padding untouched
    i3 = _tensor((4, 4, 4), base)
    wk.tensor.identity(pipe, i3)
    exp = np.zeros((4, 4, 4))
    for d in range(4):
        exp[d, d, d] = 1
    _assert_real(gh.to_np(i3), exp, base)
    src = _tensor((3, 5), base, _pairs(np.arange(15), np.arange(15) + 100, base))
    dst = _tensor((5, 3), base)
    wk.tensor.transpose(pipe, dst, src, 0, 1)
    r = gh.to_np(dst)
    np.testing.assert_array_equal(r["re"], np.arange(15).reshape(3, 5).T.astype(base))
    np.testing.assert_array_equal(r["im"], (np.arange(15) + 100).reshape(3, 5).T.astype(base))


# ------------------------------------------------------------------------------ random complex inputs vs the oracle
def _pair(oracle, base, shape, data):
    t = _tensor(shape, base, data)
    o = oracle.OTensor(gh.oracle_dev(oracle), oracle.cx(base), shape).read_from(data)
    assert (t.row_pitch, t.slice_pitch, t.number_of_elements) == (o.layout.row_pitch, o.layout.slice_pitch, o.layout.number_of_elements)
    return t, o


def _same(a, b):
    """bit-exact on both components (NaN-free inputs)"""
    np.testing.assert_array_equal(a["re"], b["re"])
    np.testing.assert_array_equal(a["im"], b["im"])


@pytest.mark.parametrize("base", REALS)
@pytest.mark.parametrize("shape", [(7,), (5, 9), (3, 4, 6), (2, 3, 33, 17)])
def test_streaming_bit_exact_vs_oracle(oracle, base, shape):
    """axpy (three variants), Hadamard and transpose-free fill on random complex data: bit-exact for every base type"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(11)
    is_f = np.dtype(base).kind == "f"
    alphas = [None, (-1, 0), (0.75, -1.25) if is_f else (3, 5), (2, 0)]
    if np.dtype(base).kind == "u":
        alphas[1] = (1, 0)
    for alpha in alphas:
        xd, yd = _rand_cx(rng, base, shape), _rand_cx(rng, base, shape)
        x, ox = _pair(oracle, base, shape, xd)
        y, oy = _pair(oracle, base, shape, yd)
        wk.blas.axpy(pipe, x, alpha, y)
        oracle.axpy(ox, alpha, oy)
        _same(gh.to_np(y), oy.to_host())
        wk.math.dot(pipe, x, y)
        oracle.hadamard(ox, oy)
        _same(gh.to_np(x), ox.to_host())
        _same(gh.padded(x), ox.buf)  # padding stays as the reference leaves it


@pytest.mark.parametrize("base", REALS)
def test_sum_mean_vs_oracle(oracle, base):
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(12)
    is_f = np.dtype(base).kind == "f"
    for shape in [(9,), (6, 11), (3, 5, 8)]:
        xd = _rand_cx(rng, base, shape)
        if not is_f:  # keep the mean's truncating division away from overflow-dependent results only for i8/u8 wrap
            pass
        x, ox = _pair(oracle, base, shape, xd)
        s, so = wk.math.sum(pipe, x), oracle.tsum(ox)
        if is_f:
            tol = 8 * xd.size * np.finfo(base).eps
            assert abs(float(s["re"]) - float(so["re"])) <= tol and abs(float(s["im"]) - float(so["im"])) <= tol
        else:
            assert s["re"] == so["re"] and s["im"] == so["im"]
    if is_f:
        xd = _rand_cx(rng, base, (7, 6))
        x, ox = _pair(oracle, base, (7, 6), xd)
        m, mo = wk.math.mean(pipe, x), oracle.mean(ox)
        assert abs(float(m["re"]) - float(mo["re"])) <= 1e-5 and abs(float(m["im"]) - float(mo["im"])) <= 1e-5


@pytest.mark.parametrize("base", FLOATS)
@pytest.mark.parametrize("op", ["sin", "cos", "tan", "sinh", "cosh", "tanh"])
def test_trig_random_vs_oracle(oracle, base, op):
    """complex arguments: CUDA libm vs glibc through the same trig.cl formulas (abs 1e-5 like the reference, plus a
    relative term because sinh/cosh grow)"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(13)
    xd = _pairs(rng.uniform(-1.2, 1.2, (9, 14)), rng.uniform(-1.2, 1.2, (9, 14)), base)
    x, ox = _pair(oracle, base, (9, 14), xd)
    getattr(wk.math, op)(pipe, x)
    oracle.unary(ox, op)
    got, ref = _c128(gh.to_np(x)), _c128(ox.to_host())
    np.testing.assert_allclose(got, ref, rtol=1e-5 if base == np.float32 else 1e-12, atol=1e-5 if base == np.float32 else 1e-12)
    z = _c128(xd)
    np.testing.assert_allclose(got, getattr(np, op)(z), rtol=2e-5 if base == np.float32 else 1e-11, atol=1e-5)


@pytest.mark.parametrize("base", REALS)
def test_uniform_vs_oracle(oracle, base):
    """uniform.cl:80-93: the components hash (index << 1) and (index << 1) + 1 of the PADDED index -- bit-exact"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    for shape in [(64, 100), (5, 7, 9)]:
        t, o = _pair(oracle, base, shape, np.zeros(shape))
        wk.tensor.random.uniform(pipe, t, 42)
        o.uniform(42)
        _same(gh.to_np(t), o.to_host())
        lo, hi = (-5, 5) if np.dtype(base).kind != "u" else (10, 100)
        wk.tensor.random.uniform(pipe, t, 43, lo, hi)
        o.uniform(43, lo, hi)
        got, ref = gh.to_np(t), o.to_host()
        if np.dtype(base).kind == "f":
            # min + normalized*range is evaluated in double then rounded to T on both sides; glibc and CUDA agree
            # bit for bit on IEEE double mul/add, so this is exact too
            _same(got, ref)
        else:
            _same(got, ref)
        for comp in ("re", "im"):
            assert got[comp].min() >= lo and got[comp].max() <= hi


@pytest.mark.parametrize("base", REALS)
def test_transpose_nd_vs_oracle(oracle, base):
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(14)
    shape = (3, 4, 5)
    for d0, d1 in [(0, 1), (0, 2), (1, 2)]:
        xd = _rand_cx(rng, base, shape)
        rshape = list(shape)
        rshape[d0], rshape[d1] = rshape[d1], rshape[d0]
        x, ox = _pair(oracle, base, shape, xd)
        r, orr = _pair(oracle, base, tuple(rshape), np.zeros(rshape))
        wk.tensor.transpose(pipe, r, x, d0, d1)
        oracle.transpose(orr, ox, d0, d1)
        _same(gh.to_np(r), orr.to_host())
        _same(gh.to_np(r), np.swapaxes(xd, d0, d1))


GEMM_SHAPES = [(1, 1, 1), (2, 3, 5), (17, 33, 9), (64, 64, 64), (100, 65, 70), (129, 130, 65)]


@pytest.mark.parametrize("base", [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64])
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_complex_integers_bit_exact(oracle, base, op_a, op_b):
    """Gaussian integers mod 2^bits: the one-real-GEMM formulation is the same ring element as the reference's
    3-multiplication tile loops, whatever alpha / beta -- bit-exact against the restated kernels"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(21 + 2 * op_a + op_b)
    for (M, N, K) in GEMM_SHAPES[:5]:
        for alpha, beta in [(None, None), ((3, 0), None), ((2, 5), (7, 3)), (None, (1, 2)), ((1, 1), (4, 0))]:
            a_shape = (K, M) if op_a else (M, K)
            b_shape = (N, K) if op_b else (K, N)
            a, oa = _pair(oracle, base, a_shape, _rand_cx(rng, base, a_shape))
            b, ob = _pair(oracle, base, b_shape, _rand_cx(rng, base, b_shape))
            c, oc = _pair(oracle, base, (M, N), _rand_cx(rng, base, (M, N)))
            wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
            oracle.gemm(alpha, oa, op_a, ob, op_b, beta, oc)
            _same(gh.to_np(c), oc.to_host())
            for t in (a, b, c):
                t.release(pipe)


@pytest.mark.parametrize("base", FLOATS)
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("path", [0, 1])
def test_gemm_complex_floats_vs_ideal(oracle, base, op_a, op_b, path):
    """random complex operands: |c - c_ideal| <= (8*2K + 32) eps (|alpha| |A||B| + |beta||C|) componentwise, where
    |X| = |re| + |im|; c_ideal in complex128.  The restated reference kernels are held to the same bound (their
    3-multiplication product cancels, so they use most of it; the tensor-core path uses ~1/10)."""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(31 + 2 * op_a + op_b)
    eps = float(np.finfo(base).eps)
    wk.capi.lib().wk_gemm_set_path(path)
    try:
        for (M, N, K) in GEMM_SHAPES:
            for alpha, beta in [(None, None), ((1.25, 0), None), ((0.75, -0.5), (0.25, 1.5)), (None, (2, 0))]:
                a_shape = (K, M) if op_a else (M, K)
                b_shape = (N, K) if op_b else (K, N)
                ad, bd, cd = (_rand_cx(rng, base, s) for s in (a_shape, b_shape, (M, N)))
                a, oa = _pair(oracle, base, a_shape, ad)
                b, ob = _pair(oracle, base, b_shape, bd)
                c, oc = _pair(oracle, base, (M, N), cd)
                wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
                oracle.gemm(alpha, oa, op_a, ob, op_b, beta, oc)
                A, B, C0 = _c128(ad), _c128(bd), _c128(cd)
                A, B = (A.T if op_a else A), (B.T if op_b else B)
                al = complex(*alpha) if alpha is not None else 1.0
                be = complex(*beta) if beta is not None else 0.0
                ideal = al * (A @ B) + be * C0
                mag = lambda X: np.abs(X.real) + np.abs(X.imag)  # noqa: E731
                bound = (8 * 2 * K + 32) * eps * (mag(np.array(al)) * (mag(A) @ mag(B)) + mag(np.array(be)) * mag(C0)) + 1e-300
                got, ref = _c128(gh.to_np(c)), _c128(oc.to_host())
                for name, val in (("cuda", got), ("oracle", ref)):
                    err = np.maximum(np.abs((val - ideal).real), np.abs((val - ideal).imag))
                    assert np.all(err <= bound), f"{name} c{np.dtype(base).name} {M}x{N}x{K} worst {np.max(err / bound)}"
                for t in (a, b, c):
                    t.release(pipe)
    finally:
        wk.capi.lib().wk_gemm_set_path(0)


def test_gemm_complex_tensor_core_scale():
    """Complex(f32) 512 x 384 x 640 through the tcgen05 path (real 512 x 768 x 1280) against complex128"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(41)
    M, N, K = 512, 384, 640
    for base in FLOATS:
        ad, bd = _rand_cx(rng, base, (M, K)), _rand_cx(rng, base, (N, K))
        a, b, c = _tensor((M, K), base, ad), _tensor((N, K), base, bd), _tensor((M, N), base)
        wk.capi.lib().wk_gemm_set_path(2)  # fail rather than fall back if the layout were not eligible
        try:
            wk.blas.gemm(pipe, (0.5, 0.25), a, 0, b, 1, None, c)
        finally:
            wk.capi.lib().wk_gemm_set_path(0)
        ideal = (0.5 + 0.25j) * (_c128(ad) @ _c128(bd).T)
        err = np.abs(_c128(gh.to_np(c)) - ideal).max()
        assert err <= (2 * K + 32) * np.finfo(base).eps * 4, err
        for t in (a, b, c):
            t.release(pipe)
