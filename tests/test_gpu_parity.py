"""-m gpu: the CUDA path (through the C ABI) against the reference's own test cases and against the oracle.

Bar: bit-exact for integer dtypes and for every float kernel whose arithmetic is a fixed sequence of IEEE
add/mul/div/sqrt (axpy, hadamard, derivatives, mse, bias, optimizers); stated tolerance for GEMM
((K+16) eps sum|a||b|, gpu_helpers.gemm_float_bound) and for libm functions (4 ulp-ish rtol / 1e-5 abs as in the
reference's own trig tests).
"""
import math

import numpy as np
import pytest

from tests import gpu_helpers as gh
from tests import ref_cases as rc

pytestmark = pytest.mark.gpu

FLOATS = [np.float32, np.float64]
ALL = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
SIGNED = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]


# ------------------------------------------------------------------ the reference's GEMM tests, verbatim cases
def _run_ref_gemm_case(dtype, case):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    kind, m, x, op_a, op_b, packed, alpha, beta = case
    a_shape, b_shape, c_shape = rc.gemm_case_shapes(kind, m, x, op_a, op_b)
    a = wk.Tensor.alloc(ctx, pipe, a_shape, dtype)
    b = wk.Tensor.alloc(ctx, pipe, b_shape, dtype)
    c = wk.Tensor.alloc(ctx, pipe, c_shape, dtype)
    if kind == "AI":
        wk.tensor.memory.read_from_buffer(pipe, a, (np.arange(a_shape[0] * a_shape[1]) + 1).astype(dtype))
        wk.tensor.identity(pipe, b)
    else:
        wk.tensor.identity(pipe, a)
        wk.tensor.memory.read_from_buffer(pipe, b, (np.arange(b_shape[0] * b_shape[1]) + 1).astype(dtype))
    if beta is not None:
        wk.tensor.fill.one(pipe, c)
    pt = wk.blas.PackedTensors.init(pipe, c, x if kind == "AI" else m, True) if packed else None
    wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c, pt)
    got = gh.to_np(c)
    exp = rc.gemm_case_expected(kind, m, x, op_a, op_b, alpha, beta, np, np.dtype(dtype).type)
    np.testing.assert_array_equal(got, exp)  # the reference uses expectEqual: exact
    for t in (a, b, c):
        t.release(pipe)


@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("path", [0, 1])
def test_gemm_all_algorithms_non_complex(dtype, path):
    """gemm.zig:945 / :1133 -- path 0 = automatic back-end choice, 1 = SIMT forced"""
    gh.wk().capi.lib().wk_gemm_set_path(path)
    try:
        for case in rc.GEMM_UNPACKED:
            _run_ref_gemm_case(dtype, case)
    finally:
        gh.wk().capi.lib().wk_gemm_set_path(0)


@pytest.mark.parametrize("dtype", FLOATS)
def test_gemm_all_algorithms_with_packing(dtype):
    """gemm.zig:1054 / :1230"""
    for case in rc.GEMM_PACKED:
        _run_ref_gemm_case(dtype, case)


def test_gemm_invalid_shapes():
    """gemm.zig:900"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    for a_s, b_s, c_s, op_a, op_b in rc.GEMM_INVALID:
        a, b, c = (wk.Tensor.alloc(ctx, pipe, s, np.float32) for s in (a_s, b_s, c_s))
        with pytest.raises(wk.capi.InvalidValue):
            wk.blas.gemm(pipe, None, a, op_a, b, op_b, None, c)


# ------------------------------------------------------------------ GEMM vs oracle + ideal on random inputs
RAGGED = [(1, 1, 1), (2, 3, 5), (17, 33, 9), (64, 64, 64), (100, 130, 70), (129, 257, 65), (256, 128, 192)]


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_random_vs_oracle(oracle, dtype, op_a, op_b):
    """every dtype x every transpose pair x (alpha,beta) in {(null,null),(a,null),(a,b),(null,b)} on ragged shapes:
    integers bit-exact against the restated reference kernels AND the exact mod-2^bits ideal; floats within the
    K-scaled bound of the fp64 ideal, with the restatement's own error required to respect the same bound."""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(1234 + op_a * 2 + op_b)
    is_f = np.dtype(dtype).kind == "f"
    for (M, N, K) in RAGGED:
        for alpha, beta in [(None, None), (1.25 if is_f else 3, None), (0.75 if is_f else 5, 0.5 if is_f else 7), (None, 2)]:
            a_shape = (K, M) if op_a else (M, K)
            b_shape = (N, K) if op_b else (K, N)
            ad, bd, cd = (gh.rand_data(rng, dtype, s) for s in (a_shape, b_shape, (M, N)))
            a, oa = gh.make_pair(oracle, dtype, a_shape, ad)
            b, ob = gh.make_pair(oracle, dtype, b_shape, bd)
            c, oc = gh.make_pair(oracle, dtype, (M, N), cd)
            wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
            oracle.gemm(alpha, oa, op_a, ob, op_b, beta, oc)
            got, ref = gh.to_np(c), oc.to_host()
            if is_f:
                ideal = gh.gemm_ideal(ad, op_a, bd, op_b, alpha, beta, cd)
                bound = gh.gemm_float_bound(ad, op_a, bd, op_b, alpha, beta, cd, tol=1.0)
                assert np.all(np.abs(ref.astype(np.float64) - ideal) <= bound), "oracle outside its own bound"
                err = np.abs(got.astype(np.float64) - ideal)
                assert np.all(err <= bound), f"{dtype.__name__} {M}x{N}x{K} max err {err.max()} bound {bound.min()}"
            else:
                if M * N * K <= 200000:  # python-int ideal is slow; larger shapes lean on the (already checked) oracle
                    np.testing.assert_array_equal(ref, gh.gemm_ideal(ad, op_a, bd, op_b, alpha, beta, cd))
                np.testing.assert_array_equal(got, ref)
            for t in (a, b, c):
                t.release(pipe)


@pytest.mark.parametrize("dtype", FLOATS)
def test_gemm_padding_untouched_and_logical_k(oracle, dtype):
    """C's padding is never written and garbage in A/B padding never enters the product (odd shapes)"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(7)
    M, N, K = 5, 7, 3
    a, _ = gh.make_pair(oracle, dtype, (M, K), rng.uniform(-1, 1, (M, K)))
    b, _ = gh.make_pair(oracle, dtype, (K, N), rng.uniform(-1, 1, (K, N)))
    c, _ = gh.make_pair(oracle, dtype, (M, N))
    # poison all padding: sigmoid over the whole padded buffer turns zeros into 0.5
    ad, bd = gh.to_np(a), gh.to_np(b)
    wk.math.sinh(pipe, c)  # no-op on zeros, just exercises the padded domain
    wk.capi.check(wk.capi.lib().wk_unary(pipe.q, a.type_index, 6, a.ptr, a.number_of_elements))
    wk.capi.check(wk.capi.lib().wk_unary(pipe.q, b.type_index, 6, b.ptr, b.number_of_elements))
    sa = 1 / (1 + np.exp(-ad.astype(np.float64)))
    sb = 1 / (1 + np.exp(-bd.astype(np.float64)))
    wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    np.testing.assert_allclose(gh.to_np(c), sa @ sb, rtol=1e-5 if dtype == np.float32 else 1e-12)
    pc = gh.padded(c).reshape(c.rows_padded, c.row_pitch)
    assert np.all(pc[:, N:] == 0) and np.all(pc[M:, :] == 0)


# ------------------------------------------------------------------ axpy: the reference's tests + oracle parity
@pytest.mark.parametrize("dtype", ALL)
def test_axpy_reference_cases(dtype):
    """axpy.zig:178 (alpha=2), :285 (null), :387 (-1, signed), :504 (2-D), :619 (3-D), :817 (alpha=0)"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()

    def run(shape, xs, ys, alpha):
        x = wk.Tensor.alloc(ctx, pipe, shape, dtype)
        y = wk.Tensor.alloc(ctx, pipe, shape, dtype)
        wk.tensor.memory.read_from_buffer(pipe, x, np.asarray(xs).astype(dtype))
        wk.tensor.memory.read_from_buffer(pipe, y, np.asarray(ys).astype(dtype))
        wk.blas.axpy(pipe, x, alpha, y)
        return gh.to_np(y).reshape(-1)

    r5 = np.arange(1, 6)
    np.testing.assert_array_equal(run((5,), r5, r5 * 10, 2), (r5 * 12).astype(dtype))
    np.testing.assert_array_equal(run((5,), r5, r5 * 10, None), (r5 * 11).astype(dtype))
    if dtype in SIGNED:
        np.testing.assert_array_equal(run((5,), r5, r5 * 10, -1), (r5 * 9).astype(dtype))
    for shape in [(2, 3), (2, 2, 2)]:
        n = int(np.prod(shape))
        np.testing.assert_array_equal(run(shape, np.arange(n), np.arange(n) * 3, 3), (np.arange(n) * 6).astype(dtype))
    np.testing.assert_array_equal(run((4,), np.arange(1, 5), np.arange(1, 5) * 10, 0), (np.arange(1, 5) * 10).astype(dtype))


def test_axpy_shape_errors():
    """axpy.zig:763, :790"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    with pytest.raises(wk.capi.UnqualTensorsShape):
        wk.blas.axpy(pipe, wk.Tensor.alloc(ctx, pipe, (2, 3)), 1, wk.Tensor.alloc(ctx, pipe, (3, 2)))
    with pytest.raises(wk.capi.UnqualTensorsShape):
        wk.blas.axpy(pipe, wk.Tensor.alloc(ctx, pipe, (6,)), 1, wk.Tensor.alloc(ctx, pipe, (2, 3)))


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("shape", [(1,), (7,), (1000003,), (3, 5), (4, 6), (64, 1000), (3, 5, 7), (2, 4, 8)])
def test_axpy_hadamard_scal_vs_oracle_bit_exact(oracle, dtype, shape):
    """ragged / padded / dense shapes, three alpha modes: bit-exact against the restated axpy.cl / dot.cl"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(99)
    is_f = np.dtype(dtype).kind == "f"
    for alpha in [None, -1 if dtype in SIGNED else 1, 0.37 if is_f else 3]:
        xd, yd = gh.rand_data(rng, dtype, shape), gh.rand_data(rng, dtype, shape)
        x, ox = gh.make_pair(oracle, dtype, shape, xd)
        y, oy = gh.make_pair(oracle, dtype, shape, yd)
        wk.blas.axpy(pipe, x, alpha, y)
        oracle.axpy(ox, alpha, oy)
        np.testing.assert_array_equal(gh.padded(y), oy.buf)
        wk.math.dot(pipe, y, x)
        oracle.hadamard(oy, ox)
        np.testing.assert_array_equal(gh.padded(y), oy.buf)
        wk.blas.scal(pipe, 3, y)
        exp = oy.to_host()
        exp = (exp * np.dtype(dtype).type(3)).astype(dtype) if is_f else gh.wrap_to_dtype(exp.astype(object) * 3, dtype)
        np.testing.assert_array_equal(gh.to_np(y), exp)
        for t in (x, y):
            t.release(pipe)


# ------------------------------------------------------------------ math
@pytest.mark.parametrize("dtype", ALL)
def test_math_reference_cases(dtype):
    """basic.zig:258 (hadamard [1..4]o[2..5]), :322 (sum 1..5 = 15), :368 (mean 2,4,6,8 = 5)"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    x = wk.Tensor.alloc(ctx, pipe, (4,), dtype)
    y = wk.Tensor.alloc(ctx, pipe, (4,), dtype)
    wk.tensor.memory.read_from_buffer(pipe, x, np.array([1, 2, 3, 4], dtype=dtype))
    wk.tensor.memory.read_from_buffer(pipe, y, np.array([2, 3, 4, 5], dtype=dtype))
    wk.math.dot(pipe, x, y)
    np.testing.assert_array_equal(gh.to_np(x), np.array([2, 6, 12, 20], dtype=dtype))
    s = wk.Tensor.alloc(ctx, pipe, (5,), dtype)
    wk.tensor.memory.read_from_buffer(pipe, s, np.array([1, 2, 3, 4, 5], dtype=dtype))
    assert wk.math.sum(pipe, s) == 15
    if np.dtype(dtype).kind == "f":
        m = wk.Tensor.alloc(ctx, pipe, (4,), dtype)
        wk.tensor.memory.read_from_buffer(pipe, m, np.array([2, 4, 6, 8], dtype=dtype))
        assert abs(float(wk.math.mean(pipe, m)) - 5.0) < 1e-5


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("shape", [(5,), (1000,), (33, 77), (4, 31, 65), (1 << 20,)])
def test_sum_vs_oracle(oracle, dtype, shape):
    """ints bit-exact (any order is exact mod 2^bits); floats within n*eps*sum|x| of the fp64 sum.  The padded
    columns take part exactly as in sum.cl:33-35 (they hold zeros here; see test_sum_includes_padding)."""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(5)
    xd = gh.rand_data(rng, dtype, shape)
    x, ox = gh.make_pair(oracle, dtype, shape, xd)
    got, ref = wk.math.sum(pipe, x), oracle.tsum(ox)
    if np.dtype(dtype).kind == "f":
        ideal = xd.astype(np.float64).sum()
        bound = xd.size * np.finfo(dtype).eps * np.abs(xd.astype(np.float64)).sum()
        assert abs(float(ref) - ideal) <= bound and abs(float(got) - ideal) <= bound
    else:
        assert got == ref
    x.release(pipe)


def test_sum_includes_padding(oracle):
    """after sigmoid the padding holds 0.5 and math.sum adds the padded COLUMNS (not the padded rows)"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    x, ox = gh.make_pair(oracle, np.float64, (3, 5))
    wk.capi.check(wk.capi.lib().wk_unary(pipe.q, x.type_index, 6, x.ptr, x.number_of_elements))
    oracle.unary(ox, "sigmoid")
    assert wk.math.sum(pipe, x) == oracle.tsum(ox) == 0.5 * 3 * 6


@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("op", ["sin", "cos", "tan", "sinh", "cosh", "tanh"])
def test_trig(oracle, dtype, op):
    """trig.zig:129-446 points (abs 1e-5) + random points against the oracle (libm) at a few ulp"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    pts = [0.0, math.pi / 6, math.pi / 4] if op == "tan" else [0.0, math.pi / 6, math.pi / 2, math.pi]
    if op in ("sinh", "cosh", "tanh"):
        pts = [0.0, 0.5, 1.0, 2.0]
    x, _ = gh.make_pair(oracle, dtype, (len(pts),), pts)
    getattr(wk.math, op)(pipe, x)
    np.testing.assert_allclose(gh.to_np(x), [getattr(math, op)(p) for p in pts], atol=1e-5, rtol=0)
    rng = np.random.default_rng(3)
    xd = rng.uniform(-1.4, 1.4, size=(37, 53)).astype(dtype)
    x, ox = gh.make_pair(oracle, dtype, xd.shape, xd)
    getattr(wk.math, op)(pipe, x)
    oracle.unary(ox, op)
    np.testing.assert_allclose(gh.padded(x), ox.buf, rtol=8 * np.finfo(dtype).eps, atol=8 * np.finfo(dtype).eps)


# ------------------------------------------------------------------ tensor utilities
@pytest.mark.parametrize("dtype", ALL)
def test_fill_identity_transpose_uniform_vs_oracle(oracle, dtype):
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    t, o = gh.make_pair(oracle, dtype, (3, 5))
    wk.tensor.fill.constant(pipe, t, 7)
    o.fill(7)
    np.testing.assert_array_equal(gh.padded(t), o.buf)
    for shape in [(6, 6), (5, 5), (4, 4, 4)]:
        t, o = gh.make_pair(oracle, dtype, shape)
        wk.tensor.identity(pipe, t)
        o.identity()
        np.testing.assert_array_equal(gh.padded(t), o.buf)
    with pytest.raises(wk.capi.InvalidValue):
        wk.tensor.identity(pipe, gh.make_pair(oracle, dtype, (3, 4))[0])
    src, osrc = gh.make_pair(oracle, dtype, (37, 53), np.arange(37 * 53) % 120)
    dst, odst = gh.make_pair(oracle, dtype, (53, 37))
    wk.tensor.transpose(pipe, dst, src, 0, 1)
    oracle.transpose(odst, osrc, 0, 1)
    np.testing.assert_array_equal(gh.padded(dst), odst.buf)
    # N-D: every pair of dimensions of a ragged 3-D and a 4-D tensor (transpose.zig tests + transpose.cl:3-42)
    for shape in [(2, 3, 4), (3, 5, 7), (2, 3, 5, 6)]:
        nd = len(shape)
        n = int(np.prod(shape))
        for d0 in range(nd):
            for d1 in range(d0 + 1, nd):
                rshape = list(shape)
                rshape[d0], rshape[d1] = rshape[d1], rshape[d0]
                src, osrc = gh.make_pair(oracle, dtype, shape, (np.arange(n) % 120).reshape(shape))
                dst, odst = gh.make_pair(oracle, dtype, tuple(rshape))
                wk.tensor.transpose(pipe, dst, src, d1, d0)
                oracle.transpose(odst, osrc, d0, d1)
                np.testing.assert_array_equal(gh.padded(dst), odst.buf)
                np.testing.assert_array_equal(gh.to_np(dst), np.swapaxes((np.arange(n) % 120).reshape(shape), d0, d1).astype(dtype))
                for t in (src, dst):
                    t.release(pipe)
    for shape in [(64, 100), (5, 7), (3, 5, 7), (1001,)]:
        for seed in (42, 43):
            t, o = gh.make_pair(oracle, dtype, shape)
            wk.tensor.random.uniform(pipe, t, seed)
            o.uniform(seed)
            np.testing.assert_array_equal(gh.padded(t), o.buf)
            lo, hi = (-0.5, 0.75) if np.dtype(dtype).kind == "f" else (3, 100)
            wk.tensor.random.uniform(pipe, t, seed, lo, hi)
            o.uniform(seed, lo, hi)
            np.testing.assert_array_equal(gh.padded(t), o.buf)


def test_memory_roundtrip_put_get():
    """tensor/memory tests: host -> padded tensor -> host round trip, putValue/getValue, copy"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    for dtype in ALL:
        for shape in [(5,), (3, 5), (2, 3, 5)]:
            data = (np.arange(int(np.prod(shape))) % 100).astype(dtype).reshape(shape)
            t = wk.Tensor.alloc(ctx, pipe, shape, dtype)
            wk.tensor.memory.read_from_buffer(pipe, t, data)
            np.testing.assert_array_equal(gh.to_np(t), data)
            t2 = wk.Tensor.alloc(ctx, pipe, shape, dtype)
            wk.tensor.memory.copy(pipe, t, t2)
            np.testing.assert_array_equal(gh.to_np(t2), data)
            coords = tuple(s - 1 for s in shape)
            wk.tensor.memory.put_value(pipe, t, coords, 42)
            assert wk.tensor.memory.get_value(pipe, t, coords) == 42
            with pytest.raises(wk.capi.InvalidCoordinates):
                wk.tensor.memory.get_value(pipe, t, shape)
        with pytest.raises(wk.capi.InvalidBuffer):
            wk.tensor.memory.read_from_buffer(pipe, t, np.zeros(3, dtype=dtype))
    with pytest.raises(wk.capi.InvalidValue):
        wk.Tensor.alloc(ctx, pipe, (2, 0, 4))


# ------------------------------------------------------------------ nn kernels vs oracle (parity unpinned in the reference)
@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("shape", [(4, 1), (4, 10), (33, 77), (256, 512)])
def test_nn_kernels_vs_oracle(oracle, dtype, shape):
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    lib = wk.capi.lib()
    rng = np.random.default_rng(11)
    eps = np.finfo(dtype).eps
    out, oout = gh.make_pair(oracle, dtype, shape, rng.uniform(-3, 3, shape))
    # sigmoid (exp: few ulp), whole padded buffer => padding becomes 0.5 on both sides
    wk.nn.Sigmoid.init().run(pipe, out)
    oracle.unary(oout, "sigmoid")
    np.testing.assert_allclose(gh.padded(out), oout.buf, rtol=4 * eps, atol=0)
    # from here on feed both sides the SAME bits so the IEEE-only kernels must agree bit for bit
    oout.buf[:] = gh.padded(out)
    dev, odev = gh.make_pair(oracle, dtype, shape)
    wk.nn.Sigmoid.init().get_derivative(pipe, out, dev)
    oracle.sigmoid_dev(oout, odev)
    np.testing.assert_array_equal(gh.padded(dev), odev.buf)
    wk.nn.Tanh.init().get_derivative(pipe, out, dev)
    oracle.tanh_dev(oout, odev)
    np.testing.assert_array_equal(gh.padded(dev), odev.buf)
    # bias add + bias step
    bias, obias = gh.make_pair(oracle, dtype, (shape[1],), rng.uniform(-1, 1, shape[1]))
    wk.nn.Linear._add_bias(pipe, out, bias)
    oracle.bias(oout, obias)
    np.testing.assert_array_equal(gh.padded(out), oout.buf)
    bg, obg = gh.make_pair(oracle, dtype, (shape[1],))
    wk.nn.Linear._bias_sensitivity(pipe, out, bg)
    oracle.bias_step(oout, obg)
    if shape[0] <= 256:
        np.testing.assert_array_equal(gh.padded(bg), obg.buf)  # same summation order as bias_step.cl
    else:
        np.testing.assert_allclose(gh.padded(bg), obg.buf, rtol=shape[0] * eps)
    # mse with derivative
    exp_t, oexp = gh.make_pair(oracle, dtype, shape, rng.uniform(0, 1, shape))
    err, oerr = gh.make_pair(oracle, dtype, shape)
    wk.capi.check(lib.wk_mse(pipe.q, out.type_index, out.ptr, exp_t.ptr, err.ptr, dev.ptr, out.number_of_elements))
    oracle.mse(oout, oexp, oerr, odev)
    np.testing.assert_array_equal(gh.padded(err), oerr.buf)
    np.testing.assert_array_equal(gh.padded(dev), odev.buf)
    # optimizers: x, g, state
    for name in ("gdm", "adagrad", "rmsprop"):
        x, ox = gh.make_pair(oracle, dtype, shape, rng.uniform(-1, 1, shape))
        g, og = gh.make_pair(oracle, dtype, shape, rng.uniform(-1, 1, shape))
        h, oh = gh.make_pair(oracle, dtype, shape, rng.uniform(0, 1, shape))
        lr, p2 = np.array([0.01], dtype=dtype), np.array([0.9], dtype=dtype)
        for _ in range(3):
            if name == "gdm":
                wk.capi.check(lib.wk_gdm(pipe.q, x.type_index, x.ptr, g.ptr, h.ptr, lr.ctypes.data, p2.ctypes.data, x.number_of_elements))
                oracle.gdm(ox, og, oh, lr[0], p2[0])
            elif name == "adagrad":
                wk.capi.check(lib.wk_adagrad(pipe.q, x.type_index, x.ptr, g.ptr, h.ptr, lr.ctypes.data, x.number_of_elements))
                oracle.adagrad(ox, og, oh, lr[0])
            else:
                wk.capi.check(lib.wk_rmsprop(pipe.q, x.type_index, x.ptr, g.ptr, h.ptr, lr.ctypes.data, p2.ctypes.data, x.number_of_elements))
                oracle.rmsprop(ox, og, oh, lr[0], p2[0])
        np.testing.assert_array_equal(gh.padded(x), ox.buf)
        np.testing.assert_array_equal(gh.padded(h), oh.buf)


def test_nn_type_not_supported():
    """sigmoid.zig:21-24: integer dtypes -> TypeNotSupported"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    t = wk.Tensor.alloc(ctx, pipe, (4, 4), np.int32)
    with pytest.raises(wk.capi.TypeNotSupported):
        wk.nn.Sigmoid.init().run(pipe, t)


@pytest.mark.parametrize("dtype", FLOATS)
def test_adam_textbook(dtype):
    """Adam has no reference implementation (adam.zig is empty): checked against the textbook update in numpy"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(2)
    n = 1000
    x0, g0 = rng.uniform(-1, 1, n).astype(dtype), rng.uniform(-1, 1, n).astype(dtype)
    x, g, m, v = (wk.Tensor.alloc(ctx, pipe, (n,), dtype) for _ in range(4))
    wk.tensor.memory.read_from_buffer(pipe, x, x0)
    wk.tensor.memory.read_from_buffer(pipe, g, g0)
    lr, b1, b2, eps = (np.array([c], dtype=dtype) for c in (1e-2, 0.9, 0.999, 1e-8))
    xr, mr, vr = x0.astype(np.float64), np.zeros(n), np.zeros(n)
    for t in range(1, 4):
        wk.capi.check(wk.capi.lib().wk_adam(pipe.q, x.type_index, x.ptr, g.ptr, m.ptr, v.ptr, lr.ctypes.data, b1.ctypes.data,
                                            b2.ctypes.data, eps.ctypes.data, t, x.number_of_elements))
        mr = 0.9 * mr + 0.1 * g0
        vr = 0.999 * vr + 0.001 * g0.astype(np.float64) ** 2
        xr = xr - 1e-2 * (mr / (1 - 0.9 ** t)) / (np.sqrt(vr / (1 - 0.999 ** t)) + 1e-8)
    np.testing.assert_allclose(gh.to_np(x), xr, rtol=2e-5 if dtype == np.float32 else 1e-12)
