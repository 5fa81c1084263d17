"""Pins the CPU restatement (oracle/) against the reference's own known-answer tests.

Each test names the reference test it transcribes.  The reference's tests run on whatever OpenCL
device is present, so every GEMM case is run against all three fake devices the oracle models:
a PoCL-like CPU with vector width 16 and 8 (`.global` local memory -> gemm_2x2 / gemm_nxn kernels)
and an NVIDIA-OpenCL-like GPU (`.local` -> gemm_nxn_gpu), plus the vw=1 device our CUDA queue reports.
"""
import math

import numpy as np
import pytest

from tests import ref_cases as rc

FLOATS = [np.float32, np.float64]
ALL = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
SIGNED = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]
DEVICES = [("cpu", 16), ("cpu", 8), ("gpu", 1), ("b200", 1)]


def _dev(oracle, spec):
    return oracle.device(spec[0], spec[1])


def _run_gemm_case(oracle, dev, dtype, case):
    kind, m, x, op_a, op_b, packed, alpha, beta = case
    a_shape, b_shape, c_shape = rc.gemm_case_shapes(kind, m, x, op_a, op_b)
    a = oracle.OTensor(dev, dtype, a_shape)
    b = oracle.OTensor(dev, dtype, b_shape)
    c = oracle.OTensor(dev, dtype, c_shape)
    if kind == "AI":
        a.read_from(np.arange(a_shape[0] * a_shape[1]) + 1)
        b.identity()
    else:
        a.identity()
        b.read_from(np.arange(b_shape[0] * b_shape[1]) + 1)
    if beta is not None:
        c.fill(1)
    oracle.gemm(alpha, a, op_a, b, op_b, beta, c, packed=packed)
    exp = rc.gemm_case_expected(kind, m, x, op_a, op_b, alpha, beta, np, np.dtype(dtype).type)
    np.testing.assert_array_equal(c.to_host(), exp)  # expectEqual: exact


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", FLOATS)
def test_gemm_all_algorithms_non_complex(oracle, devspec, dtype):
    """gemm.zig:945 / :1133 'gemm cpu|gpu - all algorithms, non-complex'"""
    dev = _dev(oracle, devspec)
    for case in rc.GEMM_UNPACKED:
        _run_gemm_case(oracle, dev, dtype, case)


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", FLOATS)
def test_gemm_all_algorithms_with_packing(oracle, devspec, dtype):
    """gemm.zig:1054 / :1230 'gemm cpu|gpu - all algorithms with packing, non-complex'"""
    dev = _dev(oracle, devspec)
    for case in rc.GEMM_PACKED:
        _run_gemm_case(oracle, dev, dtype, case)


@pytest.mark.parametrize("dtype", [np.int8, np.uint16, np.int32, np.uint64])
def test_gemm_integer_identity(oracle, dtype):
    """not a reference test (integer GEMM is never executed by the reference suite): the same A*I / I*B
    closed form must hold mod 2^bits"""
    for devspec in DEVICES:
        dev = _dev(oracle, devspec)
        for case in rc.GEMM_UNPACKED[:16] + rc.GEMM_PACKED[:6]:
            kind, m, x, op_a, op_b, packed, alpha, beta = case
            a_shape, b_shape, c_shape = rc.gemm_case_shapes(kind, m, x, op_a, op_b)
            a = oracle.OTensor(dev, dtype, a_shape)
            b = oracle.OTensor(dev, dtype, b_shape)
            c = oracle.OTensor(dev, dtype, c_shape)
            data = (np.arange(a_shape[0] * a_shape[1] if kind == "AI" else b_shape[0] * b_shape[1]) + 1).astype(dtype)
            if kind == "AI":
                a.read_from(data)
                b.identity()
            else:
                a.identity()
                b.read_from(data)
            oracle.gemm(alpha, a, op_a, b, op_b, beta, c, packed=packed)
            np.testing.assert_array_equal(c.to_host().reshape(-1), data.reshape(c_shape).reshape(-1))


def test_gemm_invalid_shapes(oracle):
    """gemm.zig:900 'gemm - invalid shapes'"""
    dev = oracle.device("cpu", 16)
    for a_s, b_s, c_s, op_a, op_b in rc.GEMM_INVALID:
        a, b, c = (oracle.OTensor(dev, np.float32, s) for s in (a_s, b_s, c_s))
        with pytest.raises(ValueError):
            oracle.gemm(None, a, op_a, b, op_b, None, c)


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (1, 0), (0, 1)])
def test_pack_layout(oracle, devspec, dtype, op_a, op_b):
    """gemm.zig:1309-1623 'pack - normal / transposed packing (A), B packing with no_transpose / transpose'"""
    dev = _dev(oracle, devspec)
    n = 8
    a = oracle.OTensor(dev, dtype, (n, n))
    b = oracle.OTensor(dev, dtype, (n, n))
    c = oracle.OTensor(dev, dtype, (n, n))
    data = (np.arange(n * n) + 1).astype(dtype)
    a.read_from(data)
    b.read_from(data)
    g = oracle.packed_geom(dev, dtype, n, n, n, c.layout.gemm_algorithm, True)
    pa, pb = oracle.pack(g, a, op_a, b, op_b)
    bs = 2 << g.algorithm
    vw = g.a.vector_width if g.vectors_enabled else 1
    row_width = bs * vw
    src = data.reshape(n, n)

    def check(packed, lay, transpose):
        tile_rows, tile_cols, tile_data = lay.shape[0], lay.shape[1], lay.shape[2]
        # writeToBuffer of the 3-D packed tensor = logical view
        view = np.empty((tile_rows, tile_cols, tile_data), dtype=dtype)
        for tr in range(tile_rows):
            for tc in range(tile_cols):
                base = tr * lay.slice_pitch + tc * lay.row_pitch
                view[tr, tc] = packed[base:base + tile_data]
        for tr in range(tile_rows):
            for tc in range(tile_cols):
                for tile_row in range(bs):
                    for tile_col in range(row_width):
                        if transpose:
                            src_row, src_col = tc * row_width + tile_col, tr * bs + tile_row
                        else:
                            src_row, src_col = tr * bs + tile_row, tc * row_width + tile_col
                        if src_row < n and src_col < n:
                            assert view[tr, tc, tile_row * row_width + tile_col] == src[src_row, src_col]

    check(pa, g.a, op_a == 1)
    check(pb, g.b, op_b == 0)  # "inverted for B"


# ---------------------------------------------------------------------------------------------- axpy
@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", ALL)
def test_axpy_basic_1d(oracle, devspec, dtype):
    """axpy.zig:178 'axpy - basic operation y = alpha*x + y for 1D tensor'"""
    dev = _dev(oracle, devspec)
    x = oracle.OTensor(dev, dtype, (5,)).read_from(np.arange(1, 6))
    y = oracle.OTensor(dev, dtype, (5,)).read_from(np.arange(1, 6) * 10)
    oracle.axpy(x, 2, y)
    np.testing.assert_array_equal(y.to_host(), (np.arange(1, 6) * 12).astype(dtype))


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", ALL)
def test_axpy_alpha_null(oracle, devspec, dtype):
    """axpy.zig:285 'axpy - with alpha = null (direct sum) for all types'"""
    dev = _dev(oracle, devspec)
    x = oracle.OTensor(dev, dtype, (5,)).read_from(np.arange(1, 6))
    y = oracle.OTensor(dev, dtype, (5,)).read_from(np.arange(1, 6) * 10)
    oracle.axpy(x, None, y)
    np.testing.assert_array_equal(y.to_host(), (np.arange(1, 6) * 11).astype(dtype))


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", SIGNED)
def test_axpy_alpha_minus_one(oracle, devspec, dtype):
    """axpy.zig:387 'axpy - with alpha = -1 (subtraction) for signed types'"""
    dev = _dev(oracle, devspec)
    x = oracle.OTensor(dev, dtype, (5,)).read_from(np.arange(1, 6))
    y = oracle.OTensor(dev, dtype, (5,)).read_from(np.arange(1, 6) * 10)
    oracle.axpy(x, -1, y)
    np.testing.assert_array_equal(y.to_host(), (np.arange(1, 6) * 9).astype(dtype))


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("shape", [(2, 3), (2, 2, 2)])
def test_axpy_nd(oracle, devspec, dtype, shape):
    """axpy.zig:504 '2D tensor' (alpha=3, x=i, y=3i) and :619 '3D tensor'"""
    dev = _dev(oracle, devspec)
    n = int(np.prod(shape))
    x = oracle.OTensor(dev, dtype, shape).read_from(np.arange(n))
    y = oracle.OTensor(dev, dtype, shape).read_from(np.arange(n) * 3)
    oracle.axpy(x, 3, y)
    np.testing.assert_array_equal(y.to_host().reshape(-1), (np.arange(n) * 6).astype(dtype))


@pytest.mark.parametrize("dtype", ALL)
def test_axpy_vector_configurations_agree(oracle, dtype):
    """axpy.zig:680 'axpy - with different vector configurations'"""
    dev = oracle.device("cpu", 16)
    res = []
    for ve in (True, False):
        x = oracle.OTensor(dev, dtype, (2, 3), vectors_enabled=ve).read_from(np.arange(1, 7))
        y = oracle.OTensor(dev, dtype, (2, 3), vectors_enabled=ve).read_from(np.arange(1, 7) * 10)
        oracle.axpy(x, 2, y)
        res.append(y.to_host())
    np.testing.assert_array_equal(res[0], res[1])


def test_axpy_shape_errors(oracle):
    """axpy.zig:763 'incompatible tensor shapes', :790 'different number of dimensions'"""
    dev = oracle.device("cpu", 16)
    with pytest.raises(ValueError):
        oracle.axpy(oracle.OTensor(dev, np.float32, (2, 3)), 1, oracle.OTensor(dev, np.float32, (3, 2)))
    with pytest.raises(ValueError):
        oracle.axpy(oracle.OTensor(dev, np.float32, (6,)), 1, oracle.OTensor(dev, np.float32, (2, 3)))


@pytest.mark.parametrize("dtype", ALL)
def test_axpy_zero_alpha(oracle, dtype):
    """axpy.zig:817 'axpy - zero alpha'"""
    dev = oracle.device("cpu", 16)
    x = oracle.OTensor(dev, dtype, (4,)).read_from(np.arange(1, 5))
    y = oracle.OTensor(dev, dtype, (4,)).read_from(np.arange(1, 5) * 10)
    oracle.axpy(x, 0, y)
    np.testing.assert_array_equal(y.to_host(), (np.arange(1, 5) * 10).astype(dtype))


# ---------------------------------------------------------------------------------------------- math
@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", ALL)
def test_hadamard(oracle, devspec, dtype):
    """basic.zig:258 'dot - element-wise multiplication': [1..4] o [2..5] = [2,6,12,20]"""
    dev = _dev(oracle, devspec)
    x = oracle.OTensor(dev, dtype, (4,)).read_from([1, 2, 3, 4])
    y = oracle.OTensor(dev, dtype, (4,)).read_from([2, 3, 4, 5])
    oracle.hadamard(x, y)
    np.testing.assert_array_equal(x.to_host(), np.array([2, 6, 12, 20], dtype=dtype))


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", ALL)
def test_sum(oracle, devspec, dtype):
    """basic.zig:322 'sum - basic sum operation': sum(1..5) = 15"""
    dev = _dev(oracle, devspec)
    x = oracle.OTensor(dev, dtype, (5,)).read_from([1, 2, 3, 4, 5])
    assert oracle.tsum(x) == 15


@pytest.mark.parametrize("devspec", DEVICES)
@pytest.mark.parametrize("dtype", FLOATS)
def test_mean(oracle, devspec, dtype):
    """basic.zig:368 'mean - basic mean operation for float types': mean(2,4,6,8) = 5 (abs 1e-5)"""
    dev = _dev(oracle, devspec)
    x = oracle.OTensor(dev, dtype, (4,)).read_from([2, 4, 6, 8])
    assert abs(float(oracle.mean(x)) - 5.0) < 1e-5


@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("op", ["sin", "cos", "tan", "sinh", "cosh", "tanh"])
def test_trig(oracle, dtype, op):
    """trig.zig:129-446: values at 0, pi/6, pi/2, pi (tan: 0, pi/6, pi/4; hyperbolic: 0, 0.5, 1, 2), abs 1e-5"""
    dev = oracle.device("cpu", 16)
    pts = {"tan": [0.0, math.pi / 6, math.pi / 4]}.get(op, [0.0, math.pi / 6, math.pi / 2, math.pi])
    if op in ("sinh", "cosh", "tanh"):
        pts = [0.0, 0.5, 1.0, 2.0]
    x = oracle.OTensor(dev, dtype, (len(pts),)).read_from(pts)
    oracle.unary(x, op)
    exp = [getattr(math, op)(p) for p in pts]
    np.testing.assert_allclose(x.to_host(), exp, atol=1e-5, rtol=0)


# ---------------------------------------------------------------------------------------------- tensor
@pytest.mark.parametrize("dtype", ALL)
def test_layout_invariants(oracle, dtype):
    """tensor/main.zig:288-420 (shape bookkeeping) + the worked examples of SURVEY section 8"""
    for devspec in DEVICES:
        dev = _dev(oracle, devspec)
        t = oracle.OTensor(dev, dtype, (10,))
        assert t.layout.number_of_elements_without_padding == 10 and t.layout.number_of_elements >= 10
        assert t.layout.rows_padded == 2  # 1-D tensors are allocated 2x (main.zig:165-168)
        t = oracle.OTensor(dev, dtype, (2, 3, 4))
        assert tuple(t.layout.shape[:3]) == (2, 3, 4) and t.layout.depth == 2 and t.layout.rows_padded == 4
        assert t.layout.pitches[0] == t.layout.slice_pitch and t.layout.pitches[1] == t.layout.row_pitch
        assert t.layout.row_pitch % (2 * t.layout.vector_width) == 0
    with pytest.raises(ValueError):
        oracle.OTensor(oracle.device("cpu", 16), dtype, (2, 0, 4))
    b200 = oracle.device("b200")
    assert oracle.OTensor(b200, dtype, (1024, 1024)).layout.row_pitch == 1024
    assert oracle.OTensor(b200, dtype, (4, 1)).layout.row_pitch == 2
    assert oracle.OTensor(b200, dtype, (1 << 20,)).layout.number_of_elements == 1 << 21


def test_tile_choice(oracle):
    """SURVEY A.2 table computed from work_configuration.zig:110-193"""
    cpu16, cpu8, gpu = oracle.device("cpu", 16), oracle.device("cpu", 8), oracle.device("gpu")
    assert oracle.OTensor(cpu16, np.float32, (1024, 1024)).layout.gemm_algorithm == 3  # 16x16
    assert oracle.OTensor(cpu8, np.float32, (1024, 1024)).layout.gemm_algorithm == 3
    assert oracle.OTensor(cpu16, np.int8, (1024, 1024)).layout.gemm_algorithm == 4  # 32x32
    assert oracle.OTensor(gpu, np.float32, (1024, 1024)).layout.gemm_algorithm == 4  # 32x32, 256 WIs
    assert oracle.OTensor(gpu, np.float64, (1024, 1024)).layout.gemm_algorithm == 4
    assert oracle.lib().wko_get_algorithm(5, 48) == 3 and oracle.lib().wko_get_algorithm(2, 64) == 2


def test_calculate_work_items(oracle):
    """utils.zig:6-32"""
    assert oracle.calculate_work_items([1024], 256) == [256]
    assert oracle.calculate_work_items([100], 256) == [100]
    # pow(1000, 1/3) = 9.999.. -> 9 -> largest divisor of 1000 below it is 8 (float quirk kept on purpose)
    assert oracle.calculate_work_items([7, 1000, 3], 1000) == [7, 8, 3]
    assert oracle.calculate_work_items([64, 48], 1024) == [32, 24]


@pytest.mark.parametrize("dtype", ALL)
def test_fill_identity(oracle, dtype):
    """fill.zig / identity.zig tests: constant fill touches only the logical region; identity = zero + diagonal"""
    dev = oracle.device("cpu", 16)
    t = oracle.OTensor(dev, dtype, (3, 5)).fill(7)
    np.testing.assert_array_equal(t.to_host(), np.full((3, 5), 7, dtype=dtype))
    assert t.buf.sum() == 7 * 15  # padding untouched (still zero)
    i3 = oracle.OTensor(dev, dtype, (4, 4, 4)).identity().to_host()
    exp = np.zeros((4, 4, 4), dtype=dtype)
    for d in range(4):
        exp[d, d, d] = 1
    np.testing.assert_array_equal(i3, exp)
    with pytest.raises(ValueError):
        oracle.OTensor(dev, dtype, (3, 4)).identity()


@pytest.mark.parametrize("dtype", ALL)
def test_transpose(oracle, dtype):
    """transpose.zig tests: 2-D and N-D swap of two dims"""
    dev = oracle.device("cpu", 8)
    src = oracle.OTensor(dev, dtype, (3, 5)).read_from(np.arange(15))
    dst = oracle.OTensor(dev, dtype, (5, 3))
    oracle.transpose(dst, src, 0, 1)
    np.testing.assert_array_equal(dst.to_host(), np.arange(15).reshape(3, 5).T.astype(dtype))
    src = oracle.OTensor(dev, dtype, (2, 3, 4)).read_from(np.arange(24))
    dst = oracle.OTensor(dev, dtype, (4, 3, 2))
    oracle.transpose(dst, src, 0, 2)
    np.testing.assert_array_equal(dst.to_host(), np.arange(24).reshape(2, 3, 4).transpose(2, 1, 0).astype(dtype))


@pytest.mark.parametrize("dtype", ALL)
def test_uniform_properties(oracle, dtype):
    """random/uniform.zig:215-833: range checks, same-seed determinism, different-seed difference, statistics"""
    dev = oracle.device("cpu", 16)
    a = oracle.OTensor(dev, dtype, (64, 100)).uniform(42).to_host()
    b = oracle.OTensor(dev, dtype, (64, 100)).uniform(42).to_host()
    c = oracle.OTensor(dev, dtype, (64, 100)).uniform(43).to_host()
    np.testing.assert_array_equal(a, b)
    assert (a != c).mean() > 0.5
    if np.dtype(dtype).kind == "f":
        assert a.min() >= 0 and a.max() <= 1
        assert abs(a.mean() - 0.5) < 0.02 and abs(a.std() - (1 / 12) ** 0.5) < 0.02
        r = oracle.OTensor(dev, dtype, (64, 100)).uniform(42, -5, 5).to_host()
        assert r.min() >= -5 and r.max() <= 5 and abs(r.mean()) < 0.2
    else:
        lo, hi = (10, 100)
        r = oracle.OTensor(dev, dtype, (64, 100)).uniform(42, lo, hi).to_host()
        assert r.min() >= lo and r.max() <= hi


def test_xxhash_mixer_self_consistency(oracle):
    """uniform.cl:32-54 restated independently in Python ints (operator precedence + the dead seed2 term)"""
    M = (1 << 64) - 1

    def rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & M

    def ref(index, seed):
        key = ((0x7C01812CF721AD1C ^ 0xDED46DE9839097DB) - seed) & M
        comb = (((index & 0xFFFFFFFF) << 32) + (index >> 32)) & M
        x0 = comb ^ key
        x1 = x0 ^ rotl(x0, 49) ^ ((rotl(x0, 24) * 0x9FB21C651E98DF25) & M)
        x2 = x1 ^ ((((x1 >> 35) + 8) * 0x9FB21C651E98DF25) & M)
        return x2 ^ (x2 >> 28)

    for idx in [0, 1, 2, 12345, (1 << 32) + 7, (1 << 40) - 1]:
        for seed in [0, 42, 43, 44, (1 << 63) + 5]:
            assert oracle.xxhash64(idx, seed) == ref(idx, seed)
