"""-m gpu: the bandwidth-tuned variants of bias add (bias.cl), bias step (bias_step.cl) and the 2-D transpose
(transpose.cl) at sizes where the 128-bit / multi-chunk code paths are taken, against the oracle (bias kernels) and
numpy (transpose is a permutation: bit-exact for every element size)."""
import numpy as np
import pytest

from . import gpu_helpers as gh

pytestmark = pytest.mark.gpu

FLOATS = [np.float32, np.float64]


@pytest.fixture(scope="module")
def oracle():
    from oracle import pyoracle

    return pyoracle


@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("shape", [(2, 4), (3, 5), (64, 36), (1000, 36), (513, 1028), (4096, 1024), (7, 30002)])
def test_bias_add_exact(oracle, dtype, shape):
    """out[i] += bias[i % row_pitch] over the whole padded buffer (bias.cl:3-19): one IEEE add per element => bit-exact"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(5)
    out, oout = gh.make_pair(oracle, dtype, shape, rng.uniform(-3, 3, shape))
    bias, obias = gh.make_pair(oracle, dtype, (shape[1],), rng.uniform(-1, 1, shape[1]))
    wk.nn.Linear._add_bias(pipe, out, bias)
    oracle.bias(oout, obias)
    np.testing.assert_array_equal(gh.padded(out), oout.buf)
    for t in (out, bias):
        t.release(pipe)


@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("shape", [(2, 4), (256, 36), (256, 1028), (1000, 36), (5000, 1028), (32768, 256), (300, 30002)])
def test_bias_step(oracle, dtype, shape):
    """column sums (bias_step.cl:24-37): rows <= 256 are summed in the reference's order (bit-exact); taller inputs are
    split into row chunks (deterministic, different association) and agree within rows*eps*sum|x|"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(6)
    data = rng.uniform(-1, 1, shape).astype(dtype)
    sens, osens = gh.make_pair(oracle, dtype, shape, data)
    bg, obg = gh.make_pair(oracle, dtype, (shape[1],))
    wk.nn.Linear._bias_sensitivity(pipe, sens, bg)
    oracle.bias_step(osens, obg)
    got, want = gh.to_np(bg), obg.to_host()
    if shape[0] <= 256:
        np.testing.assert_array_equal(got, want)
    else:
        bound = shape[0] * np.finfo(dtype).eps * np.abs(data.astype(np.float64)).sum(axis=0)
        assert np.all(np.abs(got.astype(np.float64) - data.astype(np.float64).sum(axis=0)) <= bound)
        assert np.all(np.abs(want.astype(np.float64) - data.astype(np.float64).sum(axis=0)) <= bound)
    # twice the same launch => the same bits (no atomics)
    bg2, _ = gh.make_pair(oracle, dtype, (shape[1],))
    wk.nn.Linear._bias_sensitivity(pipe, sens, bg2)
    np.testing.assert_array_equal(gh.to_np(bg2), got)
    for t in (sens, bg, bg2):
        t.release(pipe)


@pytest.mark.parametrize("dtype", [np.int8, np.uint16, np.float32, np.int64, np.float64])
@pytest.mark.parametrize("shape", [(64, 64), (128, 2048), (1000, 36), (516, 1028), (2050, 130), (33, 77), (4096, 4096)])
def test_transpose2d_exact(dtype, shape):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(7)
    data = gh.rand_data(rng, dtype, shape)
    a = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    b = wk.Tensor.alloc(ctx, pipe, shape[::-1], dtype)
    wk.tensor.memory.read_from_buffer(pipe, a, data)
    wk.tensor.transpose(pipe, b, a, 0, 1)
    np.testing.assert_array_equal(gh.to_np(b), data.T)
    # padding of the destination stays zero
    pad = gh.padded(b).reshape(b.rows_padded, b.row_pitch)
    assert not pad[b.rows:, :].any() and not pad[:, b.cols:].any()
    for t in (a, b):
        t.release(pipe)


@pytest.mark.parametrize("op", ["tanh", "sigmoid", "cosh"])
def test_f64_activation_accuracy_and_special_values(op):
    """the branch-free f64 tanh / sigmoid (csrc/common.cuh) against numpy's libm over the whole range: a few ulp
    relative (the reference tests to 1e-5 absolute, trig.zig:129-446), exact limits, NaN kept, tiny arguments not
    lost to cancellation"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(17)
    x = np.concatenate([
        rng.uniform(-1.5, 1.5, 20000), rng.uniform(-25, 25, 20000), rng.uniform(-760, 760, 20000),
        rng.uniform(0, 1, 20000) * 2.0 ** -rng.integers(0, 60, 20000) * rng.choice([-1.0, 1.0], 20000),
        [0.0, -0.0, 1e-300, -1e-300, 19.9, 20.0, 20.1, 40.0, -40.0, 700.0, -700.0, 709.0, 710.47, -710.47, 710.48, 711.0, -745.0, 1e308, -1e308,
         np.inf, -np.inf, np.nan],
    ])
    x = np.resize(x, (len(x) // 2 * 2,)).reshape(2, -1)
    t = wk.Tensor.alloc(ctx, pipe, x.shape, np.float64)
    wk.tensor.memory.read_from_buffer(pipe, t, x)
    if op == "tanh":
        wk.math.tanh(pipe, t)
        want = np.tanh(x.astype(np.longdouble)).astype(np.float64)
    elif op == "cosh":
        wk.math.cosh(pipe, t)
        with np.errstate(over="ignore"):
            want = np.cosh(x.astype(np.longdouble)).astype(np.float64)
    else:
        wk.nn.Sigmoid.init().run(pipe, t)
        with np.errstate(over="ignore"):
            want = (1.0 / (1.0 + np.exp(-x.astype(np.longdouble)))).astype(np.float64)
    got = gh.to_np(t)
    t.release(pipe)
    finite = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(x))
    assert np.array_equal(np.isinf(got), np.isinf(want)) and np.all(got[np.isinf(want)] == want[np.isinf(want)])
    big = finite & (np.abs(want) > 1e-290)
    rel = np.abs(got[big] - want[big]) / np.abs(want[big])
    assert rel.max() <= 4 * np.finfo(np.float64).eps, rel.max()
    assert np.all(np.abs(got[finite & ~big]) <= 1e-290)
    if op == "tanh":
        assert np.array_equal(np.signbit(got[~np.isnan(x)]), np.signbit(x[~np.isnan(x)]))
        assert got[x == np.inf][0] == 1.0 and got[x == -np.inf][0] == -1.0


def test_f64_tan_accuracy():
    """branch-free f64 tan (csrc/common.cuh: Cody-Waite by pi/2, sin/cos kernels, one quotient) against long-double libm:
    small, moderate and large arguments, the libdevice hand-over at 1e5, signed zero"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(19)
    x = np.concatenate([rng.uniform(-1.5, 1.5, 20000), rng.uniform(-100, 100, 20000), rng.uniform(-1e5, 1e5, 20000),
                        rng.uniform(-1e7, 1e7, 20000), rng.uniform(0, 1, 5000) * 2.0 ** -rng.integers(0, 200, 5000),
                        [0.0, -0.0, 99999.9, 100000.0, 100000.1, np.pi / 4, -np.pi / 4, 1e-300, 1e300]])
    x = np.resize(x, (len(x) // 2 * 2,)).reshape(2, -1)
    t = wk.Tensor.alloc(ctx, pipe, x.shape, np.float64)
    wk.tensor.memory.read_from_buffer(pipe, t, x)
    wk.math.tan(pipe, t)
    got = gh.to_np(t)
    t.release(pipe)
    want = np.tan(x.astype(np.longdouble)).astype(np.float64)
    nz = want != 0
    rel = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
    assert rel.max() <= 6 * np.finfo(np.float64).eps, rel.max()
    assert np.array_equal(got[~nz], x[~nz]) and np.array_equal(np.signbit(got[~nz]), np.signbit(x[~nz]))


@pytest.mark.parametrize("op", ["sin", "cos"])
def test_f64_sin_cos_accuracy(op):
    """f64 sin / cos on the same reduction as tan (quadrant picks the fdlibm kernel and the sign): relative accuracy for
    moderate arguments, absolute accuracy up to the libdevice hand-over at 1e5 and beyond it, signed zero, non-finite"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(23)
    parts = [rng.uniform(-1.5, 1.5, 20000), rng.uniform(-100, 100, 20000), rng.uniform(-1e5, 1e5, 20000),
             rng.uniform(-1e7, 1e7, 20000), rng.uniform(0, 1, 5000) * 2.0 ** -rng.integers(0, 200, 5000),
             np.array([0.0, -0.0, 99999.9, 100000.0, 100000.1, np.pi / 2, -np.pi / 2, np.pi, 1e-300, 1e300, np.inf, -np.inf, np.nan])]
    x = np.concatenate(parts)
    x = np.resize(x, (len(x) // 2 * 2,)).reshape(2, -1)
    t = wk.Tensor.alloc(ctx, pipe, x.shape, np.float64)
    wk.tensor.memory.read_from_buffer(pipe, t, x)
    getattr(wk.math, op)(pipe, t)
    got = gh.to_np(t)
    t.release(pipe)
    fin = np.isfinite(x)
    assert np.all(np.isnan(got[~fin]))
    with np.errstate(invalid="ignore"):
        want = getattr(np, op)(x.astype(np.longdouble)).astype(np.float64)
    eps = np.finfo(np.float64).eps
    assert np.abs(got[fin] - want[fin]).max() <= 1.5 * eps
    mod = fin & (np.abs(x) <= 100) & (want != 0)
    assert (np.abs(got[mod] - want[mod]) / np.abs(want[mod])).max() <= 4 * eps
    if op == "sin":
        z = x == 0
        assert np.array_equal(np.signbit(got[z]), np.signbit(x[z])) and np.all(got[z] == 0)


@pytest.mark.parametrize("op", ["sin", "cos", "tan"])
def test_f32_trig_accuracy(op):
    """f32 sin / cos / tan (branch-free fast path, libdevice slow path out of line at |x| >= 105615): against float64 libm"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(29)
    parts = [rng.uniform(-1.5, 1.5, 20000), rng.uniform(-100, 100, 20000), rng.uniform(-1e5, 1e5, 20000),
             rng.uniform(-1e7, 1e7, 20000), rng.uniform(0, 1, 5000) * 2.0 ** -rng.integers(0, 100, 5000),
             np.array([0.0, -0.0, 105614.0, 105615.0, 105616.0, np.pi / 4, -np.pi / 4, np.pi, 1e-30, 1e30, np.inf, -np.inf, np.nan])]
    x = np.concatenate(parts).astype(np.float32)
    x = np.resize(x, (len(x) // 2 * 2,)).reshape(2, -1)
    t = wk.Tensor.alloc(ctx, pipe, x.shape, np.float32)
    wk.tensor.memory.read_from_buffer(pipe, t, x)
    getattr(wk.math, op)(pipe, t)
    got = gh.to_np(t).astype(np.float64)
    t.release(pipe)
    fin = np.isfinite(x)
    assert np.all(np.isnan(got[~fin]))
    with np.errstate(invalid="ignore"):
        want = getattr(np, op)(x.astype(np.float64))
    eps = float(np.finfo(np.float32).eps)
    nz = fin & (want != 0)
    rel = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
    if op == "tan":
        assert rel.max() <= 6 * eps, rel.max()
    else:
        assert np.abs(got[fin] - want[fin]).max() <= 1.5 * eps
        mod = nz & (np.abs(x) <= 100)
        assert (np.abs(got[mod] - want[mod]) / np.abs(want[mod])).max() <= 4 * eps
    if op != "cos":
        z = x == 0
        assert np.array_equal(np.signbit(got[z]), np.signbit(x[z])) and np.all(got[z] == 0)


@pytest.mark.parametrize("dtype,shape", [(np.float32, (3000, 1000)), (np.int16, (4098, 1026)), (np.float64, (2, 1500, 700))])
def test_memory_copy_large_dense_is_bit_exact(dtype, shape):
    """memory.copy of >= 8 MB dense tensors runs as a streaming kernel (csrc/runtime.cu: wk_d2d -> copy_dense) instead of a
    driver copy: every byte, padding included, must arrive; a span whose byte count is not a multiple of 16 exercises the tail"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    a = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    b = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    wk.tensor.random.uniform(pipe, a, 5)
    before = wk.capi.launch_count()
    wk.tensor.memory.copy(pipe, a, b)
    assert wk.capi.launch_count() == before + 1  # a kernel, not a DMA
    np.testing.assert_array_equal(gh.padded(b), gh.padded(a))
    assert a.size >= 8 << 20
    for t in (a, b):
        t.release(pipe)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_large_sum_and_dot_use_chunk_scheduled_reduction(dtype):
    """>= 37 MB dense inputs take the dynamically scheduled one-launch reduction (csrc/reduce.cu: reduce_dense_dyn_kernel): the
    value must not depend on which CTA summed which chunk -- repeated calls are bit-identical -- integers are exact (mod 2^32),
    floats agree with a float64 sum within n * eps * sum|x|"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    shape = (4096, 12290)  # 201 MB f32 / 403 MB f64; cols % 4 == 2
    x = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    y = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    wk.tensor.random.uniform(pipe, x, 3, -1, 1) if np.dtype(dtype).kind == "f" else wk.tensor.random.uniform(pipe, x, 3)
    wk.tensor.random.uniform(pipe, y, 4, -1, 1) if np.dtype(dtype).kind == "f" else wk.tensor.random.uniform(pipe, y, 4)
    xh, yh = gh.to_np(x), gh.to_np(y)
    sums = [wk.math.sum(pipe, x) for _ in range(4)]
    dots = [wk.blas.dot_reduce(pipe, x, y) for _ in range(4)]
    assert all(np.array_equal(np.asarray(s), np.asarray(sums[0])) for s in sums)
    assert all(np.array_equal(np.asarray(d), np.asarray(dots[0])) for d in dots)
    if np.dtype(dtype).kind == "i":
        assert np.int32(sums[0]) == xh.astype(np.int64).sum().astype(np.int32)
        with np.errstate(over="ignore"):
            assert np.int32(dots[0]) == (xh.astype(np.int64) * yh.astype(np.int64)).sum().astype(np.int32)
    else:
        eps = np.finfo(dtype).eps
        n = xh.size
        want_s, want_d = xh.astype(np.float64).sum(), (xh.astype(np.float64) * yh.astype(np.float64)).sum()
        # tree summation (4 accumulators x 256 threads per chunk, chunk partials, final tree): worst case log2(n) * eps * sum|.|
        assert abs(float(sums[0]) - want_s) <= 32 * eps * np.abs(xh).astype(np.float64).sum()
        assert abs(float(dots[0]) - want_d) <= 32 * eps * np.abs(xh.astype(np.float64) * yh.astype(np.float64)).sum()
    for t in (x, y):
        t.release(pipe)
