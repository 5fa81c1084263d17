"""-m gpu: SURVEY 8(f)4 -- the optimizer update of a whole parameter list in one launch (wk_optimizer_step_multi) must be
bit-identical to the reference-shaped loop of one kernel per tensor (gd.zig:55-94, rmsprop.zig:168-202), and the
device-resident reductions (wk_sum_async / wk_dot_reduce_async) bit-identical to their blocking forms."""
import ctypes as C

import numpy as np
import pytest

from . import gpu_helpers as gh

pytestmark = pytest.mark.gpu

FLOATS = [np.float32, np.float64]
SHAPES = [(1,), (3,), (10,), (10, 2), (1, 10), (64, 65), (4097,), (300, 333), (1 << 20,), (1025, 1023)]


def _tensors(wk, ctx, pipe, dtype, shapes, rng, lo, hi):
    out = []
    for s in shapes:
        t = wk.Tensor.alloc(ctx, pipe, s, dtype)
        wk.tensor.memory.read_from_buffer(pipe, t, rng.uniform(lo, hi, s).astype(dtype))
        out.append(t)
    return out


def _clone(wk, ctx, pipe, ts):
    out = []
    for t in ts:
        c = wk.Tensor.alloc(ctx, pipe, t.shape, t.dtype)
        wk.tensor.memory.copy(pipe, t, c)
        out.append(c)
    return out


@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("kind", ["gd", "gd_sub", "gdm", "adagrad", "rmsprop", "adam"])
def test_multi_tensor_step_equals_per_tensor(dtype, kind):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    lib = wk.capi.lib()
    from wekua_b200.nn import optimizer as opt

    rng = np.random.default_rng(21)
    xs = _tensors(wk, ctx, pipe, dtype, SHAPES, rng, -1, 1)
    gs = _tensors(wk, ctx, pipe, dtype, SHAPES, rng, -1, 1)
    s0 = _tensors(wk, ctx, pipe, dtype, SHAPES, rng, 0, 1)
    s1 = _tensors(wk, ctx, pipe, dtype, SHAPES, rng, 0, 1)
    xs2, s02, s12 = (_clone(wk, ctx, pipe, ts) for ts in (xs, s0, s1))
    sc = lambda v: np.array([v], dtype=dtype)  # noqa: E731
    lr, h0, h1, eps = sc(-1.0 if kind == "gd_sub" else 0.01), sc(0.9), sc(0.999), sc(1e-8)
    for step in range(1, 4):
        # fused: one launch (the list fits one table)
        before = wk.capi.launch_count()
        if kind.startswith("gd") and kind != "gdm":
            opt._step_multi(pipe, "gd", [(x, g, None, None) for x, g in zip(xs, gs)], lr[0])
        elif kind == "adam":
            opt._step_multi(pipe, "adam", list(zip(xs, gs, s0, s1)), lr[0], h0[0], h1[0], eps[0], step)
        elif kind == "adagrad":
            opt._step_multi(pipe, "adagrad", [(x, g, s, None) for x, g, s in zip(xs, gs, s0)], lr[0])
        else:
            opt._step_multi(pipe, kind, [(x, g, s, None) for x, g, s in zip(xs, gs, s0)], lr[0], h0[0])
        assert wk.capi.launch_count() - before == 1
        # the reference's shape: one kernel per tensor, over the whole padded buffer
        for x, g, a, b in zip(xs2, gs, s02, s12):
            n, p = x.number_of_elements, lambda v: v.ctypes.data  # noqa: E731
            if kind.startswith("gd") and kind != "gdm":
                wk.capi.check(lib.wk_axpy(pipe.q, x.type_index, 1, 1, n, p(lr), g.ptr, n, n, x.ptr, n, n))
            elif kind == "gdm":
                wk.capi.check(lib.wk_gdm(pipe.q, x.type_index, x.ptr, g.ptr, a.ptr, p(lr), p(h0), n))
            elif kind == "adagrad":
                wk.capi.check(lib.wk_adagrad(pipe.q, x.type_index, x.ptr, g.ptr, a.ptr, p(lr), n))
            elif kind == "rmsprop":
                wk.capi.check(lib.wk_rmsprop(pipe.q, x.type_index, x.ptr, g.ptr, a.ptr, p(lr), p(h0), n))
            else:
                wk.capi.check(lib.wk_adam(pipe.q, x.type_index, x.ptr, g.ptr, a.ptr, b.ptr, p(lr), p(h0), p(h1), p(eps), step, n))
    for a, b in zip(xs + s0 + s1, xs2 + s02 + s12):
        np.testing.assert_array_equal(gh.padded(a), gh.padded(b))
    for t in xs + gs + s0 + s1 + xs2 + s02 + s12:
        t.release(pipe)


def test_multi_tensor_long_list_and_errors():
    """more tensors than one table holds (24) -> several launches, same results; null buffers are reported"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    from wekua_b200.nn import optimizer as opt

    rng = np.random.default_rng(22)
    shapes = [(int(rng.integers(1, 700)),) for _ in range(61)]
    xs = _tensors(wk, ctx, pipe, np.float32, shapes, rng, -1, 1)
    gs = _tensors(wk, ctx, pipe, np.float32, shapes, rng, -1, 1)
    want = [gh.padded(x) - np.float32(0.5) * gh.padded(g) for x, g in zip(xs, gs)]
    before = wk.capi.launch_count()
    opt._step_multi(pipe, "gd", [(x, g, None, None) for x, g in zip(xs, gs)], -0.5)
    assert wk.capi.launch_count() - before == 3
    for x, w in zip(xs, want):
        np.testing.assert_array_equal(gh.padded(x), w)
    arr = (wk.capi.OptParam * 1)(wk.capi.OptParam(xs[0].ptr, None, None, None, 4))
    lr = np.array([0.1], dtype=np.float32)
    assert wk.capi.lib().wk_optimizer_step_multi(pipe.q, 8, 0, arr, 1, lr.ctypes.data, None, None, None, 0) == 3  # InvalidBuffer
    assert wk.capi.lib().wk_optimizer_step_multi(pipe.q, 4, 0, arr, 1, lr.ctypes.data, None, None, None, 0) == 9  # int32: TypeNotSupported
    for t in xs + gs:
        t.release(pipe)


@pytest.mark.parametrize("dtype", [np.int32, np.uint64, np.float32, np.float64, np.complex64])
@pytest.mark.parametrize("shape", [(7,), (33, 77), (5, 64, 130), (1 << 21,)])
def test_async_reductions_equal_blocking(dtype, shape):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    lib = wk.capi.lib()
    rng = np.random.default_rng(23)
    if np.dtype(dtype).kind == "c":
        data = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dtype)
    else:
        data = gh.rand_data(rng, dtype, shape)
    x = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    y = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    wk.tensor.memory.read_from_buffer(pipe, x, data)
    wk.tensor.memory.read_from_buffer(pipe, y, data[::-1].copy() if len(shape) == 1 else data)
    res = wk.Tensor.alloc(ctx, pipe, (2,), dtype)  # device-resident scalars
    es = x.dtype.itemsize
    host = np.zeros(2, dtype=x.dtype)
    wk.capi.check(lib.wk_sum(pipe.q, x.type_index, x.depth, x.rows, x.row_pitch, x.slice_pitch, x.ptr, host.ctypes.data))
    wk.capi.check(lib.wk_dot_reduce(pipe.q, x.type_index, x.depth, x.rows, x.cols, x.ptr, x.row_pitch, x.slice_pitch, y.ptr,
                                    y.row_pitch, y.slice_pitch, host.ctypes.data + es))
    wk.capi.check(lib.wk_sum_async(pipe.q, x.type_index, x.depth, x.rows, x.row_pitch, x.slice_pitch, x.ptr, res.ptr))
    wk.capi.check(lib.wk_dot_reduce_async(pipe.q, x.type_index, x.depth, x.rows, x.cols, x.ptr, x.row_pitch, x.slice_pitch, y.ptr,
                                          y.row_pitch, y.slice_pitch, res.buffer + es))
    got = gh.padded(res)[:2]
    assert got.tobytes() == host.tobytes()
    for t in (x, y, res):
        t.release(pipe)
