"""-m gpu: SURVEY section 8f rank 1 -- one Linear layer step (linear.zig:480-678) at a size that runs on the tensor-core
paths, unfused (gemm + bias + activation / act' + hadamard launches, the reference's sequence) and fused (bias +
activation in the GEMM epilogue, act' * sensitivity in one pass), both against the op-by-op oracle run."""
import numpy as np
import pytest

from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("act", ["sigmoid", "tanh"])
@pytest.mark.parametrize("batch,n_in,n_out", [(512, 384, 640), (130, 96, 258)])
def test_linear_forward_backward_vs_oracle(oracle, dtype, fused, act, batch, n_in, n_out):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    o, dev = oracle, gh.oracle_dev(oracle)
    eps = float(np.finfo(dtype).eps)
    rng = np.random.default_rng(21)
    activation = wk.nn.Sigmoid.init() if act == "sigmoid" else wk.nn.Tanh.init()
    lin = wk.nn.Linear.init(ctx, pipe, n_in, n_out, activation, dtype=dtype, seed=42, fused=fused)
    bd = rng.uniform(-0.5, 0.5, n_out).astype(dtype)
    wk.tensor.memory.read_from_buffer(pipe, lin.bias[0], bd)
    xd = rng.uniform(-1, 1, (batch, n_in)).astype(dtype)
    x = wk.Tensor.alloc(ctx, pipe, (batch, n_in), dtype)
    wk.tensor.memory.read_from_buffer(pipe, x, xd)
    cache = lin.prepare_cache(pipe, batch)
    sd = rng.uniform(-1, 1, (batch, n_out)).astype(dtype)
    wk.tensor.memory.read_from_buffer(pipe, cache.sensitivities[0], sd)
    xs = wk.Tensor.alloc(ctx, pipe, (batch, n_in), dtype)  # input sensitivity

    out = lin.forward(pipe, x, cache)
    lin.backward(pipe, cache, x, xs)

    # the oracle, op by op on reference-layout tensors with the same data
    T = lambda shape: o.OTensor(dev, dtype, shape)  # noqa: E731
    W = T((n_out, n_in)).read_from(gh.to_np(lin.weights[0]))
    ob, ox, oout = T((n_out,)).read_from(bd), T((batch, n_in)).read_from(xd), T((batch, n_out))
    osens, odact, ogw, ogb, oxs = T((batch, n_out)).read_from(sd), T((batch, n_out)), T((n_out, n_in)), T((n_out,)), T((batch, n_in))
    o.gemm(None, ox, 0, W, 1, None, oout, packed=True)
    o.bias(oout, ob)
    o.unary(oout, act)
    # forward: |err| <= GEMM bound through an activation with slope <= 1, plus a few ulp of exp/tanh
    fwd_tol = (8 * n_in + 64) * eps * (np.abs(xd).astype(np.float64) @ np.abs(gh.to_np(lin.weights[0])).astype(np.float64).T + np.abs(bd))
    np.testing.assert_array_less(np.abs(gh.to_np(out).astype(np.float64) - oout.to_host()), fwd_tol + 16 * eps)
    (o.sigmoid_dev if act == "sigmoid" else o.tanh_dev)(oout, odact)
    o.hadamard(osens, odact)
    o.gemm(None, osens, 1, ox, 0, None, ogw, packed=True)
    o.bias_step(osens, ogb)
    o.gemm(None, osens, 0, W, 0, None, oxs, packed=True)
    # backward quantities: relative to their own scale (every term is a K-long dot product of O(1) values)
    for got_t, want, k in ((cache.gradients[0], ogw.to_host(), batch), (cache.bias_gradients[0], ogb.to_host(), batch),
                           (xs, oxs.to_host(), n_out)):
        got = gh.to_np(got_t).astype(np.float64)
        scale = np.abs(want).max() + 1.0
        assert np.abs(got - want).max() <= (8 * k + 64) * eps * scale * 4, (np.abs(got - want).max(), scale)
    # the logical region of the padded outputs agrees too when shapes are odd (C-padding semantics of the packed path)
    np.testing.assert_allclose(gh.padded(out).reshape(out.rows_padded, out.row_pitch)[batch:, :], oout.buf.reshape(out.rows_padded, out.row_pitch)[batch:, :],
                               rtol=64 * eps, atol=64 * eps)
    for t in (x, xs):
        t.release(pipe)
    lin.release_cache(pipe, cache)
    lin.deinit(pipe)


@pytest.mark.parametrize("act", ["sigmoid", "tanh"])
@pytest.mark.parametrize("batch,n_in,n_out", [(1000, 520, 392), (256, 1024, 1024), (4100, 260, 132)])
@pytest.mark.parametrize("mode,launches", [(2, (2, 2)), (1, (3, 4))])
def test_linear_backward_fused_launch_count_and_numpy(act, batch, n_in, n_out, mode, launches):
    """wk_linear_backward on f32 layers the tensor-core kernel can address.  Mode 2: exactly TWO kernel launches (act' o
    sensitivity formed in both GEMMs' converter stage, the bias gradient summed in the first); mode 1 (default): one fused
    element-wise + column-sum pass (two kernels when the batch is chunked) and the two GEMMs.  The reference's sequence is five
    launches: sigmoid_dev / tanh_dev, dot, gemm TN, bias_step, gemm NN (linear.zig:608-660).  Results against the fp64 ideal
    of the same formulas within the K-scaled GEMM bound.  Shapes: ragged M / N / K tiles, several tile rows and columns (the
    bias gradient must come from the first tile column only), a K (batch) longer than one pipeline round."""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    lib = wk.capi.lib()
    rng = np.random.default_rng(5)
    dtype = np.float32
    eps = float(np.finfo(dtype).eps)
    sd = rng.uniform(-1, 1, (batch, n_out)).astype(dtype)
    yd = (rng.uniform(0.05, 0.95, (batch, n_out)) if act == "sigmoid" else rng.uniform(-0.95, 0.95, (batch, n_out))).astype(dtype)
    xd = rng.uniform(-1, 1, (batch, n_in)).astype(dtype)
    wd = rng.uniform(-1, 1, (n_out, n_in)).astype(dtype)
    names = {"s": (batch, n_out), "y": (batch, n_out), "x": (batch, n_in), "w": (n_out, n_in), "g": (n_out, n_in), "bg": (n_out,),
             "n": (batch, n_in)}
    t = {k: wk.Tensor.alloc(ctx, pipe, shp, dtype) for k, shp in names.items()}
    for k, d in (("s", sd), ("y", yd), ("x", xd), ("w", wd)):
        wk.tensor.memory.read_from_buffer(pipe, t[k], d)
    kind = 1 if act == "sigmoid" else 2
    pipe.wait_and_cleanup()
    wk.capi.check(lib.wk_linear_backward_set_mode(mode))
    l0 = wk.capi.launch_count()
    wk.capi.check(lib.wk_linear_backward(pipe.q, t["s"].type_index, kind, batch, n_out, n_in, t["s"].ptr, t["s"].row_pitch, t["y"].ptr,
                                         t["y"].row_pitch, t["x"].ptr, t["x"].row_pitch, t["w"].ptr, t["w"].row_pitch, t["g"].ptr,
                                         t["g"].row_pitch, t["bg"].ptr, t["n"].ptr, t["n"].row_pitch))
    pipe.wait_and_cleanup()
    assert launches[0] <= wk.capi.launch_count() - l0 <= launches[1]
    y64, s64 = yd.astype(np.float64), sd.astype(np.float64)
    d64 = y64 * (1 - y64) if act == "sigmoid" else 1 - y64 * y64
    s2 = s64 * d64
    want = {"g": s2.T @ xd.astype(np.float64), "bg": s2.sum(axis=0), "n": s2 @ wd.astype(np.float64)}
    bound = {"g": np.abs(s2).T @ np.abs(xd).astype(np.float64), "bg": np.abs(s2).sum(axis=0), "n": np.abs(s2) @ np.abs(wd).astype(np.float64)}
    kk = {"g": batch, "bg": batch, "n": n_out}
    for k in ("g", "bg", "n"):
        got = gh.to_np(t[k]).astype(np.float64)
        err = np.abs(got - want[k])
        lim = (8 * kk[k] + 64) * eps * bound[k] + 1e-30
        assert np.all(err <= lim), (k, float((err / lim).max()))
    # the same call again gives the same bits (no atomics anywhere on the path); mode 1 consumed the sensitivity: restore it
    first = {k: gh.to_np(t[k]).copy() for k in ("g", "bg", "n")}
    wk.tensor.memory.read_from_buffer(pipe, t["s"], sd)
    wk.capi.check(lib.wk_linear_backward(pipe.q, t["s"].type_index, kind, batch, n_out, n_in, t["s"].ptr, t["s"].row_pitch, t["y"].ptr,
                                         t["y"].row_pitch, t["x"].ptr, t["x"].row_pitch, t["w"].ptr, t["w"].row_pitch, t["g"].ptr,
                                         t["g"].row_pitch, t["bg"].ptr, t["n"].ptr, t["n"].row_pitch))
    for k in ("g", "bg", "n"):
        np.testing.assert_array_equal(gh.to_np(t[k]), first[k])
    wk.capi.check(lib.wk_linear_backward_set_mode(-1))
    # and the op-by-op path (forced SIMT back-end: no fusion) agrees within the same bound
    lib.wk_gemm_set_path(1)
    lib.wk_linear_backward_set_mode(0)
    try:
        s_copy = wk.Tensor.alloc(ctx, pipe, (batch, n_out), dtype)
        wk.tensor.memory.read_from_buffer(pipe, s_copy, sd)
        wk.capi.check(lib.wk_linear_backward(pipe.q, s_copy.type_index, kind, batch, n_out, n_in, s_copy.ptr, s_copy.row_pitch, t["y"].ptr,
                                             t["y"].row_pitch, t["x"].ptr, t["x"].row_pitch, t["w"].ptr, t["w"].row_pitch, t["g"].ptr,
                                             t["g"].row_pitch, t["bg"].ptr, t["n"].ptr, t["n"].row_pitch))
    finally:
        lib.wk_gemm_set_path(0)
        lib.wk_linear_backward_set_mode(-1)
    for k in ("g", "bg", "n"):
        err = np.abs(gh.to_np(t[k]).astype(np.float64) - want[k])
        assert np.all(err <= (8 * kk[k] + 64) * eps * bound[k] + 1e-30), k
    s_copy.release(pipe)
    for x in t.values():
        x.release(pipe)
