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
