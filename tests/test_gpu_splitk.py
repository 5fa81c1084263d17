"""-m gpu: the split-K path of the f32 tensor-core GEMM (small M*N, long K: output tiles alone cannot fill the chip, so
the K range of every tile is cut over several CTAs and the last one to finish folds the partials in split order).
Checked against the fp64 ideal within the K-scaled bound of SURVEY 8c, for exactness on integer-valued operands
(sums below 2^24 are exact in any association), and for run-to-run determinism (no atomics on data)."""
import numpy as np
import pytest

from . import gpu_helpers as gh

pytestmark = pytest.mark.gpu

SHAPES = [(1024, 1024, 1024), (256, 512, 4096), (100, 300, 2000), (128, 256, 8192), (64, 64, 16384), (384, 1000, 1100),
          (512, 2048, 512),
          # more than one round of tiles with a small last round: only the tail tiles are split (N = 4096: 3 x 74 + 34)
          (4096, 4096, 1024), (2304, 2304, 768)]


@pytest.mark.parametrize("op_a", [0, 1])
@pytest.mark.parametrize("op_b", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_splitk_gemm_vs_ideal(shape, op_a, op_b):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    M, N, K = shape
    rng = np.random.default_rng(M + 3 * N + 7 * K + op_a * 2 + op_b)
    dt = np.float32
    ad = rng.uniform(-1, 1, (K, M) if op_a else (M, K)).astype(dt)
    bd = rng.uniform(-1, 1, (N, K) if op_b else (K, N)).astype(dt)
    cd = rng.uniform(-1, 1, (M, N)).astype(dt)
    a, b = (wk.Tensor.alloc(ctx, pipe, d.shape, dt) for d in (ad, bd))
    wk.tensor.memory.read_from_buffer(pipe, a, ad)
    wk.tensor.memory.read_from_buffer(pipe, b, bd)
    for alpha, beta in [(None, None), (1.25, None), (0.75, 0.5)]:
        outs = []
        for _ in range(2):
            c = wk.Tensor.alloc(ctx, pipe, (M, N), dt)
            wk.tensor.memory.read_from_buffer(pipe, c, cd)
            wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
            outs.append(gh.to_np(c))
            c.release(pipe)
        assert outs[0].tobytes() == outs[1].tobytes(), "split-K result differs between two identical launches"
        ideal = gh.gemm_ideal(ad, op_a, bd, op_b, alpha, beta, cd)
        bound = gh.gemm_float_bound(ad, op_a, bd, op_b, alpha, beta, cd, tol=8.0)
        err = np.abs(outs[0].astype(np.float64) - ideal)
        assert np.all(err <= bound), f"{shape} {op_a}{op_b} alpha={alpha} beta={beta}: max err/bound {np.max(err / bound)}"
    a.release(pipe)
    b.release(pipe)


@pytest.mark.parametrize("shape", [(256, 256, 4096), (1024, 1024, 1024)])
def test_splitk_exact_on_integer_operands(shape):
    """operands in {-2..2}: every product and every partial sum is an integer below 2^24 -> exact in fp32 whatever the
    association, so the split-K result must EQUAL the integer product"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    M, N, K = shape
    rng = np.random.default_rng(99)
    ad = rng.integers(-2, 3, (M, K)).astype(np.float32)
    bd = rng.integers(-2, 3, (K, N)).astype(np.float32)
    a, b, c = (wk.Tensor.alloc(ctx, pipe, s, np.float32) for s in ((M, K), (K, N), (M, N)))
    wk.tensor.memory.read_from_buffer(pipe, a, ad)
    wk.tensor.memory.read_from_buffer(pipe, b, bd)
    wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    want = (ad.astype(np.int64) @ bd.astype(np.int64)).astype(np.float32)
    np.testing.assert_array_equal(gh.to_np(c), want)
    for t in (a, b, c):
        t.release(pipe)


@pytest.mark.parametrize("act", ["sigmoid", "tanh"])
def test_splitk_fused_linear_forward(act):
    """Linear.forward's fused epilogue (bias + activation) applied by the CTA that folds the split-K partials"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    batch, n_in, n_out = 64, 4096, 512
    rng = np.random.default_rng(5)
    xd = rng.uniform(-1, 1, (batch, n_in)).astype(np.float32)
    wd = (rng.uniform(-1, 1, (n_out, n_in)) / np.sqrt(n_in)).astype(np.float32)
    bd = rng.uniform(-1, 1, n_out).astype(np.float32)
    x, w, b, y = (wk.Tensor.alloc(ctx, pipe, s, np.float32) for s in ((batch, n_in), (n_out, n_in), (n_out,), (batch, n_out)))
    for t, d in ((x, xd), (w, wd), (b, bd)):
        wk.tensor.memory.read_from_buffer(pipe, t, d)
    lib = wk.capi.lib()
    wk.capi.check(lib.wk_gemm_bias_act(pipe.q, 8, 0, 1, batch, n_out, n_in, x.ptr, x.row_pitch, w.ptr, w.row_pitch, y.ptr,
                                       y.row_pitch, b.ptr, 1 if act == "sigmoid" else 2))
    z = xd.astype(np.float64) @ wd.astype(np.float64).T + bd
    want = 1 / (1 + np.exp(-z)) if act == "sigmoid" else np.tanh(z)
    np.testing.assert_allclose(gh.to_np(y), want, rtol=0, atol=(8 * n_in + 64) * np.finfo(np.float32).eps)
    for t in (x, w, b, y):
        t.release(pipe)


def test_splitk_graph_survives_workspace_growth():
    """a captured split-K GEMM keeps the address of its partial-tile workspace in the graph; a larger problem that grows
    the workspace afterwards must retire the old buffer, not free it under the graph"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(3)
    dt = np.float32

    def make(M, N, K):
        ad, bd = rng.uniform(-1, 1, (M, K)).astype(dt), rng.uniform(-1, 1, (K, N)).astype(dt)
        a, b, c = (wk.Tensor.alloc(ctx, pipe, s, dt) for s in ((M, K), (K, N), (M, N)))
        wk.tensor.memory.read_from_buffer(pipe, a, ad)
        wk.tensor.memory.read_from_buffer(pipe, b, bd)
        return ad, bd, a, b, c

    ad, bd, a, b, c = make(256, 256, 2048)
    wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)  # warm-up outside the capture
    first = gh.to_np(c)
    pipe.begin_capture()
    wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    graph = pipe.end_capture()
    ad2, bd2, a2, b2, c2 = make(1024, 1536, 4096)   # 4 x 6 = 24 tiles x 3 splits: a much larger workspace
    wk.blas.gemm(pipe, None, a2, 0, b2, 0, None, c2)
    big = gh.to_np(c2)
    assert np.allclose(big, ad2.astype(np.float64) @ bd2.astype(np.float64), rtol=0, atol=4096 * 8 * np.finfo(dt).eps * 4)
    wk.capi.check(wk.capi.lib().wk_memset_zero(pipe.q, c.ptr, c.size))
    graph.launch(pipe)
    np.testing.assert_array_equal(gh.to_np(c), first)
    graph.release()
    for t in (a, b, c, a2, b2, c2):
        t.release(pipe)
