"""-m gpu: float GEMM operands the tensor-core loaders cannot address directly (TMA and 16-byte cp.async need 16-byte aligned
bases and row pitches) are staged into an aligned scratch copy and still run on the tensor cores (csrc/gemm.cu: stage_aligned).

With vector width 1 the reference's layout law gives row_pitch = cols + cols % 2 (src/tensor/main.zig:174-187), so every f32
tensor with cols % 4 == 2 has such a pitch; f64 rows are always 16-byte multiples, but a column-block VIEW that starts at an
odd column is not aligned.  The tensor-core path is forced (wk_gemm_set_path(2): an ineligible operand would be an error, not
a silent SIMT fallback); results against the float64 product within the K-scaled bound."""
import numpy as np
import pytest

from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


@pytest.fixture
def tc_path():
    lib = gh.wk().capi.lib()
    gh.wk().capi.check(lib.wk_gemm_set_path(2))
    yield
    gh.wk().capi.check(lib.wk_gemm_set_path(0))


def _check(got, ad, op_a, bd, op_b, alpha, beta, cd):
    A = (ad.T if op_a else ad).astype(np.float64)
    B = (bd.T if op_b else bd).astype(np.float64)
    want = (1.0 if alpha is None else alpha) * (A @ B) + (0.0 if beta is None else beta) * cd.astype(np.float64)
    bound = gh.gemm_float_bound(ad, op_a, bd, op_b, alpha, beta, cd, tol=1.0)
    err = np.abs(got.astype(np.float64) - want)
    assert np.all(err <= bound), float(np.max(err / bound))


@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_f32_pitch_not_multiple_of_four(tc_path, op_a, op_b):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    M, N, K = 514, 1026, 518  # every operand's cols % 4 == 2
    rng = np.random.default_rng(5 + 2 * op_a + op_b)
    a_shape = (K, M) if op_a else (M, K)
    b_shape = (N, K) if op_b else (K, N)
    for alpha, beta in ((None, None), (0.75, 0.5)):
        ad, bd, cd = (rng.uniform(-1, 1, s).astype(np.float32) for s in (a_shape, b_shape, (M, N)))
        a, b, c = (wk.Tensor.alloc(ctx, pipe, s, np.float32) for s in (a_shape, b_shape, (M, N)))
        assert a.row_pitch % 4 == 2 and b.row_pitch % 4 == 2
        for t, d in ((a, ad), (b, bd), (c, cd)):
            wk.tensor.memory.read_from_buffer(pipe, t, d)
        wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
        _check(gh.to_np(c), ad, op_a, bd, op_b, alpha, beta, cd)
        for t in (a, b, c):
            t.release(pipe)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (1, 1)])
def test_column_block_views_at_odd_offsets(tc_path, dtype, op_a, op_b):
    """A and B are column blocks of wider tensors starting at column 1: base pointers 4 / 8 bytes past a 16-byte boundary"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    M, N, K = 520, 1024, 516
    rng = np.random.default_rng(9)
    a_shape = (K, M) if op_a else (M, K)
    b_shape = (N, K) if op_b else (K, N)
    es = np.dtype(dtype).itemsize
    wide_a = wk.Tensor.alloc(ctx, pipe, (a_shape[0], a_shape[1] + 8), dtype)
    wide_b = wk.Tensor.alloc(ctx, pipe, (b_shape[0], b_shape[1] + 8), dtype)
    wad = rng.uniform(-1, 1, wide_a.shape).astype(dtype)
    wbd = rng.uniform(-1, 1, wide_b.shape).astype(dtype)
    wk.tensor.memory.read_from_buffer(pipe, wide_a, wad)
    wk.tensor.memory.read_from_buffer(pipe, wide_b, wbd)
    a = wk.Tensor.wrap(ctx, pipe, a_shape, dtype, wide_a.buffer + es, row_pitch=wide_a.row_pitch)
    b = wk.Tensor.wrap(ctx, pipe, b_shape, dtype, wide_b.buffer + es, row_pitch=wide_b.row_pitch)
    c = wk.Tensor.alloc(ctx, pipe, (M, N), dtype)
    cd = rng.uniform(-1, 1, (M, N)).astype(dtype)
    wk.tensor.memory.read_from_buffer(pipe, c, cd)
    wk.blas.gemm(pipe, 1.5, a, op_a, b, op_b, -0.25, c)
    ad, bd = wad[:, 1:1 + a_shape[1]], wbd[:, 1:1 + b_shape[1]]
    _check(gh.to_np(c), ad, op_a, bd, op_b, 1.5, -0.25, cd)
    # the staging must not have written into the wide tensors
    np.testing.assert_array_equal(gh.to_np(wide_a), wad)
    np.testing.assert_array_equal(gh.to_np(wide_b), wbd)
    for t in (c, wide_a, wide_b):
        t.release(pipe)
