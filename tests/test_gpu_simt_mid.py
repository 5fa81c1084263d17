"""-m gpu: the 32 x 64-tile SIMT GEMM for mid-size float problems (csrc/gemm_simt.cu: gemm_simt_mid_kernel).

Forced onto the SIMT path (wk_gemm_set_path(1)), shapes that make the mid-tile kernel eligible (>= 74 tiles of 32 x 64, every
128-bit vector whole) for all four transpose pairs and alpha / beta variants, f32 and f64, against a float64 numpy product
with the K-scaled bound of gpu_helpers.gemm_float_bound; ragged tile edges and a partial last k-slab included.
A shape whose vectors are NOT whole (odd K) must still work (it takes the element-wise kernels)."""
import numpy as np
import pytest

from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

SHAPES = [(320, 512, 100), (324, 516, 52), (512, 512, 512), (96, 3200, 36), (2500, 64, 20), (322, 514, 51)]


@pytest.fixture
def simt_path():
    lib = gh.wk().capi.lib()
    gh.wk().capi.check(lib.wk_gemm_set_path(1))
    yield
    gh.wk().capi.check(lib.wk_gemm_set_path(0))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_simt_mid_tiles_vs_numpy(simt_path, dtype, op_a, op_b):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(77 + 2 * op_a + op_b)
    for (M, N, K) in SHAPES:
        for alpha, beta in ((None, None), (0.75, None), (1.5, -0.5)):
            a_shape = (K, M) if op_a else (M, K)
            b_shape = (N, K) if op_b else (K, N)
            ad, bd, cd = (rng.uniform(-1, 1, s).astype(dtype) for s in (a_shape, b_shape, (M, N)))
            a, b, c = (wk.Tensor.alloc(ctx, pipe, s, dtype) for s in (a_shape, b_shape, (M, N)))
            for t, d in ((a, ad), (b, bd), (c, cd)):
                wk.tensor.memory.read_from_buffer(pipe, t, d)
            wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
            got = gh.to_np(c).astype(np.float64)
            A = (ad.T if op_a else ad).astype(np.float64)
            B = (bd.T if op_b else bd).astype(np.float64)
            want = (1.0 if alpha is None else alpha) * (A @ B) + (0.0 if beta is None else beta) * cd.astype(np.float64)
            bound = gh.gemm_float_bound(ad, op_a, bd, op_b, alpha, beta, cd, tol=1.0)
            assert np.all(np.abs(got - want) <= bound), (M, N, K, alpha, beta, float(np.max(np.abs(got - want) / bound)))
            for t in (a, b, c):
                t.release(pipe)
