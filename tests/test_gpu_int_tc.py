"""-m gpu: integer blas.gemm on the tensor cores (csrc/gemm_i8_tc.cu: tcgen05.mma kind::i8 over byte planes) -- bit-exact.

The reference computes integer GEMMs in the element type (src/blas/kernels/gemm_nxn_gpu.cl:82-319; OpenCL lanes wrap), so
the result is exact arithmetic mod 2^bits and any evaluation order must reproduce it bit for bit.  Checked here, with the
tensor-core path FORCED (wk_gemm_set_path(2): fail rather than fall back):
  * every integer dtype x transpose pair x (alpha, beta) variant on ragged shapes against the oracle's restated kernels
    and against numpy's wrap-around integer matmul;
  * K ranges that need several s32-safe chunks (i8: K > 32768; i32: 4 planes x K > 32768);
  * at BASELINE's sweep size (N = 8192): A.I == A, and the Freivalds identity C.w == A.(B.w) mod 2^bits on the host;
  * the automatic choice (path 0) takes the same kernel for big problems and agrees with the SIMT kernel (path 1).
"""
import numpy as np
import pytest

from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

INTS = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64]


def _wrap_matmul(a, op_a, b, op_b, alpha, beta, c0):
    """exact mod 2^bits: numpy integer matmul wraps in uint64 (mod 2^64), truncation to the element type keeps the residue"""
    dt = c0.dtype
    A = (a.T if op_a else a).astype(np.uint64)
    B = (b.T if op_b else b).astype(np.uint64)
    r = A @ B
    if alpha is not None or beta is not None:
        r = np.uint64(np.int64(1 if alpha is None else alpha).astype(np.uint64)) * r
    if beta is not None:
        r = r + np.int64(beta).astype(np.uint64) * c0.astype(np.uint64)
    return r.astype(dt)


def _force(path):
    gh.wk().capi.check(gh.wk().capi.lib().wk_gemm_set_path(path))


@pytest.fixture
def tc_path():
    _force(2)
    yield
    _force(0)


RAGGED = [(1, 1, 1), (2, 3, 5), (17, 33, 9), (64, 64, 64), (100, 130, 70), (129, 257, 65), (256, 128, 192), (300, 520, 1000)]


@pytest.mark.parametrize("dtype", INTS)
@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_int_gemm_tensor_core_vs_oracle_and_numpy(oracle, tc_path, dtype, op_a, op_b):
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(4321 + op_a * 2 + op_b)
    for (M, N, K) in RAGGED:
        variants = [(None, None), (3, None), (5, 7), (None, 2)]
        if np.dtype(dtype).kind == "i":
            variants.append((-3, -2))  # sign-extended scalars
        for alpha, beta in variants:
            a_shape = (K, M) if op_a else (M, K)
            b_shape = (N, K) if op_b else (K, N)
            ad, bd, cd = (gh.rand_data(rng, dtype, s) for s in (a_shape, b_shape, (M, N)))
            a, oa = gh.make_pair(oracle, dtype, a_shape, ad)
            b, ob = gh.make_pair(oracle, dtype, b_shape, bd)
            c, oc = gh.make_pair(oracle, dtype, (M, N), cd)
            launches = wk.capi.launch_count()
            wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
            assert wk.capi.launch_count() > launches
            got = gh.to_np(c)
            np.testing.assert_array_equal(got, _wrap_matmul(ad, op_a, bd, op_b, alpha, beta, cd))
            if M * N * K <= 300000:  # the oracle's scalar loops: small shapes only
                oracle.gemm(alpha, oa, op_a, ob, op_b, beta, oc)
                np.testing.assert_array_equal(got, oc.to_host())
            for t in (a, b, c):
                t.release(pipe)


def test_int_gemm_padding_of_c_untouched(oracle, tc_path):
    """odd shapes: C's pad column / pad row keep their contents (the epilogue stores logical elements only)"""
    wk = gh.wk()
    _, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(5)
    for dtype in (np.int8, np.int32):
        M, N, K = 37, 45, 131
        a, _ = gh.make_pair(oracle, dtype, (M, K), gh.rand_data(rng, dtype, (M, K)))
        b, _ = gh.make_pair(oracle, dtype, (K, N), gh.rand_data(rng, dtype, (K, N)))
        c, _ = gh.make_pair(oracle, dtype, (M, N))
        before = gh.padded(c).copy()
        wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
        after = gh.padded(c).reshape(c.rows_padded, c.row_pitch)
        before = before.reshape(c.rows_padded, c.row_pitch)
        np.testing.assert_array_equal(after[:, N:], before[:, N:])
        np.testing.assert_array_equal(after[M:, :], before[M:, :])
        for t in (a, b, c):
            t.release(pipe)


@pytest.mark.parametrize("dtype,shape", [
    (np.int8, (130, 70, 40000)),     # one plane, K > 32768: two chunks fold into C with beta' = 1
    (np.uint8, (64, 300, 70001)),    # three chunks, K not a multiple of 16 (operands staged)
    (np.int16, (70, 130, 20000)),    # group 1 = 2 planes x 20096 > 32768: pairs walked one by one
    (np.int32, (130, 66, 9000)),     # 4 planes x 9088 > 32768 for groups 3; 2 x 9088 fits
    (np.uint64, (40, 72, 5000)),     # 8 planes: 36 byte products per element pair
])
def test_int_gemm_long_k_is_chunked_exactly(tc_path, dtype, shape):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    M, N, K = shape
    rng = np.random.default_rng(11)
    ad, bd, cd = (gh.rand_data(rng, dtype, s) for s in ((M, K), (N, K), (M, N)))
    a, b, c = (wk.Tensor.alloc(ctx, pipe, s, dtype) for s in ((M, K), (N, K), (M, N)))
    for t, d in ((a, ad), (b, bd), (c, cd)):
        wk.tensor.memory.read_from_buffer(pipe, t, d)
    wk.blas.gemm(pipe, 3, a, 0, b, 1, 5, c)
    np.testing.assert_array_equal(gh.to_np(c), _wrap_matmul(ad, 0, bd, 1, 3, 5, cd))
    for t in (a, b, c):
        t.release(pipe)


@pytest.mark.parametrize("dtype", [np.int8, np.uint8, np.int16, np.int32, np.uint64])
def test_int_gemm_full_size_identity_and_freivalds(dtype):
    """N = 8192 (the sweep size of profiles/sweep_gemm_int_r02.md), automatic path: A.I == A bit for bit, and for a random
    integer vector w: C.w == A.(B.w) in arithmetic mod 2^bits (uint64 wrap-around on the host, truncated)"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    n = 8192 if np.dtype(dtype).itemsize < 8 else 4096
    a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
    wk.tensor.random.uniform(pipe, a, 42)
    wk.tensor.random.uniform(pipe, b, 43)
    ah, bh = gh.to_np(a), gh.to_np(b)
    assert len(np.unique(ah[:64])) > 16  # the operands really cover the integer range
    for op_a, op_b in ((0, 0), (1, 1)):
        wk.blas.gemm(pipe, None, a, op_a, b, op_b, None, c)
        ch = gh.to_np(c)
        rng = np.random.default_rng(3)
        w = rng.integers(0, 1 << 63, n, dtype=np.uint64)
        A = (ah.T if op_a else ah).astype(np.uint64)
        B = (bh.T if op_b else bh).astype(np.uint64)
        want = (A @ (B @ w)).astype(dtype)
        got = (ch.astype(np.uint64) @ w).astype(dtype)
        np.testing.assert_array_equal(got, want)
    wk.tensor.identity(pipe, b)
    wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
    np.testing.assert_array_equal(gh.to_np(c), ah)
    for t in (a, b, c):
        t.release(pipe)


@pytest.mark.parametrize("dtype", [np.int8, np.uint16, np.int32, np.int64])
def test_int_gemm_auto_path_matches_simt(dtype):
    """1024 x 1024 x 1024 is past the switch-over (2^29 multiply-adds): path 0 and the SIMT kernel (path 1) agree bit for bit"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    n = 1024
    a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), dtype) for _ in range(3))
    wk.tensor.random.uniform(pipe, a, 7)
    wk.tensor.random.uniform(pipe, b, 8)
    res = []
    for path in (0, 1, 2):
        _force(path)
        try:
            wk.tensor.random.uniform(pipe, c, 9)
            wk.blas.gemm(pipe, 3, a, 1, b, 0, 2, c)
            res.append(gh.to_np(c))
        finally:
            _force(0)
    np.testing.assert_array_equal(res[0], res[1])
    np.testing.assert_array_equal(res[2], res[1])
    for t in (a, b, c):
        t.release(pipe)


def test_int_gemm_tensor_core_inside_a_cuda_graph(tc_path):
    """capture / replay: the byte-plane staging, every chunk launch and the workspace (allocated inside the capture on first use,
    never freed while the graph may replay) must be graph-safe; replays with new operand contents give the new product"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(21)
    M, N, K = 300, 260, 1100
    for dtype in (np.int8, np.int32):
        a, b, c = (wk.Tensor.alloc(ctx, pipe, s, dtype) for s in ((K, M), (K, N), (M, N)))
        pipe.wait_and_cleanup()
        pipe.begin_capture()
        wk.blas.gemm(pipe, 3, a, 1, b, 0, None, c)
        g = pipe.end_capture()
        for rep in range(2):
            ad, bd = gh.rand_data(rng, dtype, (K, M)), gh.rand_data(rng, dtype, (K, N))
            wk.tensor.memory.read_from_buffer(pipe, a, ad)
            wk.tensor.memory.read_from_buffer(pipe, b, bd)
            g.launch(pipe)
            pipe.wait_and_cleanup()
            np.testing.assert_array_equal(gh.to_np(c), _wrap_matmul(ad, 1, bd, 0, 3, None, np.zeros((M, N), dtype)))
        g.release()
        for t in (a, b, c):
            t.release(pipe)
