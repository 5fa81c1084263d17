"""-m gpu: parity at BASELINE.json's FULL sizes through size-independent properties (the oracle finishes only small
cases in seconds): Freivalds-style checksums C.w == A.(B.w) in fp64 on the host, exactness of A.I for operands whose
hi/lo split is exact, exact scaling by powers of two, and bit-exact axpy on 2^28-element vectors against numpy
(separate multiply and add, like the -ffp-contract=off oracle and the -fmad=false kernels)."""
import numpy as np
import pytest

from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


def _matvec64(m: np.ndarray, v: np.ndarray, transpose=False, absolute=False, chunk=2048) -> np.ndarray:
    """fp64 m @ v (or m.T @ v) in row chunks so a 4 GiB f32 matrix never needs an 8 GiB fp64 copy"""
    rows = m.shape[0]
    out = np.zeros(m.shape[1] if transpose else rows, dtype=np.float64)
    for r0 in range(0, rows, chunk):
        blk = m[r0:r0 + chunk].astype(np.float64)
        if absolute:
            np.abs(blk, out=blk)
        if transpose:
            out += blk.T @ v[r0:r0 + chunk]
        else:
            out[r0:r0 + chunk] = blk @ v
    return out


def _device_random(wk, ctx, pipe, shape, dtype, seed):
    t = wk.Tensor.alloc(ctx, pipe, shape, dtype)
    wk.tensor.random.uniform(pipe, t, seed, -1, 1)  # the reference's counter-based PRNG (uniform.cl)
    return t


@pytest.mark.parametrize("dtype,n,op_a,op_b,alpha,beta", [
    (np.float32, 16384, 0, 0, None, None),     # BASELINE config 4
    (np.float32, 16384, 1, 1, 0.75, 0.5),
    (np.float32, 8192, 0, 1, 1.25, None),
    (np.float32, 8192, 1, 0, None, None),
    (np.float64, 16384, 0, 0, None, None),
    (np.float64, 8192, 1, 1, 0.75, 0.5),
    (np.float64, 4096, 0, 1, 1.25, None),
    (np.float32, 32768, 0, 0, None, None),     # BASELINE config 5 (single-GPU leg)
])
def test_gemm_full_size_checksum(dtype, n, op_a, op_b, alpha, beta):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    a = _device_random(wk, ctx, pipe, (n, n), dtype, 42)
    b = _device_random(wk, ctx, pipe, (n, n), dtype, 43)
    c = _device_random(wk, ctx, pipe, (n, n), dtype, 44)
    c0 = gh.to_np(c) if beta is not None else None
    wk.blas.gemm(pipe, alpha, a, op_a, b, op_b, beta, c)
    ch = gh.to_np(c)
    ah, bh = gh.to_np(a), gh.to_np(b)
    for t in (a, b, c):
        t.release(pipe)
    rng = np.random.default_rng(1)
    eps = float(np.finfo(dtype).eps)
    tol = 8 if dtype == np.float32 else 4
    al = 1.0 if alpha is None else float(dtype(alpha))
    for _ in range(2):
        w = rng.uniform(-1, 1, n)
        # right checksum: C w = alpha * op(A) (op(B) w) + beta * C0 w
        bw = _matvec64(bh, w) if op_b == 0 else _matvec64(bh, w, transpose=True)
        ideal = al * (_matvec64(ah, bw) if op_a == 0 else _matvec64(ah, bw, transpose=True))
        babs = _matvec64(bh, np.abs(w), absolute=True) if op_b == 0 else _matvec64(bh, np.abs(w), transpose=True, absolute=True)
        bound = abs(al) * (_matvec64(ah, babs, absolute=True) if op_a == 0 else _matvec64(ah, babs, transpose=True, absolute=True))
        if beta is not None:
            ideal += float(dtype(beta)) * _matvec64(c0, w)
            bound += abs(beta) * _matvec64(c0, np.abs(w), absolute=True)
        got = _matvec64(ch, w)
        err = np.abs(got - ideal)
        limit = (tol * n + 16) * eps * bound
        assert np.all(err <= limit), f"checksum off: max err/limit {float((err / limit).max()):.3g}"
        assert float(np.abs(ideal).max()) > 1.0  # the checksum is not trivially zero


@pytest.mark.parametrize("n,op_b", [(8192, 0), (4096, 1)])
def test_gemm_full_size_identity_is_exact_f32(n, op_b):
    """A.I == A bit for bit when the 3xTF32 split of A is exact (21 significant bits): every tile, k-block and lane of the
    tensor-core path carries the value through unchanged"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    rng = np.random.default_rng(2)
    ad = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    ad = (ad.view(np.uint32) & np.uint32(0xFFFFFFF8)).view(np.float32)  # 21-bit significands: hi (11) + lo (10) exactly
    a = wk.Tensor.alloc(ctx, pipe, (n, n), np.float32)
    i = wk.Tensor.alloc(ctx, pipe, (n, n), np.float32)
    c = wk.Tensor.alloc(ctx, pipe, (n, n), np.float32)
    wk.tensor.memory.read_from_buffer(pipe, a, ad)
    wk.tensor.identity(pipe, i)
    wk.blas.gemm(pipe, None, a, 0, i, op_b, None, c)
    got = gh.to_np(c)
    for t in (a, i, c):
        t.release(pipe)
    assert np.array_equal(got, ad)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gemm_full_size_power_of_two_scaling_is_exact(dtype):
    """alpha = 2 scales every element exactly: gemm(2, A, B) == 2 * gemm(A, B) bit for bit at N = 8192"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    n = 8192
    a = _device_random(wk, ctx, pipe, (n, n), dtype, 42)
    b = _device_random(wk, ctx, pipe, (n, n), dtype, 43)
    c1 = wk.Tensor.alloc(ctx, pipe, (n, n), dtype)
    c2 = wk.Tensor.alloc(ctx, pipe, (n, n), dtype)
    wk.blas.gemm(pipe, None, a, 0, b, 1, None, c1)
    wk.blas.gemm(pipe, 2, a, 0, b, 1, None, c2)
    r1, r2 = gh.to_np(c1), gh.to_np(c2)
    for t in (a, b, c1, c2):
        t.release(pipe)
    assert np.array_equal(r2, dtype(2) * r1)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_axpy_full_size_bit_exact(dtype):
    """BASELINE config 2: 2^28-element vectors, the benchmark's alternating x/y scheme (8 steps, |alpha| < 1), bit-exact
    against numpy's separate multiply and add"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    n = 1 << 28
    x = wk.Tensor.alloc(ctx, pipe, (n,), dtype)
    y = wk.Tensor.alloc(ctx, pipe, (n,), dtype)
    wk.tensor.random.uniform(pipe, x, 42)
    wk.tensor.random.uniform(pipe, y, 43)
    xh, yh = gh.to_np(x).copy(), gh.to_np(y).copy()
    alphas = np.random.default_rng(1234).uniform(-1, 1, 8) / np.sqrt(2)
    for i, al in enumerate(alphas):
        al = dtype(al)
        if i % 2 == 0:
            wk.blas.axpy(pipe, x, al, y)
            yh += al * xh
        else:
            wk.blas.axpy(pipe, y, al, x)
            xh += al * yh
    gx, gy = gh.to_np(x), gh.to_np(y)
    x.release(pipe)
    y.release(pipe)
    assert np.array_equal(gx, xh) and np.array_equal(gy, yh)
    assert np.all(np.isfinite(gx)) and np.all(np.isfinite(gy))
