"""-m gpu: random.uniform past 2^32 elements -- the hash takes the PADDED linear index as a 64-bit number whose high word
lands in the LOW half of the mixer's state (uniform.cl:38-41), and the dense kernel recomputes its per-thread invariants when
that word changes (csrc/elementwise.cu: uniform_dense_kernel<..., SMALL = false>).  A 4.3 GB uint8 tensor (65600 x 65536) is
drawn on the device; rows on both sides of index 2^32 are read back and compared with the hash evaluated on the host in plain
64-bit arithmetic (numpy uint64 wraps mod 2^64), bit for bit.  The same rows of a tensor just below 2^32 elements (SMALL = true)
and a padded (pitched) one that takes the generic map kernel are checked the same way."""
import ctypes as C

import numpy as np
import pytest

from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

U = np.uint64
C64 = U(0x9FB21C651E98DF25)


def _rotl(x, k):
    return (x << U(k)) | (x >> U(64 - k))


def host_hash(index, seed):
    """uniform.cl:32-54, little-endian branch, restated on numpy uint64 (wrap-around arithmetic)"""
    with np.errstate(over="ignore"):
        key = U(0x7C01812CF721AD1C ^ 0xDED46DE9839097DB) - U(seed)
        combined = ((index & U(0xFFFFFFFF)) << U(32)) + (index >> U(32))
        x0 = combined ^ key
        x1 = x0 ^ _rotl(x0, 49) ^ (_rotl(x0, 24) * C64)
        x2 = x1 ^ (((x1 >> U(35)) + U(8)) * C64)
        return x2 ^ (x2 >> U(28))


def _rows(pipe, t, r0, n_rows):
    wk = gh.wk()
    out = np.empty((n_rows, t.cols), dtype=t.dtype)
    es = np.dtype(t.dtype).itemsize
    wk.capi.check(wk.capi.lib().wk_d2h_rect(pipe.q, out.ctypes.data_as(C.c_void_p), C.c_void_p(t.buffer + r0 * t.row_pitch * es),
                                           t.row_pitch * es, t.slice_pitch * es, t.cols * es, n_rows, 1))
    wk.capi.check(wk.capi.lib().wk_queue_finish(pipe.q))
    return out


@pytest.mark.parametrize("rows,cols,probe", [
    (65600, 65536, [0, 65535, 65536, 65599]),  # 2^32 falls between rows 65535 and 65536: SMALL = false
    (65535, 65536, [0, 40000, 65534]),         # just below 2^32 elements: SMALL = true
    (70000, 65534, [0, 65537, 65538, 69999]),  # row pitch 65536 > cols: padded index, generic pitched kernel, crosses 2^32
])
def test_uniform_u8_across_2_pow_32(rows, cols, probe):
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    t = wk.Tensor.alloc(ctx, pipe, (rows, cols), np.uint8)
    seed = 0x1234ABCD5678
    wk.tensor.random.uniform(pipe, t, seed)
    for r in probe:
        got = _rows(pipe, t, r, 1)[0]
        idx = U(r) * U(t.row_pitch) + np.arange(cols, dtype=U)
        want = (host_hash(idx, seed) & U(0xFF)).astype(np.uint8)
        np.testing.assert_array_equal(got, want, err_msg=f"row {r}")
    t.release(pipe)


def test_uniform_f32_range_small_vs_host():
    """dense f32 with bounds: min + (double)h / 2^64 * range evaluated on the host in float64, cast to float32"""
    wk = gh.wk()
    ctx, pipe = gh.ctx_pipe()
    t = wk.Tensor.alloc(ctx, pipe, (300, 1024), np.float32)
    wk.tensor.random.uniform(pipe, t, 99, -2.0, 3.0)
    got = gh.to_np(t)
    idx = (np.arange(300, dtype=U)[:, None] * U(t.row_pitch) + np.arange(1024, dtype=U)[None, :])
    h = host_hash(idx, 99)
    want = (np.float64(-2.0) + (h.astype(np.float64) / 18446744073709551616.0) * np.float64(np.float32(3.0) - np.float32(-2.0))).astype(np.float32)
    # the device may contract min + n * range into one FMA (as the OpenCL compiler may): allow the last bit
    assert np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) <= 2.4e-7 * 3.0
    assert got.min() >= -2.0 and got.max() <= 3.0
    t.release(pipe)
