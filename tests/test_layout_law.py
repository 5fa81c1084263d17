"""-m "not gpu": the padded / pitched layout law of Tensor(T) (src/tensor/main.zig:142-222) in the Python host mirror against
the oracle's restatement, for random shapes, every dtype, vectors on / off and fake devices of every vector width the
reference can meet (1 on the CUDA backend, 4 / 8 / 16 on OpenCL CPUs) -- no device memory is touched: the tensors wrap a
dummy pointer."""
import numpy as np
import pytest

import wekua_b200 as wk
from oracle import pyoracle as o


class _FakeQueue:
    def __init__(self, widths):
        self.vector_widths = list(widths)


class _FakeContext:
    def __init__(self, widths_per_queue):
        self.command_queues = [_FakeQueue(w) for w in widths_per_queue]


def _oracle_device(widths):
    d = o.device("cpu", 16)
    for i, w in enumerate(widths):
        d.vector_widths[i] = w
    return d


@pytest.mark.parametrize("vw_f32", [1, 4, 8, 16])
def test_layout_matches_the_oracle(vw_f32):
    rng = np.random.default_rng(100 + vw_f32)
    # OpenCL preferred widths scale with the element size: bytes per vector constant (clamped to >= 1)
    widths = [max(1, vw_f32 * 4 // np.dtype(t).itemsize) if vw_f32 > 1 else 1 for t in wk.core.REAL_TYPES]
    widths = [min(16, w) for w in widths]  # command_queue.zig:236-423: clamp
    ctx = _FakeContext([widths, [1] * 10])  # the tensor takes the MAX over the context's queues (main.zig:145-148)
    dev = _oracle_device(widths)
    for _ in range(300):
        ndim = int(rng.integers(1, 5))
        shape = tuple(int(v) for v in rng.integers(1, 40, ndim))
        ti = int(rng.integers(0, 20))
        dtype = wk.core.SUPPORTED_TYPES[ti]
        vec = bool(rng.integers(0, 2))
        t = wk.Tensor.empty(ctx, None, shape, dtype, vectors_enabled=vec, _external_ptr=0x1000)
        ot = o.OTensor(dev, dtype, shape, vectors_enabled=vec)
        L = ot.layout
        assert (t.row_pitch, t.slice_pitch, t.number_of_elements, t.number_of_elements_without_padding) == \
            (L.row_pitch, L.slice_pitch, L.number_of_elements, L.number_of_elements_without_padding), (shape, ti, vec)
        assert list(t.pitches) == [int(L.pitches[i]) for i in range(ndim)], (shape, ti, vec)
        assert list(t.vl_shape) == [int(L.vl_shape[i]) for i in range(ndim)], (shape, ti, vec)
        assert t.size == t.number_of_elements * np.dtype(wk.core.storage_dtype(dtype)).itemsize


def test_one_dimensional_tensors_allocate_twice_their_length():
    """main.zig:165-168,192-193 (SURVEY layout note): a 1-D tensor has one padded 'row' of 2 => 2x the logical size"""
    ctx = _FakeContext([[1] * 10])
    t = wk.Tensor.empty(ctx, None, (1 << 20,), np.float32, _external_ptr=0x1000)
    assert (t.rows, t.rows_padded, t.row_pitch, t.number_of_elements) == (1, 2, 1 << 20, 2 << 20)
    with pytest.raises(wk.capi.InvalidValue):
        wk.Tensor.empty(ctx, None, (4, 0), np.float32, _external_ptr=0x1000)


def _t(ctx, shape, dtype=np.float32):
    return wk.Tensor.empty(ctx, None, shape, dtype, _external_ptr=0x1000)


def test_operator_argument_validation_happens_on_the_host():
    """gemm.zig:900-941 (InvalidValue / UnqualTensorsContext) and axpy.zig:763-815 (shape mismatch): the host mirror rejects
    bad arguments before anything is enqueued -- exercised here on tensors that wrap a dummy pointer, pipeline = None"""
    ctx, other = _FakeContext([[1] * 10]), _FakeContext([[1] * 10])
    a, b, c = _t(ctx, (6, 10)), _t(ctx, (10, 4)), _t(ctx, (6, 4))
    bad = [
        (a, 0, _t(ctx, (9, 4)), 0, c),      # K mismatch
        (a, 0, b, 0, _t(ctx, (6, 5))),      # C columns
        (a, 0, b, 0, _t(ctx, (5, 4))),      # C rows
        (a, 1, b, 0, c),                    # op_a = T makes A [K, M] = [6, 10]: M = 10 != 6
        (a, 0, b, 1, c),                    # op_b = T makes B [N, K] = [10, 4]: K = 4 != 10
        (_t(ctx, (2, 6, 10)), 0, b, 0, c),  # not a matrix
    ]
    for (x, oa, y, ob, z) in bad:
        with pytest.raises(wk.capi.InvalidValue):
            wk.blas.gemm(None, None, x, oa, y, ob, None, z)
    with pytest.raises(wk.capi.UnqualTensorsContext):
        wk.blas.gemm(None, None, a, 0, _t(other, (10, 4)), 0, None, c)
    with pytest.raises(wk.capi.InvalidValue):
        wk.blas.gemm(None, None, a, 0, _t(ctx, (10, 4), np.float64), 0, None, c)  # dtype mismatch
    # transposed operands that DO match pass validation (and only then reach the library): checked through the validator
    wk.blas._validate_gemm(_t(ctx, (10, 6)), _t(ctx, (4, 10)), c, 1, 1)
    with pytest.raises(wk.capi.WekuaError):
        wk.blas.axpy(None, _t(ctx, (5,)), 2.0, _t(ctx, (6,)))
    with pytest.raises(wk.capi.WekuaError):
        wk.blas.axpy(None, _t(ctx, (2, 3)), 2.0, _t(ctx, (3, 2)))
