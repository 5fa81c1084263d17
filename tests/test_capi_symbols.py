"""not gpu: the C-ABI library loads without a GPU, exports every symbol include/wekua_b200.h declares, reports
'no device' instead of computing on the CPU, and the host-side layout math equals the reference restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "wekua_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from wekua_b200 import capi

    lib = capi.lib()
    syms = _header_symbols()
    assert len(syms) >= 55
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/wekua_b200.h but not exported"
    bound = set(capi.SIGNATURES) | set(capi._NON_STATUS)
    assert set(syms) == bound, (set(syms) ^ bound)


def test_product_never_imports_the_oracle():
    """the product path must not route through oracle/ (or any CPU fallback)"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wekua_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in ("import oracle", "from oracle", "pyoracle", "wekua_oracle", "wko_", "libwekua_oracle"):
                    assert pat not in src, f"{f} references the oracle ({pat})"


def test_no_gpu_means_no_device_error_not_a_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wekua_b200 import Context, capi

    n = C.c_int32(-1)
    rc = capi.lib().wk_device_count(C.byref(n))
    assert rc != 0 and n.value == 0
    with pytest.raises(capi.WekuaError):
        Context.init_from_device_type("all")


class _FakeQueue:
    vector_widths = [1] * 10


class _FakeCtx:
    command_queues = [_FakeQueue()]


@pytest.mark.parametrize("dtype", [np.int8, np.uint16, np.float32, np.float64])
def test_host_layout_math_equals_reference_restatement(oracle, dtype):
    """Tensor.empty's layout (src/tensor/main.zig:142-222) as mirrored in wekua_b200/tensor.py vs the oracle's"""
    from wekua_b200 import Tensor

    dev = oracle.device("b200")
    for shape in [(1,), (5,), (10,), (4, 1), (4, 2), (4, 10), (5, 7), (1024, 1024), (2, 3, 4), (3, 5, 7), (2, 3, 4, 5)]:
        t = Tensor.empty(_FakeCtx(), None, shape, dtype, _external_ptr=0x1000)
        o = oracle.OTensor(dev, dtype, shape).layout
        assert (t.row_pitch, t.row_pitch_for_vectors, t.slice_pitch, t.number_of_elements, t.number_of_vectors) == \
               (o.row_pitch, o.row_pitch_for_vectors, o.slice_pitch, o.number_of_elements, o.number_of_vectors)
        assert t.pitches == list(o.pitches[:len(shape)])
        assert (t.depth, t.rows, t.rows_padded, t.cols) == (o.depth, o.rows, o.rows_padded, o.cols)
        assert t.number_of_elements_without_padding == o.number_of_elements_without_padding


def test_host_layout_math_complex(oracle):
    """complex tensors are never vectorised (main.zig:141) and their type ids 10-19 lie past vector_widths[10]"""
    from wekua_b200 import Tensor, core

    dev = oracle.device("b200")
    for dtype in (np.complex64, np.complex128, core.Complex(np.int16)):
        for shape in [(5,), (4, 1), (5, 7), (2, 3, 4)]:
            t = Tensor.empty(_FakeCtx(), None, shape, dtype, _external_ptr=0x1000)
            o = oracle.OTensor(dev, dtype, shape).layout
            assert not t.vectors_enabled and t.vector_width == 1
            assert (t.row_pitch, t.slice_pitch, t.number_of_elements) == (o.row_pitch, o.slice_pitch, o.number_of_elements)


def test_gemm_validation_is_host_side():
    """validateTensors (gemm.zig:442-485) and PackedTensors.validateTensors (:250-270) run before any launch"""
    from wekua_b200 import Tensor, blas, capi

    ctx = _FakeCtx()
    a = Tensor.empty(ctx, None, (4, 5), np.float32, _external_ptr=0x1000)
    b = Tensor.empty(ctx, None, (6, 7), np.float32, _external_ptr=0x1000)
    c = Tensor.empty(ctx, None, (4, 7), np.float32, _external_ptr=0x1000)
    with pytest.raises(capi.InvalidValue):
        blas.gemm(None, None, a, 0, b, 0, None, c)
    with pytest.raises(capi.UnqualTensorsContext):
        blas.gemm(None, None, a, 0, Tensor.empty(_FakeCtx(), None, (5, 7), np.float32, _external_ptr=0x1000), 0, None, c)
    pt = blas.PackedTensors.init(None, c, 5, True)
    with pytest.raises(capi.InvalidValue):
        pt.pack(None, a, 0, b, 0)
    with pytest.raises(capi.UnqualTensorsShape):
        blas.axpy(None, a, 1, b)


def test_views_describe_only_the_memory_they_were_given():
    """Tensor.wrap (row / column blocks of a larger matrix): no pad row (it would be the next block's first row), metadata
    from the pitch the view really has, and whole-buffer ops refuse a view that does not own its flat span"""
    from wekua_b200 import Tensor, capi, math as wmath
    from wekua_b200.tensor import fill, memory

    ctx = _FakeCtx()
    rows_blk = Tensor.wrap(ctx, None, (5, 8), np.float32, 0x1000)               # a row block: odd rows, natural pitch
    assert rows_blk.is_view and rows_blk.owns_flat_span
    assert (rows_blk.rows_padded, rows_blk.slice_pitch, rows_blk.number_of_elements, rows_blk.size) == (5, 40, 40, 160)
    col_blk = Tensor.wrap(ctx, None, (5, 8), np.float32, 0x1000, row_pitch=32)  # a column block of a 32-wide matrix
    assert col_blk.is_view and not col_blk.owns_flat_span
    assert (col_blk.row_pitch, col_blk.slice_pitch, col_blk.number_of_elements, col_blk.pitches) == (32, 160, 160, [32, 1])
    for op in (lambda: wmath.sin(None, col_blk), lambda: memory.padded_to_numpy(None, col_blk)):
        with pytest.raises(capi.InvalidValue):
            op()
    owned = Tensor.empty(ctx, None, (5, 8), np.float32, _external_ptr=0x1000)
    assert not owned.is_view and owned.rows_padded == 6
