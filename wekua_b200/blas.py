"""Host-side mirror of src/blas: gemm (gemm.zig:834-874), PackedTensors (gemm.zig:60-359), axpy (axpy.zig:93-169),
plus scal / dot_reduce which the north star names but src/ lacks."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .core import Pipeline
from .tensor import Tensor, _np_ptr, _scalar, eql_tensors_shape


class Operation:  # gemm.zig:25-29
    no_transpose = 0
    transpose = 1


class PackedTensors:
    """GemmPackedTensors(T) (gemm.zig:60-359).  The reference re-tiles op(A) and op(B) into scratch tensors on
    every call; TMA reads the operands in place, so this stays an API-compatible handle that only remembers and
    validates the problem shape (validateTensors, gemm.zig:250-270) and owns no memory."""

    def __init__(self, n_size, m_size, k_size, vectors_enabled):
        self.n_size, self.m_size, self.k_size = int(n_size), int(m_size), int(k_size)
        self.vectors_enabled = False
        self.packed_a = self.packed_b = None

    @classmethod
    def init(cls, pipeline: Pipeline, result_tensor: Tensor, k_size: int, vectors_enabled: bool) -> "PackedTensors":
        if len(result_tensor.shape) != 2:
            raise capi.InvalidValue("InvalidValue")
        return cls(result_tensor.shape[0], result_tensor.shape[1], k_size, vectors_enabled)

    @classmethod
    def init_with_dimensions(cls, pipeline, n_size, m_size, k_size, recommended_algorithm=None, vectors_enabled=True):
        return cls(n_size, m_size, k_size, vectors_enabled)

    def validate(self, a: Tensor, op_a: int, b: Tensor, op_b: int) -> None:
        ok = (a.shape[1] == self.n_size and a.shape[0] == self.k_size) if op_a else \
             (a.shape[0] == self.n_size and a.shape[1] == self.k_size)
        ok &= (b.shape[1] == self.k_size and b.shape[0] == self.m_size) if op_b else \
              (b.shape[0] == self.k_size and b.shape[1] == self.m_size)
        if not ok:
            raise capi.InvalidValue("InvalidValue")

    def pack(self, pipeline, a, op_a, b, op_b) -> None:
        self.validate(a, op_a, b, op_b)

    def deinit(self, pipeline) -> None:
        pass


def _validate_gemm(a: Tensor, b: Tensor, c: Tensor, op_a: int, op_b: int):
    """validateTensors, gemm.zig:442-485"""
    if a.context is not b.context or a.context is not c.context:
        raise capi.UnqualTensorsContext("UnqualTensorsContext")
    if len(a.shape) != 2 or len(b.shape) != 2 or len(c.shape) != 2:
        raise capi.InvalidValue("InvalidValue")
    a_m, a_k = a.shape
    b_k, b_n = b.shape
    c_m, c_n = c.shape
    if op_a:
        match = (a_m == b_n and b_k == c_n and a_k == c_m) if op_b else (a_m == b_k and b_n == c_n and a_k == c_m)
    else:
        match = (a_k == b_n and b_k == c_n and a_m == c_m) if op_b else (a_k == b_k and b_n == c_n and a_m == c_m)
    if not match:
        raise capi.InvalidValue("InvalidValue")


def gemm(pipeline: Pipeline, alpha, a: Tensor, op_a: int, b: Tensor, op_b: int, beta, c: Tensor,
         packed_tensors: PackedTensors | None = None) -> None:
    """blas.gemm(T, pipeline, alpha, a, op_a, b, op_b, beta, c, packed) -- C = alpha*op(A)*op(B) + beta*C"""
    if a.dtype != b.dtype or a.dtype != c.dtype:
        raise capi.InvalidValue("InvalidValue: dtype mismatch")
    _validate_gemm(a, b, c, op_a, op_b)
    if packed_tensors is not None:
        packed_tensors.pack(pipeline, a, op_a, b, op_b)
    M, N = c.shape
    K = a.shape[1 - op_a]
    al, pal = _scalar(c.dtype, alpha)
    be, pbe = _scalar(c.dtype, beta)
    capi.check(capi.lib().wk_gemm(pipeline.q, c.type_index, op_a, op_b, M, N, K, pal, a.ptr, a.row_pitch, b.ptr,
                                  b.row_pitch, pbe, c.ptr, c.row_pitch))
    finish_c_padding(pipeline, c, beta)


def _pad_regions(c: Tensor):
    """(byte offset, rows, cols) of the padding of a 2-D tensor: the pad columns of every padded row, then the pad
    row's logical columns (vector width 1: at most one pad column and one pad row, main.zig:168,185-187)"""
    M, N = c.shape
    es = c.dtype.itemsize
    regions = []
    if c.row_pitch > N:
        regions.append((N * es, c.rows_padded, c.row_pitch - N))
    if c.rows_padded > M:
        regions.append((M * c.row_pitch * es, c.rows_padded - M, N))
    return regions


def finish_c_padding(pipeline: Pipeline, c: Tensor, beta, value=None) -> None:
    """The reference's GEMM kernels run over the whole PADDED C (global size rows_padded/2 x row_pitch/2,
    work_configuration.zig:166-186) and, on the packed path Linear and the benchmark use, the packed operands are
    zero outside the logical shape (gemm_pack.cl:89-91): C's padding becomes beta*padding, or 0 when beta is null.
    wk_gemm writes the logical M x N block only (C may be a row block of a larger matrix), so the host mirror
    finishes the padding of whole tensors here -- at most two tiny launches, and none for even shapes."""
    if not getattr(c, "_owns", True):
        return  # a view over someone else's memory has no padding of its own
    for off, rows, cols in _pad_regions(c):
        ptr = C.c_void_p(c.buffer + off)
        if beta is None or value is not None:
            v, pv = _scalar(c.dtype, 0 if value is None else value)
            capi.check(capi.lib().wk_fill(pipeline.q, c.type_index, 1, rows, cols, ptr, c.row_pitch, c.slice_pitch, pv))
        else:
            v, pv = _scalar(c.dtype, beta)
            capi.check(capi.lib().wk_scal(pipeline.q, c.type_index, 1, rows, cols, pv, ptr, c.row_pitch, c.slice_pitch))


def axpy(pipeline: Pipeline, x: Tensor, alpha, y: Tensor) -> None:
    """blas.axpy(T, pipeline, x, alpha, y) -- y += alpha*x over the logical region"""
    if x.dtype != y.dtype:
        raise capi.InvalidValue("InvalidValue: dtype mismatch")
    eql_tensors_shape(x, y)
    al, pal = _scalar(x.dtype, alpha)
    capi.check(capi.lib().wk_axpy(pipeline.q, x.type_index, x.depth, x.rows, x.cols, pal, x.ptr, x.row_pitch,
                                  x.slice_pitch, y.ptr, y.row_pitch, y.slice_pitch))


def scal(pipeline: Pipeline, alpha, x: Tensor) -> None:
    """x *= alpha (legacy blas scal, old_src/blas.c:41-67)"""
    al, pal = _scalar(x.dtype, alpha)
    capi.check(capi.lib().wk_scal(pipeline.q, x.type_index, x.depth, x.rows, x.cols, pal, x.ptr, x.row_pitch,
                                  x.slice_pitch))


def dot_reduce(pipeline: Pipeline, x: Tensor, y: Tensor):
    """BLAS-style reduction dot = sum(x*y) (blocking)"""
    eql_tensors_shape(x, y)
    out = np.zeros(1, dtype=x.dtype)
    capi.check(capi.lib().wk_dot_reduce(pipeline.q, x.type_index, x.depth, x.rows, x.cols, x.ptr, x.row_pitch,
                                        x.slice_pitch, y.ptr, y.row_pitch, y.slice_pitch, _np_ptr(out)))
    return out[0]
