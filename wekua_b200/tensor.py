"""Host-side mirror of src/tensor: Tensor(T) with the reference's padded/pitched layout (main.zig:62-279),
memory.{readFromBuffer, writeToBuffer, copy, getValue, putValue}, fill.{constant, one, zeroes}, identity,
transpose and random.uniform.  Only layout math lives here; every byte moves through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import capi
from .core import Context, Pipeline, as_elements, base_type, get_type_index, storage_dtype


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _scalar(dtype, v):
    """pointer to one host element of the tensor's type, or None for Zig `null`"""
    if v is None:
        return None, None
    a = as_elements(v, dtype).reshape(1)
    return a, _np_ptr(a)


class Tensor:
    """Tensor(T).empty / alloc / release (main.zig:113-279)."""

    def __init__(self):
        raise TypeError("use Tensor.empty / Tensor.alloc")

    @classmethod
    def empty(cls, context: Context, pipeline: Pipeline, shape, dtype=np.float32, vectors_enabled: bool = True,
              _external_ptr: int | None = None) -> "Tensor":
        self = object.__new__(cls)
        shape = [int(s) for s in shape]
        if len(shape) == 0 or any(s == 0 for s in shape):
            raise capi.InvalidValue("InvalidValue: empty shape / zero dimension")
        self.context = context
        self.dtype = storage_dtype(dtype)
        self.type_index = get_type_index(dtype)
        self.shape = tuple(shape)
        # main.zig:142-150: vector width = max over the context's queues (all 1 on the CUDA backend)
        vw = 1
        vectors_enabled = vectors_enabled and self.type_index < 10  # main.zig:141: complex tensors are never vectorised
        if vectors_enabled:
            for cq in context.command_queues:
                vw = max(vw, int(cq.vector_widths[self.type_index]))
            vectors_enabled = vw > 1
        self.vectors_enabled = vectors_enabled
        self.vector_width = vw
        ndim = len(shape)
        last = ndim - 1
        pen = max(last - 1, 0)
        depth = 1
        for e in shape[:pen]:
            depth *= e
        pen_size = shape[pen] if ndim >= 2 else 1
        last_size = shape[last]
        padded_pen = pen_size + (pen_size % 2)                      # :168
        self.depth, self.rows, self.rows_padded, self.cols = depth, pen_size, padded_pen, last_size
        self.number_of_elements_without_padding = depth * last_size * pen_size
        row_pitch = last_size
        if vectors_enabled and vw > 1 and row_pitch % vw:
            row_pitch += vw - row_pitch % vw                        # :174-180
        rpv = row_pitch // vw
        self.vl_shape = list(shape)
        self.vl_shape[last] = rpv
        rem = rpv % 2
        rpv += rem
        row_pitch += vw * rem                                       # :185-187
        self.row_pitch, self.row_pitch_for_vectors = row_pitch, rpv
        self.slice_pitch = row_pitch * padded_pen
        self.number_of_elements = self.slice_pitch * depth
        self.slice_pitch_for_vectors = self.slice_pitch // vw
        self.number_of_vectors = self.number_of_elements // vw
        pitches = [0] * ndim
        ante = max(pen - 1, 0)
        pitch = self.number_of_elements
        for i in range(ante):
            pitch //= shape[i]
            pitches[i] = pitch
        if ndim >= 3:
            pitches[ante] = self.slice_pitch
        if ndim >= 2:
            pitches[pen] = row_pitch
        pitches[last] = 1
        self.pitches = pitches
        self.size = self.number_of_elements * self.dtype.itemsize
        self._owns = _external_ptr is None
        self.is_view = False
        if _external_ptr is None:
            p = C.c_void_p()
            capi.check(capi.lib().wk_malloc(pipeline.q, self.size, C.byref(p)))
            self.buffer = p.value
        else:
            self.buffer = int(_external_ptr)
        return self

    @classmethod
    def alloc(cls, context, pipeline, shape, dtype=np.float32, vectors_enabled=True) -> "Tensor":
        t = cls.empty(context, pipeline, shape, dtype, vectors_enabled)
        fill.zeroes(pipeline, t)
        return t

    @classmethod
    def wrap(cls, context, pipeline, shape, dtype, device_ptr: int, row_pitch: int | None = None) -> "Tensor":
        """view over device memory someone else owns (a row / column block of a larger tensor) with this layout;
        `row_pitch` (elements) overrides the pitch the shape alone would give, for column blocks of a wider matrix"""
        t = cls.empty(context, pipeline, shape, dtype, True, _external_ptr=device_ptr)
        natural_pitch = t.row_pitch
        if row_pitch is not None:
            if len(t.shape) != 2 or row_pitch < t.shape[1]:
                raise capi.InvalidValue("InvalidValue: row_pitch override needs a 2-D view and pitch >= cols")
            t.row_pitch = int(row_pitch)
        # A view describes exactly the memory it was given: it owns no pad row (for an odd-row block of a larger matrix
        # that row is the NEXT block's first row) and its metadata follows the pitch it really has.
        t.is_view = True
        t.rows_padded = t.rows
        t.row_pitch_for_vectors = t.row_pitch // t.vector_width
        t.slice_pitch = t.row_pitch * t.rows_padded
        t.slice_pitch_for_vectors = t.slice_pitch // t.vector_width
        t.number_of_elements = t.slice_pitch * t.depth
        t.number_of_vectors = t.number_of_elements // t.vector_width
        t.size = t.number_of_elements * t.dtype.itemsize
        nd = len(t.shape)
        if nd >= 2:
            t.pitches[nd - 2] = t.row_pitch
        if nd >= 3:
            t.pitches[nd - 3] = t.slice_pitch
            for i in range(nd - 4, -1, -1):
                t.pitches[i] = t.pitches[i + 1] * t.shape[i + 1]
        # the flat span [buffer, buffer + size) belongs to the view only when its rows are as wide as its pitch allows
        # (a row block); a column block's span runs through its neighbours' columns
        t.owns_flat_span = t.row_pitch == natural_pitch
        return t

    def flat_elements(self, what: str) -> int:
        """number_of_elements for an op that streams over the whole padded buffer (checks that the tensor owns it)"""
        self.require_flat_span(what)
        return self.number_of_elements

    def require_flat_span(self, what: str) -> None:
        """ops that stream over the whole padded buffer (the reference's 1-D launches) are only defined for a
        tensor that owns that span: column-block views must use the pitched (logical-region) ops"""
        if self.is_view and not self.owns_flat_span:
            raise capi.InvalidValue(f"InvalidValue: {what} runs over the whole buffer; this view (pitch {self.row_pitch}, "
                                    f"{self.cols} columns) does not own it")

    def release(self, pipeline: Pipeline) -> None:
        if self.buffer is not None and self._owns:
            capi.check(capi.lib().wk_free(pipeline.q, C.c_void_p(self.buffer)))  # syncs first, like main.zig:254
        self.buffer = None

    @property
    def ptr(self):
        return C.c_void_p(self.buffer)

    def pitch_sum(self) -> int:
        return sum(self.pitches)


def eql_tensors_shape(a: Tensor, b: Tensor) -> None:
    """tensor/helpers.zig:53-57"""
    if a.shape != b.shape:
        raise capi.UnqualTensorsShape("UnqualTensorsShape")


def eql_tensors(a: Tensor, b: Tensor) -> None:
    """tensor/helpers.zig eqlTensors: shape + layout attributes"""
    eql_tensors_shape(a, b)
    if a.vectors_enabled != b.vectors_enabled or a.number_of_elements != b.number_of_elements:
        raise capi.UnqualTensorsAttribute("UnqualTensorsAttribute")


class memory:
    @staticmethod
    def read_from_buffer(pipeline: Pipeline, tensor: Tensor, buffer) -> None:
        """memory.readFromBuffer (host -> tensor), read_from_buffer.zig:13-63.  The host array must stay alive
        until the pipeline is waited on (same lifetime rule as the reference's non-blocking writeRect)."""
        host = as_elements(buffer, tensor.dtype).reshape(-1)
        if host.size != tensor.number_of_elements_without_padding:
            raise capi.InvalidBuffer("InvalidBuffer")
        es = tensor.dtype.itemsize
        tensor._keepalive = host
        capi.check(capi.lib().wk_h2d_rect(pipeline.q, tensor.ptr, tensor.row_pitch * es, tensor.slice_pitch * es,
                                          _np_ptr(host), tensor.cols * es, tensor.rows, tensor.depth))

    @staticmethod
    def write_to_buffer(pipeline: Pipeline, tensor: Tensor, buffer: np.ndarray) -> None:
        """memory.writeToBuffer (tensor -> host), write_to_buffer.zig:13-63; valid after wait_and_cleanup()"""
        if buffer.dtype != tensor.dtype or buffer.size != tensor.number_of_elements_without_padding \
                or not buffer.flags["C_CONTIGUOUS"]:
            raise capi.InvalidBuffer("InvalidBuffer")
        es = tensor.dtype.itemsize
        capi.check(capi.lib().wk_d2h_rect(pipeline.q, _np_ptr(buffer), tensor.ptr, tensor.row_pitch * es,
                                          tensor.slice_pitch * es, tensor.cols * es, tensor.rows, tensor.depth))

    @staticmethod
    def to_numpy(pipeline: Pipeline, tensor: Tensor) -> np.ndarray:
        out = np.empty(tensor.number_of_elements_without_padding, dtype=tensor.dtype)
        memory.write_to_buffer(pipeline, tensor, out)
        pipeline.wait_and_cleanup()
        return out.reshape(tensor.shape)

    @staticmethod
    def padded_to_numpy(pipeline: Pipeline, tensor: Tensor) -> np.ndarray:
        """the whole padded buffer (tests compare padding behaviour with the oracle)"""
        tensor.require_flat_span("padded_to_numpy")
        out = np.empty(tensor.number_of_elements, dtype=tensor.dtype)
        capi.check(capi.lib().wk_d2h_rect(pipeline.q, _np_ptr(out), tensor.ptr, tensor.size, tensor.size, tensor.size, 1, 1))
        pipeline.wait_and_cleanup()
        return out

    @staticmethod
    def copy(pipeline: Pipeline, src: Tensor, dst: Tensor) -> None:
        """memory.copy, copy.zig:85-98"""
        eql_tensors_shape(src, dst)
        es = src.dtype.itemsize
        flat = not (src.is_view and not src.owns_flat_span) and not (dst.is_view and not dst.owns_flat_span)
        if flat and src.row_pitch == dst.row_pitch and src.slice_pitch == dst.slice_pitch and src.size == dst.size:
            capi.check(capi.lib().wk_d2d(pipeline.q, dst.ptr, src.ptr, src.size))
        else:
            capi.check(capi.lib().wk_d2d_rect(pipeline.q, dst.ptr, dst.row_pitch * es, dst.slice_pitch * es, src.ptr,
                                              src.row_pitch * es, src.slice_pitch * es, src.cols * es, src.rows, src.depth))

    @staticmethod
    def _offset(tensor: Tensor, coords) -> int:
        if len(coords) != len(tensor.shape) or any(c >= s for c, s in zip(coords, tensor.shape)):
            raise capi.InvalidCoordinates("InvalidCoordinates")
        return sum(c * p for c, p in zip(coords, tensor.pitches)) * tensor.dtype.itemsize

    @staticmethod
    def put_value(pipeline: Pipeline, tensor: Tensor, coords, value) -> None:
        a, p = _scalar(tensor.dtype, value)
        capi.check(capi.lib().wk_put_value(pipeline.q, tensor.ptr, memory._offset(tensor, coords), p, tensor.dtype.itemsize))

    @staticmethod
    def get_value(pipeline: Pipeline, tensor: Tensor, coords):
        out = np.zeros(1, dtype=tensor.dtype)
        capi.check(capi.lib().wk_get_value(pipeline.q, tensor.ptr, memory._offset(tensor, coords), _np_ptr(out),
                                           tensor.dtype.itemsize))
        return out[0]


class fill:
    @staticmethod
    def constant(pipeline: Pipeline, tensor: Tensor, value) -> None:
        """fill.constant (fill.zig:15-70): logical region only"""
        a, p = _scalar(tensor.dtype, value)
        capi.check(capi.lib().wk_fill(pipeline.q, tensor.type_index, tensor.depth, tensor.rows, tensor.cols, tensor.ptr,
                                      tensor.row_pitch, tensor.slice_pitch, p))

    @staticmethod
    def one(pipeline: Pipeline, tensor: Tensor) -> None:
        fill.constant(pipeline, tensor, 1)

    @staticmethod
    def zeroes(pipeline: Pipeline, tensor: Tensor) -> None:
        """fill.zeroes (fill.zig:72-95): the whole padded buffer (a column-block view: its logical region)"""
        if tensor.is_view and not tensor.owns_flat_span:
            return fill.constant(pipeline, tensor, 0)
        capi.check(capi.lib().wk_memset_zero(pipeline.q, tensor.ptr, tensor.size))


def identity(pipeline: Pipeline, tensor: Tensor) -> None:
    """identity.zig:16-70"""
    size = tensor.shape[0]
    if any(s != size for s in tensor.shape[1:]):
        raise capi.InvalidValue("InvalidValue")
    capi.check(capi.lib().wk_identity(pipeline.q, tensor.type_index, tensor.ptr, tensor.number_of_elements, size,
                                      tensor.pitch_sum()))


def transpose(pipeline: Pipeline, result_tensor: Tensor, tensor: Tensor, dim0: int, dim1: int) -> None:
    """transpose.zig:15-113: swap two dimensions (any rank; 2-D takes the shared-memory tiled kernel)"""
    if len(result_tensor.shape) != len(tensor.shape):
        raise capi.UnqualTensorsDimension("UnqualTensorsDimension")
    nd = len(tensor.shape)
    if dim0 >= nd or dim1 >= nd:
        raise capi.InvalidValue("InvalidValue")
    if tensor.number_of_elements_without_padding != result_tensor.number_of_elements_without_padding:
        raise capi.UnqualTensorsDimension("UnqualTensorsDimension")
    if result_tensor.shape[dim0] != tensor.shape[dim1] or result_tensor.shape[dim1] != tensor.shape[dim0]:
        raise capi.InvalidValue("InvalidValue")
    if dim0 == dim1:
        memory.copy(pipeline, tensor, result_tensor)
        return
    if nd != 2:
        d0, d1 = min(dim0, dim1), max(dim0, dim1)
        pa = (C.c_uint64 * nd)(*tensor.pitches)
        pb = (C.c_uint64 * nd)(*result_tensor.pitches)
        capi.check(capi.lib().wk_transpose_nd(pipeline.q, tensor.type_index, nd, tensor.ptr, pa, result_tensor.ptr, pb,
                                              tensor.row_pitch, tensor.slice_pitch, tensor.rows * tensor.row_pitch, tensor.cols,
                                              tensor.number_of_elements, d0, d1))
        return
    capi.check(capi.lib().wk_transpose2d(pipeline.q, tensor.type_index, tensor.rows, tensor.cols, tensor.ptr,
                                         tensor.row_pitch, result_tensor.ptr, result_tensor.row_pitch))


class random:
    @staticmethod
    def uniform(pipeline: Pipeline, tensor: Tensor, seed=None, min_value=None, max_value=None) -> None:
        """random.uniform (uniform.zig:60-123); seed None = wall clock (uniform.zig:82)"""
        if seed is None:
            seed = int(time.time())
        a, pa = _scalar(base_type(tensor.dtype), min_value)  # bounds are scalars of the base type (uniform.zig:64-65)
        b, pb = _scalar(base_type(tensor.dtype), max_value)
        capi.check(capi.lib().wk_uniform(pipeline.q, tensor.type_index, tensor.depth, tensor.rows, tensor.cols, tensor.ptr,
                                         tensor.row_pitch, tensor.slice_pitch, C.c_uint64(seed & (2**64 - 1)), pa, pb))
