"""Host-side mirror of src/math: dot (Hadamard, in place), sum, mean (basic.zig:17,131,206) and the six in-place
trigonometric / hyperbolic kernels (trig.zig:67-113)."""
from __future__ import annotations

import numpy as np

from . import capi
from .core import Pipeline
from .tensor import Tensor, _np_ptr, eql_tensors_shape

_OPS = {"sin": 0, "cos": 1, "tan": 2, "sinh": 3, "cosh": 4, "tanh": 5, "sigmoid": 6}


def dot(pipeline: Pipeline, x: Tensor, y: Tensor) -> None:
    """math.dot: x *= y element-wise (NOT a reduction), basic.zig:17-76"""
    eql_tensors_shape(x, y)
    capi.check(capi.lib().wk_hadamard(pipeline.q, x.type_index, x.depth, x.rows, x.cols, x.ptr, x.row_pitch, x.slice_pitch,
                                      y.ptr, y.row_pitch, y.slice_pitch))


def sum(pipeline: Pipeline, x: Tensor):  # noqa: A001 - the reference's name
    """math.sum, basic.zig:131-203: row sums over the PADDED row (sum.cl:33-35) then a total; blocking."""
    out = np.zeros(1, dtype=x.dtype)
    last_dim = x.shape[-1]
    if last_dim > 1:
        capi.check(capi.lib().wk_sum(pipeline.q, x.type_index, x.depth, x.rows, x.row_pitch, x.slice_pitch, x.ptr,
                                     _np_ptr(out)))
    else:
        # basic.zig:150-152,193-202: the tensor itself is mapped and its first row_length elements are added
        row_length = 1
        for s in x.shape[:-1]:
            row_length *= s
        capi.check(capi.lib().wk_sum(pipeline.q, x.type_index, 1, 1, row_length, row_length, x.ptr, _np_ptr(out)))
    return out[0]


def mean(pipeline: Pipeline, x: Tensor):
    """math.mean, basic.zig:206-240: @divTrunc for ints, / for floats, by the UNPADDED element count"""
    s = sum(pipeline, x)
    n = x.number_of_elements_without_padding
    if x.dtype.names is not None:  # complex: each component on its own (basic.zig:216-227)
        out = np.zeros((), dtype=x.dtype)
        for comp in ("re", "im"):
            out[comp] = _mean_component(s[comp], n)
        return out[()]
    return _mean_component(s, n)


def _mean_component(s, n):
    dt = np.dtype(type(s))
    if dt.kind == "f":
        return dt.type(s / dt.type(n))
    si, ni = int(s), int(n)
    q = abs(si) // ni
    return dt.type(q if si >= 0 else -q)


def _unary(name):
    def f(pipeline: Pipeline, tensor: Tensor) -> None:
        # trig.zig:45-51: 1-D over the whole padded buffer
        n = tensor.flat_elements(name)
        capi.check(capi.lib().wk_unary(pipeline.q, tensor.type_index, _OPS[name], tensor.ptr, n))
    f.__name__ = name
    f.__doc__ = f"math.{name}: in place, trig.zig:67-113"
    return f


sin, cos, tan, sinh, cosh, tanh = (_unary(n) for n in ("sin", "cos", "tan", "sinh", "cosh", "tanh"))
