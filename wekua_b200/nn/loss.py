"""src/nn/loss: mse (mse.zig:63-132 + mse.cl)."""
from __future__ import annotations

from .. import capi
from .. import math as wmath
from ..tensor import Tensor, eql_tensors


def mse(pipeline, output: Tensor, expected: Tensor, cache, calculate_derivative: bool = True, want_error: bool = False):
    """loss.mse(T, calc_dev, pipeline, output, expected, cache, ?*error_result).  Returns mean(error) when
    want_error (the reference's `error_result != null`), else None.  err = (t-o)^2, sensitivity = -2(t-o)."""
    err = cache.error_tensor
    eql_tensors(output, expected)
    eql_tensors(err, output)
    dev_ptr = None
    if calculate_derivative:
        last = cache.slots[-1]
        dev_ptr = last.layer.get_sensitivity(last.cache).ptr
    capi.check(capi.lib().wk_mse(pipeline.q, output.type_index, output.ptr, expected.ptr, err.ptr, dev_ptr,
                                 output.flat_elements("mse")))
    if not want_error:
        return None
    return wmath.mean(pipeline, err)
