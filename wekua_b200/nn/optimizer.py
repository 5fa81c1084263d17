"""src/nn/optimizers: Optimizer(T) vtable {step, zero, deinit} with GD (gd.zig), GDM (gdm.zig), Adagrad
(adagrad.zig), RMSProp (rmsprop.zig) and Adam (new: adam.zig is an empty file in the reference).

Deviation: the reference's history index never advances, so every layer shares history[0]; here each
parameter tensor owns its own state tensor, which is what the kernels' signatures intend."""
from __future__ import annotations

import numpy as np

from .. import blas, capi
from ..tensor import Tensor, _scalar, fill


def _params(cache):
    """(weight, gradient) then (bias, bias_gradient) pairs in the order gd.zig:55-94 walks them"""
    for slot in cache.slots:
        layer = slot.layer
        for w, g in zip(layer.get_weights(), layer.get_gradients(slot.cache)):
            yield w, g
        bgs = layer.get_bias_gradients(slot.cache)
        if bgs is not None:
            for b, bg in zip(layer.get_bias(), bgs):
                if bg is not None and b is not None:
                    yield b, bg


_KIND = {"gd": 0, "gdm": 1, "adagrad": 2, "rmsprop": 3, "adam": 4}


def _step_multi(pipeline, kind, records, lr, h0=None, h1=None, h2=None, t=0):
    """wk_optimizer_step_multi: one launch for the whole parameter list instead of the reference's one
    kernel per weight / bias tensor; bit-identical to the per-tensor calls.  records = [(x, g, state0, state1[, n])];
    n defaults to the whole padded buffer (what gdm.zig / adagrad.zig / rmsprop.zig launch over)."""
    if not records:
        return
    dtype = records[0][0].dtype
    arr = (capi.OptParam * len(records))()
    for i, rec in enumerate(records):
        x, g, s0, s1 = rec[:4]
        if x.dtype != dtype:
            raise capi.UnqualTensorsAttribute("UnqualTensorsAttribute: mixed dtypes in one optimizer step")
        arr[i] = capi.OptParam(x.ptr, g.ptr, s0.ptr if s0 is not None else None, s1.ptr if s1 is not None else None,
                               rec[4] if len(rec) > 4 else x.flat_elements("optimizer step"))
    keep = [_scalar(dtype, v) for v in (lr, h0, h1, h2)]
    capi.check(capi.lib().wk_optimizer_step_multi(pipeline.q, records[0][0].type_index, _KIND[kind], arr, len(records),
                                                  keep[0][1], keep[1][1], keep[2][1], keep[3][1], t))


class Optimizer:
    #: fused = True applies the update to every parameter tensor in ONE launch; False is the reference's per-tensor loop
    fused = True

    def step(self, pipeline, cache):
        raise NotImplementedError

    def zero(self, pipeline):
        for s in getattr(self, "_state", {}).values():
            for t in s:
                fill.zeroes(pipeline, t)

    def deinit(self, pipeline):
        for s in getattr(self, "_state", {}).values():
            for t in s:
                t.release(pipeline)
        self._state = {}

    def _state_for(self, pipeline, x: Tensor, n: int):
        """state tensors of parameter `x`, keyed by the parameter's device buffer and validated against its shape / dtype:
        a Python object id can be recycled by a new tensor after a layer is rebuilt, and would silently inherit stale
        momentum of possibly another shape"""
        key = (x.buffer, x.type_index)
        st = self._state.get(key)
        if st is not None and (st[0].shape != x.shape or st[0].buffer is None):
            for t in st:  # the buffer address was reused by a different parameter: start from fresh state
                t.release(pipeline)
            st = None
        if st is None:
            st = [Tensor.alloc(x.context, pipeline, x.shape, x.dtype) for _ in range(n)]
            self._state[key] = st
        return st


class GD(Optimizer):
    """GD.init(allocator, .{.lr}) -- step = axpy(g, -lr, w) (gd.zig:30-94); lr == 1 takes the SUBSTRACT kernel"""

    def __init__(self, lr):
        self.lr = -lr
        self._state = {}

    @classmethod
    def init(cls, allocator=None, lr=1.0):
        return cls(lr)

    def step(self, pipeline, cache):
        if self.fused:
            # axpy touches the LOGICAL region only (axpy.cl's non-vector variant; padding of a bias gradient is not zero):
            # tensors whose logical region is one contiguous prefix of the buffer go into the one-launch list, the
            # rest (padded rows) keep the pitched per-tensor kernel
            recs, rest = [], []
            for x, g in _params(cache):
                prefix = all(t.depth == 1 and (t.rows == 1 or t.row_pitch == t.cols) for t in (x, g))
                (recs if prefix else rest).append((x, g, None, None, x.rows * x.cols))
            _step_multi(pipeline, "gd", recs, self.lr)
            for x, g, *_ in rest:
                blas.axpy(pipeline, g, self.lr, x)
            return
        for x, g in _params(cache):
            blas.axpy(pipeline, g, self.lr, x)


class GDM(Optimizer):
    def __init__(self, lr, beta):
        self.lr, self.beta, self._state = lr, beta, {}

    def step(self, pipeline, cache):
        if self.fused:
            recs = [(x, g, self._state_for(pipeline, x, 1)[0], None) for x, g in _params(cache)]
            return _step_multi(pipeline, "gdm", recs, self.lr, self.beta)
        for x, g in _params(cache):
            (v,) = self._state_for(pipeline, x, 1)
            _a, plr = _scalar(x.dtype, self.lr)
            _b, pb = _scalar(x.dtype, self.beta)
            capi.check(capi.lib().wk_gdm(pipeline.q, x.type_index, x.ptr, g.ptr, v.ptr, plr, pb, x.flat_elements("optimizer step")))


class Adagrad(Optimizer):
    def __init__(self, lr):
        self.lr, self._state = lr, {}

    def step(self, pipeline, cache):
        if self.fused:
            recs = [(x, g, self._state_for(pipeline, x, 1)[0], None) for x, g in _params(cache)]
            return _step_multi(pipeline, "adagrad", recs, self.lr)
        for x, g in _params(cache):
            (h,) = self._state_for(pipeline, x, 1)
            _a, plr = _scalar(x.dtype, self.lr)
            capi.check(capi.lib().wk_adagrad(pipeline.q, x.type_index, x.ptr, g.ptr, h.ptr, plr, x.flat_elements("optimizer step")))


class RMSProp(Optimizer):
    def __init__(self, lr, gamma=0.9):
        self.lr, self.gamma, self._state = lr, gamma, {}

    def step(self, pipeline, cache):
        if self.fused:
            recs = [(x, g, self._state_for(pipeline, x, 1)[0], None) for x, g in _params(cache)]
            return _step_multi(pipeline, "rmsprop", recs, self.lr, self.gamma)
        for x, g in _params(cache):
            (h,) = self._state_for(pipeline, x, 1)
            _a, plr = _scalar(x.dtype, self.lr)
            _b, pg = _scalar(x.dtype, self.gamma)
            capi.check(capi.lib().wk_rmsprop(pipeline.q, x.type_index, x.ptr, g.ptr, h.ptr, plr, pg, x.flat_elements("optimizer step")))


class Adam(Optimizer):
    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.beta1, self.beta2, self.eps, self.t, self._state = lr, beta1, beta2, eps, 0, {}

    def step(self, pipeline, cache):
        self.t += 1
        if self.fused:
            recs = [(x, g, *self._state_for(pipeline, x, 2)) for x, g in _params(cache)]
            return _step_multi(pipeline, "adam", recs, self.lr, self.beta1, self.beta2, self.eps, self.t)
        for x, g in _params(cache):
            m, v = self._state_for(pipeline, x, 2)
            sc = [_scalar(x.dtype, s) for s in (self.lr, self.beta1, self.beta2, self.eps)]
            capi.check(capi.lib().wk_adam(pipeline.q, x.type_index, x.ptr, g.ptr, m.ptr, v.ptr, sc[0][1], sc[1][1],
                                          sc[2][1], sc[3][1], self.t, x.flat_elements("optimizer step")))
