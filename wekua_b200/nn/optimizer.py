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


class Optimizer:
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
        st = self._state.get(id(x))
        if st is None:
            st = [Tensor.alloc(x.context, pipeline, x.shape, x.dtype) for _ in range(n)]
            self._state[id(x)] = st
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
        for x, g in _params(cache):
            blas.axpy(pipeline, g, self.lr, x)


class GDM(Optimizer):
    def __init__(self, lr, beta):
        self.lr, self.beta, self._state = lr, beta, {}

    def step(self, pipeline, cache):
        for x, g in _params(cache):
            (v,) = self._state_for(pipeline, x, 1)
            _a, plr = _scalar(x.dtype, self.lr)
            _b, pb = _scalar(x.dtype, self.beta)
            capi.check(capi.lib().wk_gdm(pipeline.q, x.type_index, x.ptr, g.ptr, v.ptr, plr, pb, x.number_of_elements))


class Adagrad(Optimizer):
    def __init__(self, lr):
        self.lr, self._state = lr, {}

    def step(self, pipeline, cache):
        for x, g in _params(cache):
            (h,) = self._state_for(pipeline, x, 1)
            _a, plr = _scalar(x.dtype, self.lr)
            capi.check(capi.lib().wk_adagrad(pipeline.q, x.type_index, x.ptr, g.ptr, h.ptr, plr, x.number_of_elements))


class RMSProp(Optimizer):
    def __init__(self, lr, gamma=0.9):
        self.lr, self.gamma, self._state = lr, gamma, {}

    def step(self, pipeline, cache):
        for x, g in _params(cache):
            (h,) = self._state_for(pipeline, x, 1)
            _a, plr = _scalar(x.dtype, self.lr)
            _b, pg = _scalar(x.dtype, self.gamma)
            capi.check(capi.lib().wk_rmsprop(pipeline.q, x.type_index, x.ptr, g.ptr, h.ptr, plr, pg, x.number_of_elements))


class Adam(Optimizer):
    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.beta1, self.beta2, self.eps, self.t, self._state = lr, beta1, beta2, eps, 0, {}

    def step(self, pipeline, cache):
        self.t += 1
        for x, g in _params(cache):
            m, v = self._state_for(pipeline, x, 2)
            sc = [_scalar(x.dtype, s) for s in (self.lr, self.beta1, self.beta2, self.eps)]
            capi.check(capi.lib().wk_adam(pipeline.q, x.type_index, x.ptr, g.ptr, m.ptr, v.ptr, sc[0][1], sc[1][1],
                                          sc[2][1], sc[3][1], self.t, x.number_of_elements))
