"""src/nn/layer: Layer(T) interface (main.zig:18-56), Linear (linear.zig), Sequential (sequential.zig), Cache
(cache.zig).  Same call sequence as the reference; `fused=True` swaps the per-op launches of Linear.forward /
backward for the fused kernels (GEMM epilogue with bias + activation; act' * sensitivity in one pass)."""
from __future__ import annotations

import ctypes as C
import math as _math

import numpy as np

from .. import blas, capi
from .. import math as wmath
from ..core import Context, Pipeline
from ..tensor import Tensor, _scalar, fill, random
from .activation import ACT_NONE, Activation


class LinearCache:
    """linear.zig:57-71"""

    def __init__(self):
        self.outputs, self.sensitivities, self.acti_derivatives = [], [], []
        self.gradients, self.bias_gradients = [], []
        self.forward_packed, self.grad_packed, self.sensitivity_packed = [], [], []


def _random_limits(dtype, n_in, n_out):
    """getRandomLimits, linear.zig:28-45 (Xavier-uniform)"""
    if np.dtype(dtype).kind == "f":
        limit = _math.sqrt(6.0 / float(n_in + n_out))
        return -limit, limit
    limit = int(_math.isqrt(6 // (n_in + n_out)))
    return (-limit if np.dtype(dtype).kind == "i" else 0), limit


class Linear:
    """Linear(T).init(context, pipeline, input, output, activation, ExtraParams{deep, enable_bias}) linear.zig:88-178"""

    def __init__(self, context: Context, pipeline: Pipeline, n_input: int, n_output: int, acti: Activation | None = None,
                 deep: int = 1, enable_bias: bool = True, dtype=np.float32, seed=None, fused: bool = True):
        if n_input == 0 or n_output == 0 or deep == 0:
            raise capi.InvalidValue("InvalidValue")
        self.context, self.activation, self.bias_enabled, self.dtype = context, acti, enable_bias, np.dtype(dtype)
        # the fused epilogue / backward kernels exist for f32 and f64; every other dtype takes the reference's op-by-op
        # sequence (gemm, bias.cl, bias_step.cl are generic over the real dtypes)
        self.fused = bool(fused) and self.dtype in (np.dtype(np.float32), np.dtype(np.float64))
        self.weights, self.bias = [], []
        lo, hi = _random_limits(dtype, n_input, n_output)
        w = Tensor.alloc(context, pipeline, (n_output, n_input), dtype)
        random.uniform(pipeline, w, seed, lo, hi)  # the reference seeds from the wall clock (Q6); tests pass a seed
        self.weights.append(w)
        lo, hi = _random_limits(dtype, n_output, n_output)
        for i in range(1, deep):
            w = Tensor.alloc(context, pipeline, (n_output, n_output), dtype)
            random.uniform(pipeline, w, None if seed is None else seed + i, lo, hi)
            self.weights.append(w)
        if enable_bias:
            self.bias = [Tensor.alloc(context, pipeline, (n_output,), dtype) for _ in range(deep)]

    @classmethod
    def init(cls, context, pipeline, n_input, n_output, acti=None, extra_params=None, **kw):
        extra_params = extra_params or {}
        return cls(context, pipeline, n_input, n_output, acti, extra_params.get("deep", 1),
                   extra_params.get("enable_bias", True), **kw)

    # ---- Layer vtable (layer/main.zig:18-56)
    def deinit(self, pipeline):
        for w in self.weights:
            w.release(pipeline)
        for b in self.bias:
            b.release(pipeline)

    def get_cached_output(self, cache: LinearCache) -> Tensor:
        return cache.outputs[-1]

    def get_weights(self):
        return self.weights

    def get_bias(self):
        return self.bias if self.bias_enabled else None

    def prepare_cache(self, pipeline: Pipeline, number_of_elements: int) -> LinearCache:
        """linear.zig:219-378"""
        c = LinearCache()
        ctx, dt = self.context, self.dtype
        for w in self.weights:
            out = w.shape[0]
            c.outputs.append(Tensor.alloc(ctx, pipeline, (number_of_elements, out), dt))
            s = Tensor.alloc(ctx, pipeline, (number_of_elements, out), dt)
            fill.constant(pipeline, s, 1)
            c.sensitivities.append(s)
            c.acti_derivatives.append(Tensor.alloc(ctx, pipeline, (number_of_elements, out), dt))
            c.gradients.append(Tensor.alloc(ctx, pipeline, w.shape, dt))
            c.bias_gradients.append(Tensor.alloc(ctx, pipeline, (out,), dt))
        for w, o, g in zip(self.weights, c.outputs, c.gradients):
            c.forward_packed.append(blas.PackedTensors.init(pipeline, o, w.shape[1], True))
            c.grad_packed.append(blas.PackedTensors.init(pipeline, g, number_of_elements, True))
        for i, w in enumerate(self.weights):
            if i > 0:
                c.sensitivity_packed.append(blas.PackedTensors.init(pipeline, c.sensitivities[i - 1], w.shape[0], True))
            else:
                c.sensitivity_packed.append(blas.PackedTensors.init_with_dimensions(
                    pipeline, number_of_elements, self.weights[0].shape[1], self.weights[0].shape[0]))
        return c

    def release_cache(self, pipeline, cache: LinearCache):
        for lst in (cache.outputs, cache.gradients, cache.bias_gradients, cache.acti_derivatives, cache.sensitivities):
            for t in lst:
                t.release(pipeline)

    @staticmethod
    def _add_bias(pipeline, output: Tensor, bias: Tensor):
        """addBias, linear.zig:424-478"""
        capi.check(capi.lib().wk_bias_add(pipeline.q, output.type_index, output.ptr, bias.ptr, output.row_pitch,
                                          output.flat_elements("bias")))

    def forward(self, pipeline: Pipeline, input_tensor: Tensor, cache: LinearCache) -> Tensor:
        """linear.zig:480-525: out = act(in . W^T + b) per sub-layer"""
        inp = input_tensor
        for idx, (weight, output, fwd_pt) in enumerate(zip(self.weights, cache.outputs, cache.forward_packed)):
            if self.fused:
                if inp.dtype != weight.dtype or output.dtype != weight.dtype:  # the check blas.gemm makes
                    raise capi.UnqualTensorsAttribute("UnqualTensorsAttribute: dtype mismatch")
                fwd_pt.pack(pipeline, inp, 0, weight, 1)
                M, N = output.shape
                capi.check(capi.lib().wk_gemm_bias_act(
                    pipeline.q, output.type_index, 0, 1, M, N, inp.shape[1], inp.ptr, inp.row_pitch, weight.ptr,
                    weight.row_pitch, output.ptr, output.row_pitch, self.bias[idx].ptr if self.bias_enabled else None,
                    self.activation.kind if self.activation else ACT_NONE))
                self._fused_padding(pipeline, output, idx)
            else:
                blas.gemm(pipeline, None, inp, blas.Operation.no_transpose, weight, blas.Operation.transpose, None, output, fwd_pt)
                if self.bias_enabled:
                    self._add_bias(pipeline, output, self.bias[idx])
                if self.activation is not None:
                    self.activation.run(pipeline, output)
            inp = output
        return inp

    def _fused_padding(self, pipeline, output: Tensor, idx: int):
        """what the unfused sequence leaves in the padding of `output` (odd shapes only): the packed GEMM writes 0
        there (gemm_pack.cl:89-91), bias.cl adds bias[i % row_pitch] over the whole buffer (bias.cl:3-19) and the
        activation runs over the whole buffer (sigmoid.zig:62-80) -- reproduced so math.sum / mse see the same values"""
        M, N = output.shape
        if output.row_pitch == N and output.rows_padded == M:
            return
        blas.finish_c_padding(pipeline, output, None)
        lib, q, ti, es = capi.lib(), pipeline.q, output.type_index, output.dtype.itemsize
        act_op = {1: 6, 2: 5}.get(self.activation.kind) if self.activation is not None else None
        if output.rows_padded > M:  # the pad row sees the logical bias values: act(0 + bias[j])
            ptr = C.c_void_p(output.buffer + M * output.row_pitch * es)
            n = (output.rows_padded - M) * output.row_pitch
            if self.bias_enabled:
                capi.check(lib.wk_bias_add(q, ti, ptr, self.bias[idx].ptr, output.row_pitch, n))
            if act_op is not None:
                capi.check(lib.wk_unary(q, ti, act_op, ptr, n))
        if output.row_pitch > N and act_op is not None:  # pad column of the logical rows: act(0 + 0)
            at_zero = 0.5 if self.activation.kind == 1 else 0.0
            _, pv = _scalar(output.dtype, at_zero)
            capi.check(lib.wk_fill(q, ti, 1, M, output.row_pitch - N, C.c_void_p(output.buffer + N * es),
                                   output.row_pitch, output.slice_pitch, pv))

    def get_sensitivity(self, cache: LinearCache) -> Tensor:
        return cache.sensitivities[-1]

    @staticmethod
    def _bias_sensitivity(pipeline, sensitivity: Tensor, bias_gradient: Tensor):
        """getBiasSensitivity, linear.zig:534-577"""
        capi.check(capi.lib().wk_bias_step(pipeline.q, sensitivity.type_index, sensitivity.ptr, bias_gradient.ptr,
                                           sensitivity.row_pitch_for_vectors, sensitivity.shape[0],
                                           bias_gradient.row_pitch_for_vectors))

    def backward(self, pipeline: Pipeline, cache: LinearCache, input_tensor: Tensor, input_sensitivity: Tensor | None):
        """linear.zig:579-678"""
        sens = cache.sensitivities[-1]
        index = len(self.weights) - 1
        output = cache.outputs[index]
        while True:
            acti_derivative = cache.acti_derivatives[index]
            prev_output = cache.outputs[index - 1] if index >= 1 else input_tensor
            if index >= 1:
                next_sens = cache.sensitivities[index - 1]
            else:
                next_sens = input_sensitivity
            if self.fused:
                # one call: act' o sensitivity formed inside the two GEMMs (f32, tensor-core path: 2 launches), the bias
                # gradient summed inside the first; otherwise the library runs the op-by-op sequence itself
                grad = cache.gradients[index]
                bgrad = cache.bias_gradients[index] if self.bias_enabled else None
                capi.check(capi.lib().wk_linear_backward(
                    pipeline.q, sens.type_index, self.activation.kind if self.activation is not None else ACT_NONE,
                    sens.shape[0], sens.shape[1], self.weights[index].shape[1], sens.ptr, sens.row_pitch, output.ptr,
                    output.row_pitch, prev_output.ptr, prev_output.row_pitch, self.weights[index].ptr,
                    self.weights[index].row_pitch, grad.ptr, grad.row_pitch, bgrad.ptr if bgrad is not None else None,
                    next_sens.ptr if next_sens is not None else None, next_sens.row_pitch if next_sens is not None else 0))
                blas.finish_c_padding(pipeline, grad, None)
                if next_sens is not None:
                    blas.finish_c_padding(pipeline, next_sens, None)
            else:
                if self.activation is not None:
                    self.activation.get_derivative(pipeline, output, acti_derivative)
                wmath.dot(pipeline, sens, acti_derivative)
                blas.gemm(pipeline, None, sens, blas.Operation.transpose, prev_output, blas.Operation.no_transpose, None,
                          cache.gradients[index], cache.grad_packed[index])
                if self.bias_enabled:
                    self._bias_sensitivity(pipeline, sens, cache.bias_gradients[index])
                if next_sens is not None:
                    blas.gemm(pipeline, None, sens, blas.Operation.no_transpose, self.weights[index], blas.Operation.no_transpose,
                              None, next_sens, cache.sensitivity_packed[index])
            if next_sens is None or index == 0:
                break
            output = prev_output
            index -= 1
            sens = next_sens

    def get_gradients(self, cache: LinearCache):
        return cache.gradients

    def get_bias_gradients(self, cache: LinearCache):
        return cache.bias_gradients if self.bias_enabled else None


class SequentialCache:
    def __init__(self, caches, gradients, bias_gradients):
        self.caches, self.gradients, self.bias_gradients = caches, gradients, bias_gradients


class Sequential:
    """sequential.zig: an ordered container of layers exposing the same Layer interface"""

    def __init__(self):
        self.layers = []

    @classmethod
    def init(cls, allocator=None):
        return cls()

    def append(self, layer) -> None:
        self.layers.append(layer)

    def layer(self) -> "Sequential":
        return self

    def deinit(self, pipeline):
        for l in self.layers:
            l.deinit(pipeline)

    def get_weights(self):
        return [w for l in self.layers for w in l.get_weights()]

    def get_bias(self):
        out = []
        for l in self.layers:
            b = l.get_bias()
            out.extend(b if b is not None else [None] * len(l.get_weights()))
        return out

    def prepare_cache(self, pipeline, number_of_elements) -> SequentialCache:
        caches = [l.prepare_cache(pipeline, number_of_elements) for l in self.layers]
        grads = [g for l, c in zip(self.layers, caches) for g in l.get_gradients(c)]
        bgrads = []
        for l, c in zip(self.layers, caches):
            bg = l.get_bias_gradients(c)
            bgrads.extend(bg if bg is not None else [None] * len(l.get_weights()))
        return SequentialCache(caches, grads, bgrads)

    def release_cache(self, pipeline, cache: SequentialCache):
        for l, c in zip(self.layers, cache.caches):
            l.release_cache(pipeline, c)

    def get_cached_output(self, cache: SequentialCache):
        return self.layers[-1].get_cached_output(cache.caches[-1])

    def forward(self, pipeline, input_tensor, cache: SequentialCache):
        """sequential.zig:211-228"""
        out = input_tensor
        for l, c in zip(self.layers, cache.caches):
            out = l.forward(pipeline, out, c)
        return out

    def get_sensitivity(self, cache: SequentialCache):
        return self.layers[-1].get_sensitivity(cache.caches[-1])

    def backward(self, pipeline, cache: SequentialCache, input_tensor, input_gradient):
        """sequential.zig:242-274"""
        for index in range(len(self.layers) - 1, -1, -1):
            _input = input_tensor if index == 0 else self.layers[index - 1].get_cached_output(cache.caches[index - 1])
            _grad = input_gradient if index == 0 else self.layers[index - 1].get_sensitivity(cache.caches[index - 1])
            self.layers[index].backward(pipeline, cache.caches[index], _input, _grad)

    def get_gradients(self, cache: SequentialCache):
        return cache.gradients

    def get_bias_gradients(self, cache: SequentialCache):
        return cache.bias_gradients


class CacheSlot:
    def __init__(self, cache, layer):
        self.cache, self.layer = cache, layer


class Cache:
    """Cache(T).init(context, pipeline, number_of_elements, layers) cache.zig:19-52"""

    def __init__(self, context, pipeline, number_of_elements, layers):
        self.slots = [CacheSlot(l.prepare_cache(pipeline, number_of_elements), l) for l in layers]
        last = self.slots[-1]
        last_sens = last.layer.get_sensitivity(last.cache)
        self.error_tensor = Tensor.alloc(context, pipeline, last_sens.shape, last_sens.dtype)

    @classmethod
    def init(cls, context, pipeline, number_of_elements, layers):
        return cls(context, pipeline, number_of_elements, layers)

    def get_layer_cache(self, index):
        return self.slots[index].cache

    def deinit(self, pipeline):
        for s in self.slots:
            s.layer.release_cache(pipeline, s.cache)
        self.error_tensor.release(pipeline)
