"""src/nn/activation: Activation(T) vtable {run, getDerivative} (main.zig), Sigmoid (sigmoid.zig:39-137),
Tanh (tanh.zig:25-84).  f32/f64 only (sigmoid.zig:21-24)."""
from __future__ import annotations

from .. import capi
from ..core import Pipeline
from ..tensor import Tensor, eql_tensors

ACT_NONE, ACT_SIGMOID, ACT_TANH = 0, 1, 2


class Activation:
    kind = ACT_NONE

    def run(self, pipeline: Pipeline, net_output: Tensor) -> None:
        raise NotImplementedError

    def get_derivative(self, pipeline: Pipeline, output: Tensor, derivative: Tensor) -> None:
        eql_tensors(output, derivative)  # activation/main.zig getDerivative
        self._derivative(pipeline, output, derivative)


class Sigmoid(Activation):
    kind = ACT_SIGMOID

    @classmethod
    def init(cls):
        return cls()

    def run(self, pipeline, net_output):
        # sigmoid.zig:62-80: 1-D over the whole padded buffer
        capi.check(capi.lib().wk_unary(pipeline.q, net_output.type_index, 6, net_output.ptr, net_output.flat_elements("activation")))

    def _derivative(self, pipeline, output, derivative):
        capi.check(capi.lib().wk_sigmoid_dev(pipeline.q, output.type_index, output.ptr, derivative.ptr,
                                             output.flat_elements("activation derivative")))


class Tanh(Activation):
    kind = ACT_TANH

    @classmethod
    def init(cls):
        return cls()

    def run(self, pipeline, net_output):
        capi.check(capi.lib().wk_unary(pipeline.q, net_output.type_index, 5, net_output.ptr, net_output.flat_elements("activation")))

    def _derivative(self, pipeline, output, derivative):
        capi.check(capi.lib().wk_tanh_dev(pipeline.q, output.type_index, output.ptr, derivative.ptr,
                                          output.flat_elements("activation derivative")))
