//! Replaces src/nn/activation/sigmoid.zig:39-137 -- Sigmoid(T): run = 1 / (1 + exp(-x)) in place, getDerivative = y (1 - y)
//! from the layer OUTPUT; both over the whole padded buffer like sigmoid.zig:62-68 (SURVEY Q1).  f32 / f64 only (:21-24).
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const activation = @import("main.zig");

pub fn Sigmoid(comptime T: type) type {
    const ActivationTensor = Tensor(T);
    const Activation = activation.Activation(T);

    switch (@typeInfo(T)) {
        .float => {},
        else => @compileError("Sigmoid activation only supports f32 and f64 types"),
    }

    return struct {
        /// the fused Linear paths (wk_gemm_bias_act / wk_act_backward) recognise the activation by this id
        pub const kind: i32 = b200.ACT_SIGMOID;

        pub fn init() Activation {
            return Activation{
                .vtable = .{ .run = &run, .getDerivative = &getDerivative },
                .ptr = undefined,
            };
        }

        pub fn deinit(_: *const anyopaque) void {}

        pub fn run(_: *const anyopaque, pipeline: *Pipeline, net_output: *ActivationTensor) !void {
            try b200.check(b200.wk_unary(pipeline.q(), core.types.getTypeIndex(T), b200.OP_SIGMOID, net_output.buffer, net_output.dimensions.number_of_elements));
        }

        pub fn getDerivative(_: *const anyopaque, pipeline: *Pipeline, output: *ActivationTensor, derivative: *ActivationTensor) !void {
            try b200.check(b200.wk_sigmoid_dev(pipeline.q(), core.types.getTypeIndex(T), output.buffer, derivative.buffer, output.dimensions.number_of_elements));
        }
    };
}

test {
    const std = @import("std");
    std.testing.refAllDecls(Sigmoid(f32));
    std.testing.refAllDecls(Sigmoid(f64));
}
