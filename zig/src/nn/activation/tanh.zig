//! Replaces src/nn/activation/tanh.zig:25-84 -- Tanh(T): run = math.tanh in place, getDerivative = 1 - y^2.
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const math = @import("math");
const activation = @import("main.zig");

pub fn Tanh(comptime T: type) type {
    const ActivationTensor = Tensor(T);
    const Activation = activation.Activation(T);

    switch (@typeInfo(T)) {
        .float => {},
        else => @compileError("Tanh activation only supports f32 and f64 types"),
    }

    return struct {
        pub const kind: i32 = b200.ACT_TANH;

        pub fn init() Activation {
            return Activation{
                .vtable = .{ .run = &run, .getDerivative = &getDerivative },
                .ptr = undefined,
            };
        }

        pub fn run(_: *const anyopaque, pipeline: *Pipeline, net_output: *ActivationTensor) !void {
            try math.trig.tanh(T, pipeline, net_output);
        }

        pub fn getDerivative(_: *const anyopaque, pipeline: *Pipeline, input: *ActivationTensor, derivative: *ActivationTensor) !void {
            try b200.check(b200.wk_tanh_dev(pipeline.q(), core.types.getTypeIndex(T), input.buffer, derivative.buffer, input.dimensions.number_of_elements));
        }
    };
}

test {
    const std = @import("std");
    std.testing.refAllDecls(Tanh(f32));
    std.testing.refAllDecls(Tanh(f64));
}
