//! Replaces src/tensor/helpers.zig: the cl_event helpers are gone with the events; the two shape predicates stay.
const std = @import("std");
const tensor_module = @import("main.zig");
const Tensor = tensor_module.Tensor;
const Errors = tensor_module.Errors;

/// helpers.zig:53-57
pub inline fn eqlTensorsShape(comptime T: type, tensor_a: *Tensor(T), tensor_b: *Tensor(T)) Errors!void {
    if (!std.mem.eql(u64, tensor_a.dimensions.shape, tensor_b.dimensions.shape)) return Errors.UnqualTensorsShape;
}

/// helpers.zig:59-65
pub inline fn eqlTensors(comptime T: type, tensor_a: *Tensor(T), tensor_b: *Tensor(T)) Errors!void {
    try eqlTensorsShape(T, tensor_a, tensor_b);
    if (tensor_a.flags.vectors_enabled != tensor_b.flags.vectors_enabled) return Errors.UnqualTensorsAttribute;
}
