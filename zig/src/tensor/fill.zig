//! Replaces src/tensor/fill.zig:15-95.
const std = @import("std");
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

/// fill.zig:16-57: the LOGICAL region only (3-D range over [depth, rows, cols])
pub fn constant(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T), scalar: T) TensorErrors!void {
    const e = tensor.extent();
    const l = tensor.memory_layout;
    try b200.check(b200.wk_fill(pipeline.q(), core.types.getTypeIndex(T), e.depth, e.rows, e.cols, tensor.buffer, l.row_pitch, l.slice_pitch, @ptrCast(&scalar)));
}

/// fill.zig:59-68
pub inline fn one(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) !void {
    try constant(T, pipeline, tensor, core.types.getOne(T));
}

/// fill.zig:70-95: clEnqueueFillBuffer over the WHOLE padded buffer -> cudaMemsetAsync
pub fn zeroes(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try b200.check(b200.wk_memset_zero(pipeline.q(), tensor.buffer, tensor.memory_layout.size));
}
