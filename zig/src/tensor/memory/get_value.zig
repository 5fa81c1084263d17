//! Replaces src/tensor/memory/get_value.zig:12-47: one element, blocking read.
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("../main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

pub fn getValue(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T), coor: []const u64, scalar: *T) TensorErrors!void {
    if (coor.len != tensor.dimensions.shape.len) return tensor_module.Errors.InvalidCoordinates;
    var offset: usize = 0;
    for (tensor.dimensions.pitches, tensor.dimensions.shape, coor) |p, ds, c| {
        if (c >= ds) return tensor_module.Errors.InvalidCoordinates;
        offset += c * p;
    }
    try b200.check(b200.wk_get_value(pipeline.q(), tensor.buffer, offset * @sizeOf(T), @ptrCast(scalar), @sizeOf(T)));
}
