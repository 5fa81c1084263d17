//! Replaces src/tensor/random/uniform.zig:60-123.  The device kernel reproduces uniform.cl's counter-based generator
//! bit for bit (hash of the PADDED linear index and the seed), so a seed gives the same tensor as the reference does on a
//! device with the same vector width.
const std = @import("std");
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("../main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

pub fn uniform(
    comptime T: type,
    pipeline: *Pipeline,
    tensor: *Tensor(T),
    seed: ?u64,
    min_value: ?core.types.getType(T),
    max_value: ?core.types.getType(T),
) TensorErrors!void {
    const S = core.types.getType(T);
    const e = tensor.extent();
    const l = tensor.memory_layout;
    const s: u64 = seed orelse @bitCast(std.time.timestamp()); // uniform.zig:82: wall-clock seed when null
    try b200.check(b200.wk_uniform(pipeline.q(), core.types.getTypeIndex(T), e.depth, e.rows, e.cols, tensor.buffer, l.row_pitch, l.slice_pitch, s, b200.optPtr(S, &min_value), b200.optPtr(S, &max_value)));
}
