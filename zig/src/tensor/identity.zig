//! Replaces src/tensor/identity.zig:16-70.
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

/// zero the buffer, then 1 on the hyper-diagonal: element i sits at i * (sum of all pitches) (identity.cl:3-20)
pub fn identity(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    const size = tensor.dimensions.shape[0];
    for (tensor.dimensions.shape[1..]) |s| if (s != size) return tensor_module.Errors.InvalidValue;
    try tensor_module.fill.zeroes(T, pipeline, tensor);
    var pitch_sum: u64 = 0;
    for (tensor.dimensions.pitches) |p| pitch_sum += p;
    try b200.check(b200.wk_identity(pipeline.q(), core.types.getTypeIndex(T), tensor.buffer, tensor.dimensions.number_of_elements, size, pitch_sum));
}
