//! Replaces src/nn/loss/mse.zig:63-132 -- err = (expected - output)^2 into the cache's error tensor and, when
//! `calculate_derivative`, -2 (expected - output) into the last layer's sensitivity, in ONE pass over the padded buffers
//! (mse.cl:3-35); optional mean of the error (:128-131).
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;
const math = @import("math");
const cache_module = @import("../layer/cache.zig");

pub fn mse(
    comptime T: type,
    comptime calculate_derivative: bool,
    pipeline: *Pipeline,
    output: *Tensor(T),
    expected: *Tensor(T),
    cache: *const cache_module.Cache(T),
    error_result: ?*T,
) TensorErrors!void {
    const error_tensor = cache.error_tensor;
    try tensor_module.helpers.eqlTensors(T, output, expected);
    try tensor_module.helpers.eqlTensors(T, error_tensor, output);

    var sensitivity_buffer: ?*anyopaque = null;
    if (calculate_derivative) {
        const last_slot = cache.slots[cache.slots.len - 1];
        sensitivity_buffer = last_slot.layer.getSensitivity(last_slot.cache).buffer;
    }
    try b200.check(b200.wk_mse(pipeline.q(), core.types.getTypeIndex(T), output.buffer, expected.buffer, error_tensor.buffer, sensitivity_buffer, output.dimensions.number_of_elements));

    if (error_result) |res| res.* = try math.basic.mean(T, pipeline, error_tensor);
}
