//! The device side of src/nn/layer/linear.zig on the B200 backend.
//!
//! linear.zig is 700 lines of host glue (cache construction, the Layer vtable, the forward / backward call order) around
//! exactly two kernel launches; the glue stays the reference's file.  The overlay is this module plus a three-line edit:
//!
//!     const linear_b200 = @import("linear_b200.zig");
//!     fn addBias(pipeline, output, bias_tensor)                 -> return linear_b200.addBias(T, pipeline, output, bias_tensor);
//!     fn getBiasSensitivity(pipeline, sensitivity, bias_gradient, lwi) -> return linear_b200.getBiasSensitivity(T, pipeline, sensitivity, bias_gradient);
//!
//! and, optionally, the fused fast paths below in forward (:480-525) and backward (:608-613).
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

/// linear.zig:424-478 + bias.cl:3-19: out[i] += bias[i % row_pitch] over the whole padded buffer
pub fn addBias(comptime T: type, pipeline: *Pipeline, output: *Tensor(T), bias_tensor: *Tensor(T)) TensorErrors!void {
    try b200.check(b200.wk_bias_add(pipeline.q(), core.types.getTypeIndex(T), output.buffer, bias_tensor.buffer, output.memory_layout.row_pitch, output.dimensions.number_of_elements));
}

/// linear.zig:534-577 + bias_step.cl:3-39: column sums of the sensitivity over its shape[0] logical rows
pub fn getBiasSensitivity(comptime T: type, pipeline: *Pipeline, sensitivity: *Tensor(T), bias_gradient: *Tensor(T)) TensorErrors!void {
    try b200.check(b200.wk_bias_step(pipeline.q(), core.types.getTypeIndex(T), sensitivity.buffer, bias_gradient.buffer, sensitivity.memory_layout.row_pitch, sensitivity.dimensions.shape[0], bias_gradient.memory_layout.row_pitch));
}

/// fused forward of one sub-layer: output = act(input . weight^T + bias) in the GEMM epilogue (one launch instead of
/// gemm + bias.cl + sigmoid.cl / tanh; bit-identical to the unfused sequence on logical elements)
pub fn forwardFused(comptime T: type, pipeline: *Pipeline, input: *Tensor(T), weight: *Tensor(T), bias_tensor: ?*Tensor(T), activation_kind: i32, output: *Tensor(T)) TensorErrors!void {
    const m = output.dimensions.shape[0];
    const n = output.dimensions.shape[1];
    const k = input.dimensions.shape[1];
    try b200.check(b200.wk_gemm_bias_act(pipeline.q(), core.types.getTypeIndex(T), 0, 1, m, n, k, input.buffer, input.memory_layout.row_pitch, weight.buffer, weight.memory_layout.row_pitch, output.buffer, output.memory_layout.row_pitch, if (bias_tensor) |b| b.buffer else null, activation_kind));
}

/// fused first step of backward: derivative = act'(output) (optional) and sensitivity *= act'(output) in one pass
/// (replaces getDerivative + math.dot, linear.zig:608-613)
pub fn backwardActivation(comptime T: type, pipeline: *Pipeline, activation_kind: i32, output: *Tensor(T), derivative: ?*Tensor(T), sensitivity: *Tensor(T)) TensorErrors!void {
    try b200.check(b200.wk_act_backward(pipeline.q(), core.types.getTypeIndex(T), activation_kind, output.buffer, if (derivative) |d| d.buffer else null, sensitivity.buffer, sensitivity.dimensions.number_of_elements));
}

/// the whole backward step of one sub-layer (linear.zig:579-678: getDerivative, math.dot, gemm TN, getBiasSensitivity, gemm NN)
/// as ONE call -- two launches for f32 layers on the tensor-core path; `sensitivity` is consumed
pub fn backwardFused(comptime T: type, pipeline: *Pipeline, activation_kind: i32, sensitivity: *Tensor(T), output: *Tensor(T), prev_output: *Tensor(T), weight: *Tensor(T), gradient: *Tensor(T), bias_gradient: ?*Tensor(T), next_sensitivity: ?*Tensor(T)) TensorErrors!void {
    const batch = sensitivity.dimensions.shape[0];
    const n_out = sensitivity.dimensions.shape[1];
    const n_in = weight.dimensions.shape[1];
    try b200.check(b200.wk_linear_backward(pipeline.q(), core.types.getTypeIndex(T), activation_kind, batch, n_out, n_in, sensitivity.buffer, sensitivity.memory_layout.row_pitch, output.buffer, output.memory_layout.row_pitch, prev_output.buffer, prev_output.memory_layout.row_pitch, weight.buffer, weight.memory_layout.row_pitch, gradient.buffer, gradient.memory_layout.row_pitch, if (bias_gradient) |b| b.buffer else null, if (next_sensitivity) |n| n.buffer else null, if (next_sensitivity) |n| n.memory_layout.row_pitch else 0));
}
