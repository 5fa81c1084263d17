//! Replaces src/math/trig.zig:15-113 -- six in-place unary maps over the WHOLE padded buffer (trig.zig:45-51).
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

fn genericTrigFunction(comptime T: type, comptime op: i32, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try b200.check(b200.wk_unary(pipeline.q(), core.types.getTypeIndex(T), op, tensor.buffer, tensor.dimensions.number_of_elements));
}

pub inline fn sin(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try genericTrigFunction(T, b200.OP_SIN, pipeline, tensor);
}
pub inline fn cos(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try genericTrigFunction(T, b200.OP_COS, pipeline, tensor);
}
pub inline fn tan(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try genericTrigFunction(T, b200.OP_TAN, pipeline, tensor);
}
pub inline fn sinh(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try genericTrigFunction(T, b200.OP_SINH, pipeline, tensor);
}
pub inline fn cosh(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try genericTrigFunction(T, b200.OP_COSH, pipeline, tensor);
}
pub inline fn tanh(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) TensorErrors!void {
    try genericTrigFunction(T, b200.OP_TANH, pipeline, tensor);
}
