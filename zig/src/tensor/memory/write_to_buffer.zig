//! Replaces src/tensor/memory/write_to_buffer.zig:13-63 (clEnqueueReadBufferRect): tensor -> host, pitched, asynchronous.
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("../main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

pub fn writeToBuffer(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T), buffer: []T) TensorErrors!void {
    if (buffer.len != tensor.dimensions.number_of_elements_without_padding) return tensor_module.Errors.InvalidBuffer;
    const e = tensor.extent();
    const l = tensor.memory_layout;
    try b200.check(b200.wk_d2h_rect(pipeline.q(), @ptrCast(buffer.ptr), tensor.buffer, l.row_pitch * @sizeOf(T), l.slice_pitch * @sizeOf(T), e.cols * @sizeOf(T), e.rows, e.depth));
}
