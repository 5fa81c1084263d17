//! Replaces src/tensor/memory/read_from_buffer.zig:13-63 (clEnqueueWriteBufferRect): host -> tensor, pitched,
//! asynchronous -- `buffer` must stay alive until pipeline.waitAndCleanup(), the reference's own rule (:45-59).
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("../main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

pub fn readFromBuffer(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T), buffer: []const T) TensorErrors!void {
    if (buffer.len != tensor.dimensions.number_of_elements_without_padding) return tensor_module.Errors.InvalidBuffer;
    const e = tensor.extent();
    const l = tensor.memory_layout;
    try b200.check(b200.wk_h2d_rect(pipeline.q(), tensor.buffer, l.row_pitch * @sizeOf(T), l.slice_pitch * @sizeOf(T), @ptrCast(buffer.ptr), e.cols * @sizeOf(T), e.rows, e.depth));
}
