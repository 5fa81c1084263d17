//! Replaces src/tensor/memory/copy.zig:85-98 (clEnqueueCopyBuffer / clEnqueueCopyBufferRect).
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("../main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;
const helpers = tensor_module.helpers;

pub fn copy(comptime T: type, pipeline: *Pipeline, src: *Tensor(T), dst: *Tensor(T)) TensorErrors!void {
    try helpers.eqlTensorsShape(T, src, dst);
    const s = src.memory_layout;
    const d = dst.memory_layout;
    if (s.row_pitch == d.row_pitch) { // same layout: one flat copy of the padded buffer
        return b200.check(b200.wk_d2d(pipeline.q(), dst.buffer, src.buffer, s.size));
    }
    const e = src.extent();
    try b200.check(b200.wk_d2d_rect(pipeline.q(), dst.buffer, d.row_pitch * @sizeOf(T), d.slice_pitch * @sizeOf(T), src.buffer, s.row_pitch * @sizeOf(T), s.slice_pitch * @sizeOf(T), e.cols * @sizeOf(T), e.rows, e.depth));
}
