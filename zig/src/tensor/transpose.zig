//! Replaces src/tensor/transpose.zig:15-113: swap two dimensions of an N-D tensor.
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("main.zig");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

pub fn transpose(
    comptime T: type,
    pipeline: *Pipeline,
    result_tensor: *Tensor(T),
    tensor: *Tensor(T),
    dim0: u64,
    dim1: u64,
) TensorErrors!void {
    const shape_a = result_tensor.dimensions.shape;
    const shape_b = tensor.dimensions.shape;
    if (shape_a.len != shape_b.len) return tensor_module.Errors.UnqualTensorsDimension;
    if (dim0 >= shape_a.len or dim1 >= shape_a.len) return tensor_module.Errors.InvalidValue;
    if (tensor.dimensions.number_of_elements_without_padding != result_tensor.dimensions.number_of_elements_without_padding)
        return tensor_module.Errors.UnqualTensorsDimension;
    if (shape_a[dim0] != shape_b[dim1] or shape_a[dim1] != shape_b[dim0]) return tensor_module.Errors.InvalidValue;
    if (dim0 == dim1) return tensor_module.memory.copy(T, pipeline, tensor, result_tensor);

    const dtype = core.types.getTypeIndex(T);
    const e = tensor.extent();
    const l = tensor.memory_layout;
    if (shape_a.len == 2) { // shared-memory tiled kernel, 128-bit accesses on both sides
        return b200.check(b200.wk_transpose2d(pipeline.q(), dtype, e.rows, e.cols, tensor.buffer, l.row_pitch, result_tensor.buffer, result_tensor.memory_layout.row_pitch));
    }
    try b200.check(b200.wk_transpose_nd(pipeline.q(), dtype, @intCast(shape_a.len), tensor.buffer, tensor.dimensions.pitches.ptr, result_tensor.buffer, result_tensor.dimensions.pitches.ptr, l.row_pitch, l.slice_pitch, e.rows * l.row_pitch, e.cols, tensor.dimensions.number_of_elements, @intCast(@min(dim0, dim1)), @intCast(@max(dim0, dim1))));
}
