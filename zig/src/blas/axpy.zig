//! Replaces src/blas/axpy.zig:93-169 -- y += alpha * x over the logical region (128-bit coalesced streaming kernel).
//! The reference picks one of three compiled variants on the host (`y += x`, `y += alpha x`, `y -= x` when alpha == -1,
//! axpy.zig:66-91); the library makes the same choice from the scalar it is handed, so the body is one call.
//! `scal` and `dotReduce` are new ops (no definition under src/: SURVEY a13; semantics of old_src/blas.c:41-67 and
//! the textbook reduction).
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

pub fn axpy(comptime T: type, pipeline: *Pipeline, x: *Tensor(T), alpha: ?T, y: *Tensor(T)) TensorErrors!void {
    try tensor_module.helpers.eqlTensorsShape(T, x, y);
    const e = x.extent();
    const lx = x.memory_layout;
    const ly = y.memory_layout;
    try b200.check(b200.wk_axpy(pipeline.q(), core.types.getTypeIndex(T), e.depth, e.rows, e.cols, b200.optPtr(T, &alpha), x.buffer, lx.row_pitch, lx.slice_pitch, y.buffer, ly.row_pitch, ly.slice_pitch));
}

/// x *= alpha
pub fn scal(comptime T: type, pipeline: *Pipeline, alpha: T, x: *Tensor(T)) TensorErrors!void {
    const e = x.extent();
    const l = x.memory_layout;
    try b200.check(b200.wk_scal(pipeline.q(), core.types.getTypeIndex(T), e.depth, e.rows, e.cols, @ptrCast(&alpha), x.buffer, l.row_pitch, l.slice_pitch));
}

/// sum(x * y) -- blocking, deterministic (fixed-order two-stage reduction)
pub fn dotReduce(comptime T: type, pipeline: *Pipeline, x: *Tensor(T), y: *Tensor(T)) TensorErrors!T {
    try tensor_module.helpers.eqlTensorsShape(T, x, y);
    const e = x.extent();
    const lx = x.memory_layout;
    const ly = y.memory_layout;
    var result: T = undefined;
    try b200.check(b200.wk_dot_reduce(pipeline.q(), core.types.getTypeIndex(T), e.depth, e.rows, e.cols, x.buffer, lx.row_pitch, lx.slice_pitch, y.buffer, ly.row_pitch, ly.slice_pitch, @ptrCast(&result)));
    return result;
}
