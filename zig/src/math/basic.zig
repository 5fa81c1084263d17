//! Replaces src/math/basic.zig:17-240 -- dot (in-place Hadamard product), sum, mean.
const std = @import("std");
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;

/// basic.zig:17-76: x *= y element by element (NOT a reduction; dot.cl:33)
pub fn dot(comptime T: type, pipeline: *Pipeline, x: *Tensor(T), y: *Tensor(T)) TensorErrors!void {
    try tensor_module.helpers.eqlTensorsShape(T, x, y);
    const e = x.extent();
    const lx = x.memory_layout;
    const ly = y.memory_layout;
    try b200.check(b200.wk_hadamard(pipeline.q(), core.types.getTypeIndex(T), e.depth, e.rows, e.cols, x.buffer, lx.row_pitch, lx.slice_pitch, y.buffer, ly.row_pitch, ly.slice_pitch));
}

/// basic.zig:128-203.  The reference launches sum.cl (row sums over the PADDED row, `k < row_pitch`), maps the
/// [1, rows] temporary and adds it on the host.  Here one launch reduces the same elements on the device
/// (warp shuffles, per-block partials folded in a fixed order) and the call blocks where the reference maps.
pub fn sum(comptime T: type, pipeline: *Pipeline, x: *Tensor(T)) TensorErrors!T {
    const shape = x.dimensions.shape;
    const last_dim = shape[shape.len - 1];
    const l = x.memory_layout;
    var result: T = std.mem.zeroes(T);
    if (last_dim > 1) {
        const e = x.extent();
        try b200.check(b200.wk_sum(pipeline.q(), core.types.getTypeIndex(T), e.depth, e.rows, l.row_pitch, l.slice_pitch, x.buffer, @ptrCast(&result)));
    } else {
        // basic.zig:150-152,193-202: the tensor itself is mapped and its first `row_length` elements are added
        var row_length: u64 = 1;
        for (shape[0 .. shape.len - 1]) |s| row_length *= s;
        try b200.check(b200.wk_sum(pipeline.q(), core.types.getTypeIndex(T), 1, 1, row_length, row_length, x.buffer, @ptrCast(&result)));
    }
    return result;
}

/// basic.zig:206-240: sum / number of UNPADDED elements (@divTrunc for integers, each component for complex)
pub fn mean(comptime T: type, pipeline: *Pipeline, x: *Tensor(T)) TensorErrors!T {
    var result = try sum(T, pipeline, x);
    const n = x.dimensions.number_of_elements_without_padding;
    const SubType = core.types.getType(T);
    if (comptime core.types.isComplex(T)) {
        result.real = divide(SubType, result.real, n);
        result.imag = divide(SubType, result.imag, n);
    } else {
        result = divide(SubType, result, n);
    }
    return result;
}

inline fn divide(comptime S: type, v: S, n: u64) S {
    return switch (@typeInfo(S)) {
        .float => v / @as(S, @floatFromInt(n)),
        .int => @divTrunc(v, @as(S, @intCast(n))),
        else => unreachable,
    };
}
