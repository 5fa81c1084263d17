//! Replaces src/blas/gemm.zig:25-29,60-359,442-485,834-874 -- blas.gemm and its PackedTensors handle.
//!
//!   C = alpha * op(A) * op(B) + beta * C       alpha == null and beta == null: C = A.B;  beta == null: C is overwritten;
//!                                              alpha == null with beta: alpha = 1 (gemm.zig:500-501,592-596)
//! The reference JIT-compiles one of seven OpenCL-C kernel texts per (tile, vector, alpha, beta, opA, opB, dtype) and
//! re-tiles both operands into scratch tensors on every packed call (gemm.zig:272-357).  Here ONE call reaches the library:
//! f32 runs as 3xTF32 on tcgen05 tensor cores with TMEM accumulators, f64 on FP64 tensor-core MMA, integers on
//! tcgen05 kind::i8 over byte planes (exact mod 2^bits); TMA reads the operands where they lie, so PackedTensors
//! owns no memory -- it stays as a shape-checked handle because Linear and the benchmark construct one.
const std = @import("std");
const core = @import("core");
const b200 = core.b200;
const Pipeline = core.Pipeline;
const tensor_module = @import("tensor");
const Tensor = tensor_module.Tensor;
const TensorErrors = tensor_module.Errors;
const GemmAlgorithm = tensor_module.GemmAlgorithm;

/// gemm.zig:25-29
pub const Operation = enum(u8) {
    no_transpose = 0,
    transpose = 1,
};

pub fn PackedTensors(comptime T: type) type {
    const TensorT = Tensor(T);
    return struct {
        m_size: u64,
        n_size: u64,
        k_size: u64,
        vectors_enabled: bool,
        algorithm: GemmAlgorithm,

        const Self = @This();

        /// gemm.zig:77-97: sized from the result tensor [n, m] and the inner dimension
        pub fn init(pipeline: *Pipeline, result_tensor: *TensorT, k_size: u64, vectors_enabled: bool) TensorErrors!*Self {
            const shape = result_tensor.dimensions.shape;
            if (shape.len != 2) return tensor_module.Errors.InvalidValue;
            return initWithDimensions(pipeline, shape[0], shape[1], k_size, .@"64x64", vectors_enabled);
        }

        /// gemm.zig:99-115
        pub fn initWithDimensions(
            pipeline: *Pipeline,
            n_size: u64,
            m_size: u64,
            k_size: u64,
            recommended_algorithm: GemmAlgorithm,
            vectors_enabled: bool,
        ) TensorErrors!*Self {
            if (n_size == 0 or m_size == 0 or k_size == 0) return tensor_module.Errors.InvalidValue;
            const self = try pipeline.allocator.create(Self);
            self.* = .{ .m_size = m_size, .n_size = n_size, .k_size = k_size, .vectors_enabled = vectors_enabled, .algorithm = recommended_algorithm };
            return self;
        }

        /// gemm.zig:188-193
        pub fn deinit(self: *Self, pipeline: *Pipeline) void {
            pipeline.allocator.destroy(self);
        }

        /// gemm.zig:250-270
        inline fn validateTensors(self: *Self, a: *TensorT, op_a: Operation, b: *TensorT, op_b: Operation) TensorErrors!void {
            const as = a.dimensions.shape;
            const bs = b.dimensions.shape;
            var valid = switch (op_a) {
                .no_transpose => (as[0] == self.n_size and as[1] == self.k_size),
                .transpose => (as[1] == self.n_size and as[0] == self.k_size),
            };
            valid = valid and switch (op_b) {
                .no_transpose => (bs[0] == self.k_size and bs[1] == self.m_size),
                .transpose => (bs[1] == self.k_size and bs[0] == self.m_size),
            };
            if (!valid) return tensor_module.Errors.InvalidValue;
        }

        /// gemm.zig:272-357 launches two re-tiling kernels; here only the shape contract is checked
        pub fn pack(self: *Self, pipeline: *Pipeline, a: *TensorT, op_a: Operation, b: *TensorT, op_b: Operation) TensorErrors!void {
            _ = pipeline;
            try self.validateTensors(a, op_a, b, op_b);
        }
    };
}

/// gemm.zig:442-485
inline fn validateTensors(comptime T: type, a: *Tensor(T), b: *Tensor(T), c: *Tensor(T), op_a: Operation, op_b: Operation) TensorErrors!void {
    if (a.context != b.context or a.context != c.context) return tensor_module.Errors.UnqualTensorsContext;
    const as = a.dimensions.shape;
    const bs = b.dimensions.shape;
    const cs = c.dimensions.shape;
    if (cs.len != 2 or as.len != 2 or bs.len != 2) return tensor_module.Errors.InvalidValue;
    // op(A) is [m, k], op(B) is [k, n]
    const m = as[@intFromEnum(op_a)];
    const k = as[1 - @intFromEnum(op_a)];
    const kb = bs[@intFromEnum(op_b)];
    const n = bs[1 - @intFromEnum(op_b)];
    if (k != kb or m != cs[0] or n != cs[1]) return tensor_module.Errors.InvalidValue;
}

/// gemm.zig:834-874
pub fn gemm(
    comptime T: type,
    pipeline: *Pipeline,
    alpha: ?T,
    a: *Tensor(T),
    op_a: Operation,
    b: *Tensor(T),
    op_b: Operation,
    beta: ?T,
    c: *Tensor(T),
    packed_tensors: ?*PackedTensors(T),
) TensorErrors!void {
    try validateTensors(T, a, b, c, op_a, op_b);
    if (packed_tensors) |p| try p.pack(pipeline, a, op_a, b, op_b);
    const m = c.dimensions.shape[0];
    const n = c.dimensions.shape[1];
    const k = a.dimensions.shape[1 - @intFromEnum(op_a)];
    try b200.check(b200.wk_gemm(pipeline.q(), core.types.getTypeIndex(T), @intFromEnum(op_a), @intFromEnum(op_b), m, n, k, b200.optPtr(T, &alpha), a.buffer, a.memory_layout.row_pitch, b.buffer, b.memory_layout.row_pitch, b200.optPtr(T, &beta), c.buffer, c.memory_layout.row_pitch));
    try finishPadding(T, pipeline, c, beta);
}

/// The reference's kernels run over the whole PADDED C: the packed path leaves beta * padding (or 0) there, and later
/// whole-buffer ops (sum.cl adds the padded columns, SURVEY Q2) see it.  The library writes logical elements only, so the
/// pad column / pad row are finished here (no-ops for even shapes, i.e. every benchmark shape).
fn finishPadding(comptime T: type, pipeline: *Pipeline, c: *Tensor(T), beta: ?T) TensorErrors!void {
    const m = c.dimensions.shape[0];
    const n = c.dimensions.shape[1];
    const l = c.memory_layout;
    const rows_padded = l.slice_pitch / l.row_pitch;
    const dtype = core.types.getTypeIndex(T);
    const base: [*]u8 = @ptrCast(c.buffer.?);
    if (l.row_pitch > n) { // pad columns of every (padded) row
        const region: ?*anyopaque = @ptrCast(base + n * @sizeOf(T));
        if (beta) |bv| {
            try b200.check(b200.wk_scal(pipeline.q(), dtype, 1, rows_padded, l.row_pitch - n, @ptrCast(&bv), region, l.row_pitch, l.slice_pitch));
        } else {
            const zero: T = std.mem.zeroes(T);
            try b200.check(b200.wk_fill(pipeline.q(), dtype, 1, rows_padded, l.row_pitch - n, region, l.row_pitch, l.slice_pitch, @ptrCast(&zero)));
        }
    }
    if (rows_padded > m) { // the pad row, logical columns
        const region: ?*anyopaque = @ptrCast(base + m * l.row_pitch * @sizeOf(T));
        if (beta) |bv| {
            try b200.check(b200.wk_scal(pipeline.q(), dtype, 1, rows_padded - m, n, @ptrCast(&bv), region, l.row_pitch, l.slice_pitch));
        } else {
            const zero: T = std.mem.zeroes(T);
            try b200.check(b200.wk_fill(pipeline.q(), dtype, 1, rows_padded - m, n, region, l.row_pitch, l.slice_pitch, @ptrCast(&zero)));
        }
    }
}

test {
    std.testing.refAllDecls(@This());
}
