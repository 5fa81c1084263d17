//! Replaces src/blas/main.zig (same exports; `scal` and `dotReduce` are new ops the north star names, SURVEY a13).
const axpy_module = @import("axpy.zig");
const gemm_module = @import("gemm.zig");

pub const axpy = axpy_module.axpy;
pub const scal = axpy_module.scal;
pub const dotReduce = axpy_module.dotReduce;
pub const gemm = gemm_module.gemm;
pub const GemmPackedTensors = gemm_module.PackedTensors;
pub const GemmOperation = gemm_module.Operation;

test {
    const std = @import("std");
    std.testing.refAllDecls(@This());
}
