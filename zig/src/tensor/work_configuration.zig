//! Replaces src/tensor/work_configuration.zig:48-193.  The reference pre-computes OpenCL ND-range local sizes and a GEMM
//! tile ("algorithm") per device for every tensor.  Launch geometry on the B200 backend is chosen inside the library, so
//! only the declarations other Zig files still name remain: the GemmAlgorithm enum (:9-16, the
//! `recommended_algorithm` argument of PackedTensors.initWithDimensions) and `gemm_algorithm_per_device`, which
//! src/nn/layer/linear.zig:358 reads when it builds its packed-tensor handles.
const std = @import("std");

pub const GemmAlgorithm = enum(u8) {
    @"2x2" = 0,
    @"4x4" = 1,
    @"8x8" = 2,
    @"16x16" = 3,
    @"32x32" = 4,
    @"64x64" = 5,
};

gemm_algorithm_per_device: []GemmAlgorithm,

/// one entry per CommandQueue of the context; the value is a hint nothing on this backend consumes
pub fn init(self: *WorkConfiguration, allocator: std.mem.Allocator, number_of_queues: usize) error{OutOfMemory}!void {
    self.gemm_algorithm_per_device = try allocator.alloc(GemmAlgorithm, number_of_queues);
    @memset(self.gemm_algorithm_per_device, .@"64x64");
}

const WorkConfiguration = @This();
