//! Replaces src/tensor/print.zig:316-361 (which maps the cl_mem to the host): download the logical elements, print rows.
//! Host-side convenience only (examples/xor_neural_network.zig:135-138); not on the accelerated path.
const std = @import("std");
const core = @import("core");
const Pipeline = core.Pipeline;
const tensor_module = @import("main.zig");
const Tensor = tensor_module.Tensor;

pub fn print(comptime T: type, pipeline: *Pipeline, tensor: *Tensor(T)) !void {
    const allocator = tensor.context.allocator;
    const host = try allocator.alloc(T, tensor.dimensions.number_of_elements_without_padding);
    defer allocator.free(host);
    try tensor_module.memory.writeToBuffer(T, pipeline, tensor, host);
    pipeline.waitAndCleanup();
    const cols = tensor.dimensions.shape[tensor.dimensions.shape.len - 1];
    std.debug.print("Tensor{any} =\n", .{tensor.dimensions.shape});
    for (host, 0..) |v, i| {
        std.debug.print("{any} ", .{v});
        if ((i + 1) % cols == 0) std.debug.print("\n", .{});
    }
}
