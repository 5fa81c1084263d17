//! core module of the B200 overlay (replaces src/core/main.zig: same exports).
pub const Context = @import("context.zig");
pub const CommandQueue = @import("command_queue.zig");
pub const KernelsSet = @import("kernel.zig");
pub const Pipeline = @import("pipeline.zig");

/// src/core/types.zig is host-only (dtype table, Complex(T), getTypeIndex / getTypeId): used unchanged from the reference tree
pub const types = @import("types.zig");

/// the FFI surface (new)
pub const b200 = @import("b200.zig");

test {
    const std = @import("std");
    std.testing.refAllDecls(@This());
}
