//! Replaces src/tensor/memory/main.zig (same exports).
pub const getValue = @import("get_value.zig").getValue;
pub const putValue = @import("put_value.zig").putValue;
pub const readFromBuffer = @import("read_from_buffer.zig").readFromBuffer;
pub const writeToBuffer = @import("write_to_buffer.zig").writeToBuffer;
pub const copy = @import("copy.zig").copy;
