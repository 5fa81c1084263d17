//! Replaces src/tensor/random/main.zig (same export).
pub const uniform = @import("uniform.zig").uniform;
