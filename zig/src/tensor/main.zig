//! Replaces src/tensor/main.zig:62-279 -- Tensor(T): HBM allocation with the reference's padded / pitched layout law.
//!
//! Same public declarations (Errors, CreateConfig, Tensor(T).empty / alloc / release, the `dimensions`,
//! `memory_layout`, `flags` records).  What changes: `buffer` is a CUDA device pointer from wk_malloc instead of a
//! cl_mem, and there is no device copy of the pitch array (`pitches_buffer`): kernels receive pitches as arguments.
//! `convertions` (toComplex / toReal) is outside the accelerated path and not part of this overlay.
const cl = @import("opencl");
const std = @import("std");

const core = @import("core");
const b200 = core.b200;
const Context = core.Context;
const Pipeline = core.Pipeline;

pub const helpers = @import("helpers.zig");

pub const fill = @import("fill.zig");
pub const memory = @import("memory/main.zig");
pub const random = @import("random/main.zig");
pub const transpose = @import("transpose.zig").transpose;
pub const identity = @import("identity.zig").identity;
pub const print = @import("print.zig").print;

const WorkConfiguration = @import("work_configuration.zig");
pub const GemmAlgorithm = WorkConfiguration.GemmAlgorithm;

pub const Errors = error{
    InvalidValue,
    InvalidCoordinates,
    InvalidBuffer,
    UnqualTensorsAttribute,
    UnqualTensorsShape,
    UnqualTensorsDimension,
    UnqualTensorsContext,
} || std.mem.Allocator.Error || cl.errors.OpenCLError || core.KernelsSet.Errors || b200.Error;

pub const CreateConfig = struct {
    /// accepted for source compatibility; HBM buffers are always read-write and never alias host memory
    cl_mem_flags: cl.buffer.MemFlags = cl.buffer.MemFlag.read_write,
    host_ptr: ?*anyopaque = null,
    vectors_enabled: bool = true,
};

const Dimensions = struct {
    shape: []u64,
    vl_shape: []u64,
    pitches: []u64,
    number_of_elements: u64,
    number_of_elements_without_padding: u64,
};

const MemoryLayout = struct {
    row_pitch: u64,
    row_pitch_for_vectors: u64,
    slice_pitch: u64,
    slice_pitch_for_vectors: u64,
    number_of_vectors: u64,
    size: usize,
};

const Flags = struct {
    vectors_enabled: bool,
};

pub fn Tensor(comptime T: type) type {
    const type_id = core.types.getTypeId(T);
    const is_complex = core.types.isComplex(T);

    return struct {
        context: *const Context,
        arena: std.heap.ArenaAllocator,

        /// CUDA device pointer (reference: cl_mem)
        buffer: cl.buffer.Mem,

        dimensions: Dimensions,
        work_configuration: WorkConfiguration,
        memory_layout: MemoryLayout,
        flags: Flags,

        const Self = @This();

        /// depth / rows / cols of the logical region: what the pitched (3-D) kernels iterate over
        pub const Extent = struct { depth: u64, rows: u64, cols: u64 };

        pub fn extent(self: *const Self) Extent {
            const shape = self.dimensions.shape;
            const nd = shape.len;
            var depth: u64 = 1;
            if (nd > 2) for (shape[0 .. nd - 2]) |e| {
                depth *= e;
            };
            return .{ .depth = depth, .rows = if (nd >= 2) shape[nd - 2] else 1, .cols = shape[nd - 1] };
        }

        /// Tensor.empty, main.zig:113-251: the layout law (vector-width round-up of the row pitch, even row pitch in
        /// vector units, even number of rows) followed by ONE device allocation
        pub fn empty(
            context: *const Context,
            pipeline: *Pipeline,
            shape: []const u64,
            config: CreateConfig,
        ) Errors!*Self {
            if (shape.len == 0) return Errors.InvalidValue;
            for (shape) |s| if (s == 0) return Errors.InvalidValue;

            const allocator = context.allocator;
            const tensor = try allocator.create(Self);
            errdefer allocator.destroy(tensor);
            tensor.context = context;
            tensor.arena = std.heap.ArenaAllocator.init(allocator);
            errdefer tensor.arena.deinit();
            const arena = tensor.arena.allocator();

            // widest vector any queue of the context reports for this element type (main.zig:142-150)
            var vector_width: u64 = 1;
            var vectors_enabled = !is_complex and config.vectors_enabled;
            if (vectors_enabled) {
                for (context.command_queues) |cmd| vector_width = @max(vector_width, @as(u64, cmd.vector_widths[type_id]));
                vectors_enabled = vector_width > 1;
            }
            if (!vectors_enabled) vector_width = 1;
            tensor.flags.vectors_enabled = vectors_enabled;

            const ndim = shape.len;
            const last = ndim - 1;
            const pen = last -| 1;
            tensor.dimensions.shape = try arena.dupe(u64, shape);
            const vl_shape = try arena.dupe(u64, shape);
            tensor.dimensions.vl_shape = vl_shape;

            var depth: u64 = 1;
            for (shape[0..pen]) |e| depth *= e;
            const rows: u64 = if (ndim >= 2) shape[pen] else 1;
            const cols: u64 = shape[last];
            const rows_padded = rows + rows % 2; // :168
            tensor.dimensions.number_of_elements_without_padding = depth * rows * cols;

            var row_pitch = std.mem.alignForward(u64, cols, vector_width); // :174-180
            var row_pitch_for_vectors = row_pitch / vector_width;
            vl_shape[last] = row_pitch_for_vectors;
            if (row_pitch_for_vectors % 2 == 1) { // :185-187
                row_pitch_for_vectors += 1;
                row_pitch += vector_width;
            }
            const slice_pitch = row_pitch * rows_padded;
            const number_of_elements = slice_pitch * depth;
            tensor.dimensions.number_of_elements = number_of_elements;
            tensor.memory_layout = .{
                .row_pitch = row_pitch,
                .row_pitch_for_vectors = row_pitch_for_vectors,
                .slice_pitch = slice_pitch,
                .slice_pitch_for_vectors = slice_pitch / vector_width,
                .number_of_vectors = number_of_elements / vector_width,
                .size = number_of_elements * @sizeOf(T),
            };

            // element pitches of every dimension (:196-222)
            const pitches = try arena.alloc(u64, ndim);
            tensor.dimensions.pitches = pitches;
            pitches[last] = 1;
            if (ndim >= 2) pitches[pen] = row_pitch;
            if (ndim >= 3) {
                pitches[pen - 1] = slice_pitch;
                var i = pen - 1;
                while (i > 0) : (i -= 1) pitches[i - 1] = pitches[i] * shape[i];
            }

            try tensor.work_configuration.init(arena, context.command_queues.len);

            var dptr: ?*anyopaque = null;
            try b200.check(b200.wk_malloc(pipeline.q(), tensor.memory_layout.size, &dptr));
            tensor.buffer = dptr;
            return tensor;
        }

        /// main.zig:253-263: freeing a tensor synchronises the pipeline first (wk_free drains the stream)
        pub fn release(self: *Self, pipeline: *Pipeline) void {
            b200.check(b200.wk_free(pipeline.q(), self.buffer)) catch |err| {
                std.debug.panic("wk_free failed ({s}): {s}", .{ @errorName(err), b200.wk_last_error() });
            };
            const allocator = self.context.allocator;
            self.arena.deinit();
            allocator.destroy(self);
        }

        /// main.zig:265-277: empty + zero the whole padded buffer
        pub fn alloc(
            context: *const Context,
            pipeline: *Pipeline,
            shape: []const u64,
            config: CreateConfig,
        ) Errors!*Self {
            const tensor = try empty(context, pipeline, shape, config);
            errdefer tensor.release(pipeline);
            try fill.zeroes(T, pipeline, tensor);
            return tensor;
        }
    };
}

test {
    std.testing.refAllDecls(@This());
}
