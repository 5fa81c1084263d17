//! Replaces src/core/command_queue.zig:10-229 -- one in-order queue per device = one CUDA stream.
//! The capability record callers read (vector_widths, local_mem_type, max_work_group_size, compute_units, wekua_id) is
//! filled from wk_queue_info instead of ~19 clGetDeviceInfo calls (:30-123); there is no kernel cache to own.
const std = @import("std");
const cl = @import("opencl");

const b200 = @import("b200.zig");
const types = @import("types.zig");
const Context = @import("context.zig");

pub const Errors = cl.errors.OpenCLError || b200.Error;

context: *const Context,
/// reference field name kept; holds the library's wk_queue (a CUDA stream + its scratch buffers)
cl_command_queue: *b200.Queue,
device: cl.device.DeviceId,

device_name: []u8,
device_vendor_id: u32,
device_type: cl.device.Type,

local_mem_type: cl.device.LocalMemType,
local_mem_size: u64,

compute_units: u32,
vector_widths: [10]u32,
max_work_group_size: u64,
cache_line_size: u32,
wekua_id: usize,

/// the handle every FFI call takes
pub inline fn handle(self: *const CommandQueue) *b200.Queue {
    return self.cl_command_queue;
}

pub fn init(self: *CommandQueue, ctx: *const Context, index: usize) Errors!void {
    var q: ?*b200.Queue = null;
    try b200.check(b200.wk_context_queue(ctx.cl_context, @intCast(index), &q));
    var info: b200.QueueInfo = undefined;
    try b200.check(b200.wk_queue_info(q.?, &info));

    const name_len = std.mem.indexOfScalar(u8, &info.device_name, 0) orelse info.device_name.len;
    const name = try ctx.allocator.alloc(u8, name_len);
    @memcpy(name, info.device_name[0..name_len]);

    self.* = .{
        .context = ctx,
        .cl_command_queue = q.?,
        .device = info.device_ordinal,
        .device_name = name,
        .device_vendor_id = 0x10DE,
        .device_type = cl.device.Type.gpu,
        .local_mem_type = @enumFromInt(@as(u32, @intCast(info.local_mem_type))),
        .local_mem_size = info.local_mem_size,
        .compute_units = info.compute_units,
        .vector_widths = undefined,
        .max_work_group_size = info.max_work_group_size,
        .cache_line_size = info.cache_line_size,
        .wekua_id = index,
    };
    for (&self.vector_widths, info.vector_widths) |*vw, w| vw.* = @min(@as(u32, w), 16); // command_queue.zig:97
}

pub fn initMultiples(allocator: std.mem.Allocator, ctx: *const Context, n: usize) Errors![]CommandQueue {
    const command_queues = try allocator.alloc(CommandQueue, n);
    var created: usize = 0;
    errdefer {
        for (command_queues[0..created]) |*cmd| cmd.deinit();
        allocator.free(command_queues);
    }
    for (command_queues, 0..) |*cmd, index| {
        try cmd.init(ctx, index);
        created += 1;
    }
    return command_queues;
}

/// the queues belong to the wk_context (destroyed by Context.deinit); drain the stream like clFinish (:213-217)
pub fn deinit(self: *CommandQueue) void {
    self.context.allocator.free(self.device_name);
    b200.check(b200.wk_queue_finish(self.cl_command_queue)) catch |err| {
        std.debug.panic("An error ocurred while draining the CUDA stream: {s}", .{@errorName(err)});
    };
}

pub fn deinitMultiples(allocator: std.mem.Allocator, command_queues: []CommandQueue) void {
    for (command_queues) |*cmd| cmd.deinit();
    allocator.free(command_queues);
}

pub inline fn isTypeSupported(self: *const CommandQueue, comptime T: type) bool {
    return (self.vector_widths[types.getTypeId(T)] > 0);
}

const CommandQueue = @This();
