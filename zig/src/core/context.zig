//! Replaces src/core/context.zig:13-190 -- a Context owns N CommandQueues, one per CUDA device
//! (reference: one cl_context + one in-order cl_command_queue per OpenCL device).
//! Public signatures are the reference's; the bodies are one FFI call each.
const std = @import("std");
const cl = @import("opencl");

const b200 = @import("b200.zig");
const CommandQueue = @import("command_queue.zig");

pub const Errors = cl.errors.OpenCLError || b200.Error;

allocator: std.mem.Allocator,
/// reference field name kept; holds the library's wk_context
cl_context: *b200.Context,
command_queues: []CommandQueue,

/// Context.init, context.zig:13-25: `devices` are CUDA ordinals (cl.device.DeviceId = i32)
pub fn init(
    allocator: std.mem.Allocator,
    properties: ?[]const cl.context.Properties,
    devices: []cl.device.DeviceId,
) Errors!*Context {
    _ = properties;
    if (devices.len == 0) return error.DevicesArrayEmpty;
    var handle: ?*b200.Context = null;
    try b200.check(b200.wk_context_create(devices.ptr, @intCast(devices.len), &handle));
    errdefer _ = b200.wk_context_destroy(handle.?);
    return fromHandle(allocator, handle.?);
}

/// context.zig:27-37: every CUDA device of the box (any selector; there are only GPUs)
pub fn initFromDeviceType(
    allocator: std.mem.Allocator,
    properties: ?[]const cl.context.Properties,
    device_type: cl.device.Type,
) Errors!*Context {
    _ = properties;
    _ = device_type;
    var handle: ?*b200.Context = null;
    try b200.check(b200.wk_context_create_all(&handle));
    errdefer _ = b200.wk_context_destroy(handle.?);
    return fromHandle(allocator, handle.?);
}

/// context.zig:39-104 scores devices by sub-devices x work-group size (SURVEY Q8); the B200s of a box are
/// identical, so the best device is ordinal 0
pub fn initFromBestDevice(
    allocator: std.mem.Allocator,
    properties: ?[]const cl.context.Properties,
    device_type: cl.device.Type,
) Errors!*Context {
    _ = device_type;
    var ordinals = [_]cl.device.DeviceId{0};
    return init(allocator, properties, &ordinals);
}

/// context.zig:106-143: there is one "platform" (CUDA) -> one context over all devices
pub fn createOnePerPlatform(
    allocator: std.mem.Allocator,
    properties: ?[]const cl.context.Properties,
    device_type: cl.device.Type,
) Errors![]*Context {
    const contexts = try allocator.alloc(*Context, 1);
    errdefer allocator.free(contexts);
    contexts[0] = try initFromDeviceType(allocator, properties, device_type);
    return contexts;
}

fn fromHandle(allocator: std.mem.Allocator, handle: *b200.Context) Errors!*Context {
    const context = try allocator.create(Context);
    errdefer allocator.destroy(context);
    context.allocator = allocator;
    context.cl_context = handle;
    var n: i32 = 0;
    try b200.check(b200.wk_context_num_queues(handle, &n));
    context.command_queues = try CommandQueue.initMultiples(allocator, context, @intCast(n));
    return context;
}

pub fn deinit(context: *Context) void {
    const allocator = context.allocator;
    CommandQueue.deinitMultiples(allocator, context.command_queues);
    _ = b200.wk_context_destroy(context.cl_context);
    allocator.destroy(context);
}

pub fn deinitMultiples(allocator: std.mem.Allocator, contexts: []*Context) void {
    if (contexts.len == 0) std.debug.panic("Contexts array is empty", .{});
    for (contexts) |ctx| ctx.deinit();
    allocator.free(contexts);
}

const Context = @This();
