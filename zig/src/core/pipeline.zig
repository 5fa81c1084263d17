//! Replaces src/core/pipeline.zig:6-66.  In the reference a Pipeline is a list of cl_events: every op waits on
//! `prevEvents()`, enqueues, and `append`s its own event -- a serial dependency chain.  One in-order CUDA stream per
//! CommandQueue already IS that chain, so the list stays empty: `prevEvents()` returns null, `append` records
//! nothing, and `waitAndCleanup()` drains the stream.  Signatures are the reference's.
const std = @import("std");
const cl = @import("opencl");

const b200 = @import("b200.zig");
const CommandQueue = @import("command_queue.zig");

command_queue: *CommandQueue,
allocator: std.mem.Allocator,

pub fn init(command_queue: *CommandQueue) error{OutOfMemory}!*Pipeline {
    const allocator = command_queue.context.allocator;
    const self = try allocator.create(Pipeline);
    self.* = .{ .allocator = allocator, .command_queue = command_queue };
    return self;
}

pub fn deinit(self: *Pipeline) void {
    self.allocator.destroy(self);
}

/// capacity hint only (pipeline.zig:25-33)
pub fn prealloc(self: *Pipeline, capacity: usize) error{OutOfMemory}!void {
    _ = self;
    _ = capacity;
}

pub fn prevEvents(self: *Pipeline) ?[]const cl.event.Event {
    _ = self;
    return null;
}

pub fn append(self: *Pipeline, events: []const cl.event.Event) error{OutOfMemory}!void {
    _ = self;
    _ = events;
}

/// clWaitForEvents + release (pipeline.zig:47-61) -> cudaStreamSynchronize
pub fn waitAndCleanup(self: *Pipeline) void {
    b200.check(b200.wk_queue_finish(self.command_queue.handle())) catch |err| {
        std.debug.panic("Unexpected error ({s}) while waiting for the queue: {s}", .{ @errorName(err), b200.wk_last_error() });
    };
}

pub fn clear(self: *Pipeline) void {
    _ = self;
}

/// the handle every FFI call takes
pub inline fn q(self: *Pipeline) *b200.Queue {
    return self.command_queue.handle();
}

const Pipeline = @This();
