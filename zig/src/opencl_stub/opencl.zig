//! Stub of the `opencl` module (kython28/zig-opencl v0.8.1, build.zig.zon:7-10) for the B200 backend.
//!
//! `wekua.opencl` is re-exported (src/wekua.zig:1) and user code names a handful of its declarations
//! (`cl.device.Type.all` in examples/xor_neural_network.zig:25, `cl.device.Type.cpu` in benchmark/*.zig,
//! `cl.buffer.MemFlags` through `CreateTensorConfig`, `cl.errors.OpenCLError` in error sets).  This file keeps exactly
//! those names alive so such code compiles unchanged; nothing here talks to a device.

pub const errors = struct {
    /// kept so `cl.errors.OpenCLError || ...` error-set expressions still type-check; the CUDA backend reports
    /// device failures as `error.CudaFailure` (core/b200.zig)
    pub const OpenCLError = error{
        DeviceNotFound,
        DeviceNotAvailable,
        OutOfResources,
        OutOfHostMemory,
        InvalidValue,
        InvalidDevice,
    };
};

pub const device = struct {
    /// cl_device_type bit field; the CUDA backend treats every selector that includes GPUs as "all CUDA devices"
    pub const Type = enum(u64) {
        default = 1 << 0,
        cpu = 1 << 1,
        gpu = 1 << 2,
        accelerator = 1 << 3,
        custom = 1 << 4,
        all = 0xFFFFFFFF,
    };
    /// CL_LOCAL / CL_GLOBAL (CommandQueue.local_mem_type; B200 shared memory is `.local`)
    pub const LocalMemType = enum(u32) {
        none = 0,
        local = 1,
        global = 2,
    };
    /// a device is a CUDA ordinal
    pub const DeviceId = i32;
};

pub const context = struct {
    /// accepted and ignored (there are no OpenCL platforms to select between)
    pub const Properties = struct {
        platform: ?*anyopaque = null,
    };
};

pub const buffer = struct {
    pub const MemFlags = u64;
    /// CreateConfig.cl_mem_flags (src/tensor/main.zig:35-39); HBM allocations are always read-write
    pub const MemFlag = struct {
        pub const read_write: MemFlags = 1 << 0;
        pub const write_only: MemFlags = 1 << 1;
        pub const read_only: MemFlags = 1 << 2;
        pub const use_host_ptr: MemFlags = 1 << 3;
        pub const alloc_host_ptr: MemFlags = 1 << 4;
        pub const copy_host_ptr: MemFlags = 1 << 5;
    };
    /// a device buffer is a CUDA device pointer
    pub const Mem = ?*anyopaque;
};

pub const event = struct {
    /// Pipeline.prevEvents() keeps its return type; on one in-order CUDA stream the list is always empty
    pub const Event = ?*anyopaque;
};
