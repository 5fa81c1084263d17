//! b200.zig -- the whole FFI surface of the B200 backend: one `extern fn` per exported symbol of
//! libwekua_b200.so (include/wekua_b200.h), the status -> Zig error mapping, and two marshalling helpers.
//! Every other file of this overlay calls the device ONLY through these declarations
//! (tests/test_zig_shim.py checks name and arity of each one against the C header).
//!
//! Replaces the `opencl` module as the seam of src/core, src/tensor, src/blas, src/math and src/nn
//! (reference: ~35 zig-opencl calls on JIT-compiled OpenCL-C text, src/core/kernel.zig:123-334).

pub const Context = opaque {}; // wk_context
pub const Queue = opaque {}; // wk_queue
pub const Event = opaque {}; // wk_event
pub const Graph = opaque {}; // wk_graph

/// wk_queue_info_t: the fields of CommandQueue that callers read (src/core/command_queue.zig:10-28)
pub const QueueInfo = extern struct {
    device_name: [256]u8,
    device_ordinal: i32,
    wekua_id: i32,
    compute_units: u32,
    max_work_group_size: u64,
    local_mem_size: u64,
    local_mem_type: i32,
    cache_line_size: u32,
    vector_widths: [10]u16,
    global_mem_size: u64,
    cc_major: i32,
    cc_minor: i32,
};

/// wk_opt_param_t: one parameter tensor of a multi-tensor optimizer step
pub const OptParam = extern struct {
    x: ?*anyopaque,
    grad: ?*const anyopaque,
    state0: ?*anyopaque,
    state1: ?*anyopaque,
    n: u64,
};

pub const OK: i32 = 0;
pub const ERR_INVALID_VALUE: i32 = 1;
pub const ERR_INVALID_COORDINATES: i32 = 2;
pub const ERR_INVALID_BUFFER: i32 = 3;
pub const ERR_UNEQUAL_ATTRIBUTE: i32 = 4;
pub const ERR_UNEQUAL_SHAPE: i32 = 5;
pub const ERR_UNEQUAL_DIMENSION: i32 = 6;
pub const ERR_UNEQUAL_CONTEXT: i32 = 7;
pub const ERR_OUT_OF_MEMORY: i32 = 8;
pub const ERR_TYPE_NOT_SUPPORTED: i32 = 9;
pub const ERR_NO_DEVICE: i32 = 10;
pub const ERR_CUDA: i32 = 11;

/// blas.Operation, src/blas/gemm.zig:25-29 (the Zig enum's integer values are passed as they are)
pub const NO_TRANSPOSE: i32 = 0;
pub const TRANSPOSE: i32 = 1;

pub const OP_SIN: i32 = 0;
pub const OP_COS: i32 = 1;
pub const OP_TAN: i32 = 2;
pub const OP_SINH: i32 = 3;
pub const OP_COSH: i32 = 4;
pub const OP_TANH: i32 = 5;
pub const OP_SIGMOID: i32 = 6;

pub const ACT_NONE: i32 = 0;
pub const ACT_SIGMOID: i32 = 1;
pub const ACT_TANH: i32 = 2;

pub const OPT_GD: i32 = 0;
pub const OPT_GDM: i32 = 1;
pub const OPT_ADAGRAD: i32 = 2;
pub const OPT_RMSPROP: i32 = 3;
pub const OPT_ADAM: i32 = 4;

// ---- runtime: context / queue / events / graphs (src/core/context.zig, command_queue.zig, pipeline.zig)
pub extern fn wk_last_error() [*:0]const u8;
pub extern fn wk_version() [*:0]const u8;
pub extern fn wk_launch_count() u64;
pub extern fn wk_device_count(count: *i32) i32;
pub extern fn wk_context_create(device_ordinals: [*]const i32, n: i32, out: *?*Context) i32;
pub extern fn wk_context_create_all(out: *?*Context) i32;
pub extern fn wk_context_destroy(ctx: *Context) i32;
pub extern fn wk_context_num_queues(ctx: *const Context, n: *i32) i32;
pub extern fn wk_context_queue(ctx: *Context, index: i32, out: *?*Queue) i32;
pub extern fn wk_queue_wrap_stream(device_ordinal: i32, cuda_stream: ?*anyopaque, out: *?*Queue) i32;
pub extern fn wk_queue_release(q: *Queue) i32;
pub extern fn wk_queue_info(q: *const Queue, info: *QueueInfo) i32;
pub extern fn wk_queue_finish(q: *Queue) i32;
pub extern fn wk_queue_stream(q: *const Queue, cuda_stream: *?*anyopaque) i32;
pub extern fn wk_event_record(q: *Queue, out: *?*Event) i32;
pub extern fn wk_queue_wait_event(q: *Queue, ev: *Event) i32;
pub extern fn wk_event_wait(ev: *Event) i32;
pub extern fn wk_event_elapsed_ms(start: *Event, end: *Event, ms: *f32) i32;
pub extern fn wk_event_release(ev: *Event) i32;
pub extern fn wk_graph_begin_capture(q: *Queue) i32;
pub extern fn wk_graph_end_capture(q: *Queue, out: *?*Graph) i32;
pub extern fn wk_graph_launch(g: *Graph, q: *Queue) i32;
pub extern fn wk_graph_num_kernels(g: *const Graph, n: *u64) i32;
pub extern fn wk_graph_release(g: *Graph) i32;

// ---- memory (src/tensor/main.zig:238-263, fill.zig:72-95, memory/*.zig)
pub extern fn wk_malloc(q: *Queue, bytes: usize, dptr: *?*anyopaque) i32;
pub extern fn wk_free(q: *Queue, dptr: ?*anyopaque) i32;
pub extern fn wk_host_alloc(bytes: usize, hptr: *?*anyopaque) i32;
pub extern fn wk_host_free(hptr: ?*anyopaque) i32;
pub extern fn wk_memset_zero(q: *Queue, dptr: ?*anyopaque, bytes: usize) i32;
pub extern fn wk_h2d_rect(q: *Queue, dst: ?*anyopaque, dst_row_pitch: usize, dst_slice_pitch: usize, src_host: ?*const anyopaque, width_bytes: usize, height: usize, depth: usize) i32;
pub extern fn wk_d2h_rect(q: *Queue, dst_host: ?*anyopaque, src: ?*const anyopaque, src_row_pitch: usize, src_slice_pitch: usize, width_bytes: usize, height: usize, depth: usize) i32;
pub extern fn wk_d2d(q: *Queue, dst: ?*anyopaque, src: ?*const anyopaque, bytes: usize) i32;
pub extern fn wk_d2d_rect(q: *Queue, dst: ?*anyopaque, dst_row_pitch: usize, dst_slice_pitch: usize, src: ?*const anyopaque, src_row_pitch: usize, src_slice_pitch: usize, width_bytes: usize, height: usize, depth: usize) i32;
pub extern fn wk_put_value(q: *Queue, dptr: ?*anyopaque, byte_offset: usize, host_value: ?*const anyopaque, size: usize) i32;
pub extern fn wk_get_value(q: *Queue, dptr: ?*const anyopaque, byte_offset: usize, host_value: ?*anyopaque, size: usize) i32;

// ---- BLAS (src/blas/gemm.zig:834-874, src/blas/axpy.zig:93-169; scal / dot_reduce are new ops, SURVEY a13)
pub extern fn wk_gemm(q: *Queue, dtype: i32, op_a: i32, op_b: i32, m: u64, n: u64, k: u64, alpha_or_null: ?*const anyopaque, a: ?*const anyopaque, lda: u64, b: ?*const anyopaque, ldb: u64, beta_or_null: ?*const anyopaque, c: ?*anyopaque, ldc: u64) i32;
pub extern fn wk_gemm_bias_act(q: *Queue, dtype: i32, op_a: i32, op_b: i32, m: u64, n: u64, k: u64, a: ?*const anyopaque, lda: u64, b: ?*const anyopaque, ldb: u64, c: ?*anyopaque, ldc: u64, bias_or_null: ?*const anyopaque, activation: i32) i32;
pub extern fn wk_gemm_set_path(path: i32) i32;
pub extern fn wk_axpy(q: *Queue, dtype: i32, depth: u64, rows: u64, cols: u64, alpha_or_null: ?*const anyopaque, x: ?*const anyopaque, x_row_pitch: u64, x_slice_pitch: u64, y: ?*anyopaque, y_row_pitch: u64, y_slice_pitch: u64) i32;
pub extern fn wk_scal(q: *Queue, dtype: i32, depth: u64, rows: u64, cols: u64, alpha: ?*const anyopaque, x: ?*anyopaque, x_row_pitch: u64, x_slice_pitch: u64) i32;
pub extern fn wk_dot_reduce(q: *Queue, dtype: i32, depth: u64, rows: u64, cols: u64, x: ?*const anyopaque, x_row_pitch: u64, x_slice_pitch: u64, y: ?*const anyopaque, y_row_pitch: u64, y_slice_pitch: u64, host_out: ?*anyopaque) i32;
pub extern fn wk_dot_reduce_async(q: *Queue, dtype: i32, depth: u64, rows: u64, cols: u64, x: ?*const anyopaque, x_row_pitch: u64, x_slice_pitch: u64, y: ?*const anyopaque, y_row_pitch: u64, y_slice_pitch: u64, device_out: ?*anyopaque) i32;

// ---- math (src/math/basic.zig:17-240, src/math/trig.zig:15-113)
pub extern fn wk_hadamard(q: *Queue, dtype: i32, depth: u64, rows: u64, cols: u64, x: ?*anyopaque, x_row_pitch: u64, x_slice_pitch: u64, y: ?*const anyopaque, y_row_pitch: u64, y_slice_pitch: u64) i32;
pub extern fn wk_sum(q: *Queue, dtype: i32, depth: u64, rows: u64, row_pitch: u64, slice_pitch: u64, x: ?*const anyopaque, host_out: ?*anyopaque) i32;
pub extern fn wk_sum_async(q: *Queue, dtype: i32, depth: u64, rows: u64, row_pitch: u64, slice_pitch: u64, x: ?*const anyopaque, device_out: ?*anyopaque) i32;
pub extern fn wk_unary(q: *Queue, dtype: i32, op: i32, x: ?*anyopaque, n: u64) i32;

// ---- nn (src/nn/activation/*.zig, src/nn/layer/linear.zig:424-577, src/nn/loss/mse.zig:63-132, src/nn/optimizers/*.zig)
pub extern fn wk_sigmoid_dev(q: *Queue, dtype: i32, output: ?*const anyopaque, derivative: ?*anyopaque, n: u64) i32;
pub extern fn wk_tanh_dev(q: *Queue, dtype: i32, output: ?*const anyopaque, derivative: ?*anyopaque, n: u64) i32;
pub extern fn wk_bias_add(q: *Queue, dtype: i32, output: ?*anyopaque, bias: ?*const anyopaque, row_pitch: u64, n: u64) i32;
pub extern fn wk_bias_step(q: *Queue, dtype: i32, sensitivity: ?*const anyopaque, bias_grad: ?*anyopaque, row_pitch: u64, rows: u64, n_cols: u64) i32;
pub extern fn wk_mse(q: *Queue, dtype: i32, output: ?*const anyopaque, expected: ?*const anyopaque, error_tensor: ?*anyopaque, dev_or_null: ?*anyopaque, n: u64) i32;
pub extern fn wk_act_backward(q: *Queue, dtype: i32, activation: i32, output: ?*const anyopaque, derivative_or_null: ?*anyopaque, sensitivity: ?*anyopaque, n: u64) i32;
pub extern fn wk_linear_backward(q: *Queue, dtype: i32, activation: i32, batch: u64, n_out: u64, n_in: u64, sensitivity: ?*anyopaque, ld_s: u64, output: ?*const anyopaque, ld_o: u64, prev_output: ?*const anyopaque, ld_p: u64, weight: ?*const anyopaque, ld_w: u64, gradient: ?*anyopaque, ld_g: u64, bias_gradient_or_null: ?*anyopaque, next_sensitivity_or_null: ?*anyopaque, ld_n: u64) i32;
pub extern fn wk_linear_backward_set_mode(mode: i32) i32;
pub extern fn wk_gdm(q: *Queue, dtype: i32, x: ?*anyopaque, grad: ?*const anyopaque, velocity: ?*anyopaque, lr: ?*const anyopaque, beta: ?*const anyopaque, n: u64) i32;
pub extern fn wk_adagrad(q: *Queue, dtype: i32, x: ?*anyopaque, grad: ?*const anyopaque, history: ?*anyopaque, lr: ?*const anyopaque, n: u64) i32;
pub extern fn wk_rmsprop(q: *Queue, dtype: i32, x: ?*anyopaque, grad: ?*const anyopaque, history: ?*anyopaque, lr: ?*const anyopaque, gamma: ?*const anyopaque, n: u64) i32;
pub extern fn wk_adam(q: *Queue, dtype: i32, x: ?*anyopaque, grad: ?*const anyopaque, m: ?*anyopaque, v: ?*anyopaque, lr: ?*const anyopaque, beta1: ?*const anyopaque, beta2: ?*const anyopaque, eps: ?*const anyopaque, t: u64, n: u64) i32;
pub extern fn wk_optimizer_step_multi(q: *Queue, dtype: i32, kind: i32, params: [*]const OptParam, n_params: u32, lr: ?*const anyopaque, h0: ?*const anyopaque, h1: ?*const anyopaque, h2: ?*const anyopaque, t: u64) i32;

// ---- tensor utilities (src/tensor/fill.zig:15, identity.zig:16, random/uniform.zig:60, transpose.zig:15)
pub extern fn wk_fill(q: *Queue, dtype: i32, depth: u64, rows: u64, cols: u64, buf: ?*anyopaque, row_pitch: u64, slice_pitch: u64, scalar: ?*const anyopaque) i32;
pub extern fn wk_identity(q: *Queue, dtype: i32, buf: ?*anyopaque, n_total_elements: u64, size: u64, pitch_sum: u64) i32;
pub extern fn wk_uniform(q: *Queue, dtype: i32, depth: u64, rows: u64, cols: u64, buf: ?*anyopaque, row_pitch: u64, slice_pitch: u64, seed: u64, min_or_null: ?*const anyopaque, max_or_null: ?*const anyopaque) i32;
pub extern fn wk_transpose2d(q: *Queue, dtype: i32, rows: u64, cols: u64, src: ?*const anyopaque, src_pitch: u64, dst: ?*anyopaque, dst_pitch: u64) i32;
pub extern fn wk_transpose_nd(q: *Queue, dtype: i32, ndim: u32, src: ?*const anyopaque, src_pitches: [*]const u64, dst: ?*anyopaque, dst_pitches: [*]const u64, row_pitch: u64, slice_pitch: u64, height: u64, cols: u64, n_elements: u64, dim0: u32, dim1: u32) i32;

// ---- multi-GPU (new: the reference has "one queue per device" only, src/core/command_queue.zig:160-181)
pub extern fn wk_gemm_rowshard_allgather(q: *Queue, dtype: i32, op_a: i32, op_b: i32, m_local: u64, n: u64, k: u64, alpha_or_null: ?*const anyopaque, a: ?*const anyopaque, lda: u64, b: ?*const anyopaque, ldb: u64, beta_or_null: ?*const anyopaque, row0: u64, peer_c: [*]const ?*anyopaque, n_peers: i32, rank: i32, ldc: u64) i32;
pub extern fn wk_ipc_get_handle(q: *Queue, dptr: ?*anyopaque, handle64: *[64]u8) i32;
pub extern fn wk_ipc_open_handle(q: *Queue, handle64: *const [64]u8, dptr: *?*anyopaque) i32;
pub extern fn wk_ipc_close_handle(q: *Queue, dptr: ?*anyopaque) i32;
pub extern fn wk_enable_peer_access(q: *Queue, peer_device_ordinal: i32) i32;

/// The error set a status code can turn into: TensorErrors of src/tensor/main.zig:25-33, KernelsSet.Errors
/// (src/core/kernel.zig:9), Context.Errors (src/core/context.zig:6) and one catch-all for CUDA failures
/// (the text is in wk_last_error()).
pub const Error = error{
    InvalidValue,
    InvalidCoordinates,
    InvalidBuffer,
    UnqualTensorsAttribute,
    UnqualTensorsShape,
    UnqualTensorsDimension,
    UnqualTensorsContext,
    OutOfMemory,
    TypeNotSupported,
    DevicesArrayEmpty,
    CudaFailure,
};

/// int32 status -> Zig error union
pub inline fn check(status: i32) Error!void {
    return switch (status) {
        OK => {},
        ERR_INVALID_VALUE => error.InvalidValue,
        ERR_INVALID_COORDINATES => error.InvalidCoordinates,
        ERR_INVALID_BUFFER => error.InvalidBuffer,
        ERR_UNEQUAL_ATTRIBUTE => error.UnqualTensorsAttribute,
        ERR_UNEQUAL_SHAPE => error.UnqualTensorsShape,
        ERR_UNEQUAL_DIMENSION => error.UnqualTensorsDimension,
        ERR_UNEQUAL_CONTEXT => error.UnqualTensorsContext,
        ERR_OUT_OF_MEMORY => error.OutOfMemory,
        ERR_TYPE_NOT_SUPPORTED => error.TypeNotSupported,
        ERR_NO_DEVICE => error.DevicesArrayEmpty,
        else => error.CudaFailure,
    };
}

/// `?T` scalar -> pointer-or-NULL (a Zig `null` alpha / beta is a NULL pointer on the C side:
/// src/blas/gemm.zig:500-501,592-596, src/blas/axpy.zig:103-111)
pub inline fn optPtr(comptime T: type, v: *const ?T) ?*const anyopaque {
    return if (v.*) |*p| @ptrCast(p) else null;
}
