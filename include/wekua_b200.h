/*
 * wekua_b200.h -- C ABI of libwekua_b200.so: the B200 (sm_100a) backend that slots in underneath
 * wekua's Zig `Context / CommandQueue / Pipeline / Tensor` API for the dense-BLAS hot path.
 *
 * The reference (kython28/wekua) has no FFI boundary of its own: every device interaction goes through
 * the `opencl` Zig module (clSetKernelArg + clEnqueueNDRangeKernel on JIT-compiled OpenCL-C text).  This
 * header is the replacement seam one level up -- ONE extern "C" function per Zig operator.  Each entry
 * cites the reference interface (file:line under /root/reference) whose body it replaces; the Zig stubs
 * that bind them are in INTEGRATION.md and zig/.
 *
 * Conventions
 *  - every function returns an int32 status (WK_OK == 0); the Zig shim maps it onto the TensorErrors
 *    error set of src/tensor/main.zig:25-33.
 *  - `dtype` is core.types.getTypeIndex (src/core/types.zig:60-87): 0 i8, 1 u8, 2 i16, 3 u16, 4 i32, 5 u32,
 *    6 i64, 7 u64, 8 f32, 9 f64; 10-19 = Complex(T) of 0-9 (types.zig:74-83: {real, imag} stored next to each other).
 *    Complex is accepted wherever the reference compiles a complex branch (gemm, axpy, dot, sum, trig, fill, identity,
 *    uniform, transpose); the nn kernels are f32/f64 only like the reference (TypeNotSupported otherwise).
 *  - scalars (alpha, beta, lr ...) are passed by pointer to ONE host value of the element type; a NULL
 *    pointer is Zig's `null` (src/blas/gemm.zig:834-874: beta == null overwrites C, alpha == null means 1).
 *  - pitches / leading dimensions are in ELEMENTS, exactly the numbers the reference passes to its kernels
 *    (memory_layout.row_pitch / slice_pitch, src/tensor/main.zig:49-56).
 *  - every op is asynchronous on the queue's CUDA stream (the reference's `prevEvents -> enqueue -> append`
 *    chain of src/core/pipeline.zig:35-45 is one in-order stream); results are visible after
 *    wk_queue_finish (== Pipeline.waitAndCleanup, pipeline.zig:47-61).  wk_sum / wk_dot_reduce /
 *    wk_get_value block, like the reference's mapped read in src/math/basic.zig:154-171.
 *  - one host thread per queue, like the reference.
 *  - there is NO CPU path: without a CUDA device every compute entry returns WK_ERR_NO_DEVICE.
 */
#ifndef WEKUA_B200_H
#define WEKUA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ status */
enum {
    WK_OK = 0,
    WK_ERR_INVALID_VALUE = 1,        /* TensorErrors.InvalidValue */
    WK_ERR_INVALID_COORDINATES = 2,  /* TensorErrors.InvalidCoordinates */
    WK_ERR_INVALID_BUFFER = 3,       /* TensorErrors.InvalidBuffer */
    WK_ERR_UNEQUAL_ATTRIBUTE = 4,    /* TensorErrors.UnqualTensorsAttribute */
    WK_ERR_UNEQUAL_SHAPE = 5,        /* TensorErrors.UnqualTensorsShape */
    WK_ERR_UNEQUAL_DIMENSION = 6,    /* TensorErrors.UnqualTensorsDimension */
    WK_ERR_UNEQUAL_CONTEXT = 7,      /* TensorErrors.UnqualTensorsContext */
    WK_ERR_OUT_OF_MEMORY = 8,        /* error.OutOfMemory / CL_MEM_OBJECT_ALLOCATION_FAILURE */
    WK_ERR_TYPE_NOT_SUPPORTED = 9,   /* KernelsSet.Errors.TypeNotSupported (src/core/kernel.zig:9) */
    WK_ERR_NO_DEVICE = 10,           /* Context.init: DevicesArrayEmpty (src/core/context.zig:13-25) */
    WK_ERR_CUDA = 11                 /* any other CUDA failure; text in wk_last_error() */
};

enum { WK_NO_TRANSPOSE = 0, WK_TRANSPOSE = 1 }; /* blas.Operation, src/blas/gemm.zig:25-29 */

enum { /* unary math / activation ops: src/math/trig.zig:67-113, src/nn/activation/sigmoid.zig:39 */
    WK_OP_SIN = 0, WK_OP_COS = 1, WK_OP_TAN = 2, WK_OP_SINH = 3, WK_OP_COSH = 4, WK_OP_TANH = 5, WK_OP_SIGMOID = 6
};

enum { WK_ACT_NONE = 0, WK_ACT_SIGMOID = 1, WK_ACT_TANH = 2 };

/* human-readable text of the last failure on the calling thread */
const char *wk_last_error(void);
const char *wk_version(void);

/* ------------------------------------------------------------------------------- context and queue */
typedef struct wk_context wk_context; /* replaces core.Context       src/core/context.zig:13-190 */
typedef struct wk_queue wk_queue;     /* replaces core.CommandQueue  src/core/command_queue.zig:10-229 */
typedef struct wk_event wk_event;     /* replaces cl_event in Pipeline, src/core/pipeline.zig:6-66 */

/* the fields of CommandQueue that callers read (command_queue.zig:10-28) */
typedef struct {
    char device_name[256];
    int32_t device_ordinal;
    int32_t wekua_id;            /* index of the queue inside its context */
    uint32_t compute_units;      /* SM count */
    uint64_t max_work_group_size;
    uint64_t local_mem_size;     /* opt-in shared memory per block */
    int32_t local_mem_type;      /* 1 == cl.device.LocalMemType.local */
    uint32_t cache_line_size;    /* 128 */
    uint16_t vector_widths[10];  /* all 1: tensors get the minimal (even) padding, see DESIGN.md */
    uint64_t global_mem_size;
    int32_t cc_major, cc_minor;
} wk_queue_info_t;

int32_t wk_device_count(int32_t *count);                                       /* cl.device.getIds */
int32_t wk_context_create(const int32_t *device_ordinals, int32_t n, wk_context **out); /* Context.init :13 */
int32_t wk_context_create_all(wk_context **out);                               /* Context.initFromDeviceType :27 */
int32_t wk_context_destroy(wk_context *ctx);                                   /* Context.deinit :180 */
int32_t wk_context_num_queues(const wk_context *ctx, int32_t *n);
int32_t wk_context_queue(wk_context *ctx, int32_t index, wk_queue **out);      /* context.command_queues[i] */
/* adopt an existing CUDA stream (e.g. torch's current stream) as a stand-alone queue */
int32_t wk_queue_wrap_stream(int32_t device_ordinal, void *cuda_stream, wk_queue **out);
int32_t wk_queue_release(wk_queue *q); /* only for queues made by wk_queue_wrap_stream */
int32_t wk_queue_info(const wk_queue *q, wk_queue_info_t *info);               /* CommandQueue.get_device_info :30-123 */
int32_t wk_queue_finish(wk_queue *q);                                          /* cl.command_queue.finish / waitAndCleanup */
int32_t wk_queue_stream(const wk_queue *q, void **cuda_stream);

int32_t wk_event_record(wk_queue *q, wk_event **out);      /* the event an enqueue returns (pipeline.append) */
int32_t wk_queue_wait_event(wk_queue *q, wk_event *ev);    /* wait-list of an enqueue (pipeline.prevEvents) */
int32_t wk_event_wait(wk_event *ev);                       /* cl.event.wait */
int32_t wk_event_elapsed_ms(wk_event *start, wk_event *end, float *ms);
int32_t wk_event_release(wk_event *ev);                    /* cl.event.release */

/* CUDA-graph capture of a launch-bound op sequence (one Linear step is ~25 microsecond kernels; the
 * reference pays one clEnqueueNDRangeKernel per op, src/core/pipeline.zig:35-45).  Ops enqueued on `q` between begin
 * and end are recorded, not executed; wk_graph_launch replays them as one launch.  Blocking entries (wk_sum,
 * wk_dot_reduce, wk_get_value, wk_free) must not be called while capturing. */
typedef struct wk_graph wk_graph;
int32_t wk_graph_begin_capture(wk_queue *q);
int32_t wk_graph_end_capture(wk_queue *q, wk_graph **out);
int32_t wk_graph_launch(wk_graph *g, wk_queue *q);
int32_t wk_graph_num_kernels(const wk_graph *g, uint64_t *n);
int32_t wk_graph_release(wk_graph *g);

/* ------------------------------------------------------------------------------------------ memory */
int32_t wk_malloc(wk_queue *q, size_t bytes, void **dptr);  /* cl.buffer.create, src/tensor/main.zig:241-247 */
int32_t wk_free(wk_queue *q, void *dptr);                   /* cl.buffer.release, main.zig:258 */
int32_t wk_host_alloc(size_t bytes, void **hptr);           /* pinned staging for async rect copies */
int32_t wk_host_free(void *hptr);
int32_t wk_memset_zero(wk_queue *q, void *dptr, size_t bytes); /* fill.zeroes, src/tensor/fill.zig:72-95 */
/* memory.readFromBuffer (host -> tensor) src/tensor/memory/read_from_buffer.zig:13-63: rect copy of
 * width_bytes x height x depth from a dense host buffer into a pitched device buffer. Pitches in BYTES. */
int32_t wk_h2d_rect(wk_queue *q, void *dst, size_t dst_row_pitch, size_t dst_slice_pitch, const void *src_host,
                    size_t width_bytes, size_t height, size_t depth);
/* memory.writeToBuffer (tensor -> host) src/tensor/memory/write_to_buffer.zig:13-63 */
int32_t wk_d2h_rect(wk_queue *q, void *dst_host, const void *src, size_t src_row_pitch, size_t src_slice_pitch,
                    size_t width_bytes, size_t height, size_t depth);
/* memory.copy src/tensor/memory/copy.zig:85-98 */
int32_t wk_d2d(wk_queue *q, void *dst, const void *src, size_t bytes);
int32_t wk_d2d_rect(wk_queue *q, void *dst, size_t dst_row_pitch, size_t dst_slice_pitch, const void *src,
                    size_t src_row_pitch, size_t src_slice_pitch, size_t width_bytes, size_t height, size_t depth);
/* memory.putValue / getValue (src/tensor/memory/put_value.zig, get_value.zig); get blocks */
int32_t wk_put_value(wk_queue *q, void *dptr, size_t byte_offset, const void *host_value, size_t size);
int32_t wk_get_value(wk_queue *q, const void *dptr, size_t byte_offset, void *host_value, size_t size);

/* ---------------------------------------------------------------------------------------- BLAS ★★ */
/* blas.gemm  src/blas/gemm.zig:834-874 (+ the 7 kernel texts src/blas/kernels/gemm_*.cl).
 * Row-major C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C.  A is stored [M,K] (op_a = N) or [K,M]
 * (op_a = T); B is stored [K,N] or [N,K].  lda/ldb/ldc = row pitch in elements.  Only the logical
 * M x N region of C is written.  f32 runs as 3xTF32 on tcgen05/TMEM, f64 on FP64 tensor-core MMA,
 * integers on a SIMT kernel with the reference's wrap-around (mod 2^bits) semantics. */
int32_t wk_gemm(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K,
                const void *alpha_or_null, const void *A, uint64_t lda, const void *B, uint64_t ldb,
                const void *beta_or_null, void *C, uint64_t ldc);
/* Linear.forward fast path (src/nn/layer/linear.zig:480-525): gemm(N,T) + bias + activation in one epilogue */
int32_t wk_gemm_bias_act(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N,
                         uint64_t K, const void *A, uint64_t lda, const void *B, uint64_t ldb, void *C,
                         uint64_t ldc, const void *bias_or_null, int32_t activation);
/* force a GEMM implementation for tests/benchmarks: 0 auto, 1 SIMT, 2 tensor-core */
int32_t wk_gemm_set_path(int32_t path);

/* blas.axpy  src/blas/axpy.zig:93-169 + kernels/axpy.cl : y += alpha*x over [depth, rows, cols] with
 * per-tensor row/slice pitches.  alpha NULL -> y += x; alpha == -1 -> y -= x (SUBSTRACT variant). */
int32_t wk_axpy(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *alpha_or_null,
                const void *x, uint64_t x_row_pitch, uint64_t x_slice_pitch, void *y, uint64_t y_row_pitch,
                uint64_t y_slice_pitch);
/* new ops named by the north star but absent from src/ (legacy semantics old_kernels/scal.cl:3-20):
 * x *= alpha over the logical region;  dot = sum_i x_i*y_i (blocking, result to host) */
int32_t wk_scal(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *alpha, void *x,
                uint64_t x_row_pitch, uint64_t x_slice_pitch);
int32_t wk_dot_reduce(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *x,
                      uint64_t x_row_pitch, uint64_t x_slice_pitch, const void *y, uint64_t y_row_pitch,
                      uint64_t y_slice_pitch, void *host_out);

/* ----------------------------------------------------------------------------------------- math ★ */
/* math.dot  src/math/basic.zig:17-76 + kernels/dot.cl : x *= y (Hadamard, in place) */
int32_t wk_hadamard(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, void *x,
                    uint64_t x_row_pitch, uint64_t x_slice_pitch, const void *y, uint64_t y_row_pitch,
                    uint64_t y_slice_pitch);
/* math.sum  src/math/basic.zig:131-203 + kernels/sum.cl : sum over i<depth, j<rows, k<row_pitch (the
 * reference's row loop runs over the PADDED row, sum.cl:33-35).  Blocking; result to *host_out. */
int32_t wk_sum(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t row_pitch, uint64_t slice_pitch,
               const void *x, void *host_out);
/* the same reductions with the scalar left in DEVICE memory (one element of `dtype` at device_out) and no host
 * synchronisation: lets a loss value feed later kernels / be read back once per epoch instead of once per step
 * (the reference blocks on a mapped read inside every math.sum, basic.zig:154-171). */
int32_t wk_sum_async(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t row_pitch, uint64_t slice_pitch,
                     const void *x, void *device_out);
int32_t wk_dot_reduce_async(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *x,
                            uint64_t x_row_pitch, uint64_t x_slice_pitch, const void *y, uint64_t y_row_pitch,
                            uint64_t y_slice_pitch, void *device_out);
/* sin|cos|tan|sinh|cosh|tanh (src/math/trig.zig:15-113) and Sigmoid.run (src/nn/activation/sigmoid.zig:39-96):
 * in place over n contiguous elements (the reference launches over the whole padded buffer). f32/f64. */
int32_t wk_unary(wk_queue *q, int32_t dtype, int32_t op, void *x, uint64_t n);

/* ------------------------------------------------------------------------------------------- nn ★ */
/* Sigmoid.getDerivative sigmoid.zig:98-137 (d = y(1-y)); Tanh.getDerivative tanh.zig:45-84 (d = 1-y^2) */
int32_t wk_sigmoid_dev(wk_queue *q, int32_t dtype, const void *output, void *derivative, uint64_t n);
int32_t wk_tanh_dev(wk_queue *q, int32_t dtype, const void *output, void *derivative, uint64_t n);
/* Linear.addBias linear.zig:424-478 + bias.cl : out[i] += bias[i % row_pitch], i < n */
int32_t wk_bias_add(wk_queue *q, int32_t dtype, void *output, const void *bias, uint64_t row_pitch, uint64_t n);
/* Linear.getBiasSensitivity linear.zig:534-577 + bias_step.cl : bias_grad[c] = sum_{r<rows} sens[r*pitch + c], c < n_cols */
int32_t wk_bias_step(wk_queue *q, int32_t dtype, const void *sensitivity, void *bias_grad, uint64_t row_pitch,
                     uint64_t rows, uint64_t n_cols);
/* loss.mse src/nn/loss/mse.zig:63-132 + mse.cl : err = (t-o)^2 ; dev = -2(t-o) when dev_or_null != NULL */
int32_t wk_mse(wk_queue *q, int32_t dtype, const void *output, const void *expected, void *error_tensor,
               void *dev_or_null, uint64_t n);
/* fused backward prologue (linear.zig:608-613): sens *= act'(output) in one pass (sigmoid_dev o hadamard) */
int32_t wk_act_backward(wk_queue *q, int32_t dtype, int32_t activation, const void *output, void *derivative_or_null,
                        void *sensitivity, uint64_t n);
/* Linear.backward of one sub-layer, src/nn/layer/linear.zig:579-678, as ONE call: with s = sensitivity o act'(output),
 * gradient[n_out,n_in] = s^T . prev_output, bias_gradient[n_out] = column sums of s (getBiasSensitivity :534-577),
 * next_sensitivity[batch,n_in] = s . weight (skipped when NULL, :648).  f32 layers the tensor-core kernel can address run as
 * THREE launches by default (one pass forms s in place and sums its columns, then the two GEMMs), TWO in mode 2 (s formed
 * inside both GEMMs' converter stage; measured slower); everything else takes the reference's op-by-op
 * sequence.  `sensitivity` is consumed (holds s or its old contents on return).  Pitches in elements. */
int32_t wk_linear_backward(wk_queue *q, int32_t dtype, int32_t activation, uint64_t batch, uint64_t n_out, uint64_t n_in,
                           void *sensitivity, uint64_t ld_s, const void *output, uint64_t ld_o, const void *prev_output,
                           uint64_t ld_p, const void *weight, uint64_t ld_w, void *gradient, uint64_t ld_g,
                           void *bias_gradient_or_null, void *next_sensitivity_or_null, uint64_t ld_n);
/* test / tuning hook: 0 op-by-op, 1 fused element-wise pass + 2 GEMMs (default), 2 converter-stage prologue, -1 default */
int32_t wk_linear_backward_set_mode(int32_t mode);
/* optimizers: gdm.cl:3-33, adagrad.cl:3-48, rmsprop.cl:3-57 (FLT_EPSILON for both f32 and f64) */
int32_t wk_gdm(wk_queue *q, int32_t dtype, void *x, const void *grad, void *velocity, const void *lr, const void *beta,
               uint64_t n);
int32_t wk_adagrad(wk_queue *q, int32_t dtype, void *x, const void *grad, void *history, const void *lr, uint64_t n);
int32_t wk_rmsprop(wk_queue *q, int32_t dtype, void *x, const void *grad, void *history, const void *lr,
                   const void *gamma, uint64_t n);
/* Adam does not exist in src/ (adam.zig is empty); textbook bias-corrected Adam, step t >= 1. parity unpinned */
int32_t wk_adam(wk_queue *q, int32_t dtype, void *x, const void *grad, void *m, void *v, const void *lr,
                const void *beta1, const void *beta2, const void *eps, uint64_t t, uint64_t n);

/* One launch for the whole parameter list of Optimizer.step (gd.zig:55-94, gdm.zig, adagrad.zig, rmsprop.zig:168-202
 * walk cache.slots and enqueue one kernel per weight / bias tensor).  Element-wise results are
 * bit-identical to the per-tensor entries above.  kind GD: x += lr * grad with the axpy rules (lr is the value GD.init
 * stores, i.e. -config.lr; lr == -1 takes the subtract form); state0/state1 are velocity | history | (m, v).
 * n = contiguous elements of each buffer (the reference launches these kernels over the whole padded buffer). */
typedef struct {
    void *x;            /* parameter tensor (weights or bias) */
    const void *grad;   /* its gradient */
    void *state0;       /* GDM velocity / Adagrad, RMSProp history / Adam m; NULL for GD */
    void *state1;       /* Adam v; NULL otherwise */
    uint64_t n;
} wk_opt_param_t;
enum { WK_OPT_GD = 0, WK_OPT_GDM = 1, WK_OPT_ADAGRAD = 2, WK_OPT_RMSPROP = 3, WK_OPT_ADAM = 4 };
int32_t wk_optimizer_step_multi(wk_queue *q, int32_t dtype, int32_t kind, const wk_opt_param_t *params, uint32_t n_params,
                                const void *lr, const void *h0 /* beta | gamma | beta1 */, const void *h1 /* beta2 */,
                                const void *h2 /* eps */, uint64_t t /* Adam step >= 1 */);

/* -------------------------------------------------------------------------------- tensor utilities */
/* fill.constant src/tensor/fill.zig:15-70 + fill.cl (logical region only) */
int32_t wk_fill(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, void *buf, uint64_t row_pitch,
                uint64_t slice_pitch, const void *scalar);
/* identity src/tensor/identity.zig:16-70 + identity.cl : buf[i * sum(pitches)] = 1, i < size (after zeroing n_total elements) */
int32_t wk_identity(wk_queue *q, int32_t dtype, void *buf, uint64_t n_total_elements, uint64_t size, uint64_t pitch_sum);
/* random.uniform src/tensor/random/uniform.zig:60-123 + uniform.cl : counter-based hash of the padded index */
int32_t wk_uniform(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, void *buf,
                   uint64_t row_pitch, uint64_t slice_pitch, uint64_t seed, const void *min_or_null,
                   const void *max_or_null);
/* transpose (2-D case of src/tensor/transpose.zig:15-113): dst[c, r] = src[r, c] */
int32_t wk_transpose2d(wk_queue *q, int32_t dtype, uint64_t rows, uint64_t cols, const void *src, uint64_t src_pitch,
                       void *dst, uint64_t dst_pitch);

/* transpose (src/tensor/transpose.zig:15-113 + transpose.cl:3-42), any rank <= 8: swaps dim0 and dim1 through the two
 * tensors' pitch arrays.  row_pitch / slice_pitch / cols of the SOURCE, height = rows * row_pitch, n_elements = the
 * source's padded element count -- exactly the arguments transpose.zig:68-86 sets. */
int32_t wk_transpose_nd(wk_queue *q, int32_t dtype, uint32_t ndim, const void *src, const uint64_t *src_pitches, void *dst,
                        const uint64_t *dst_pitches, uint64_t row_pitch, uint64_t slice_pitch, uint64_t height, uint64_t cols,
                        uint64_t n_elements, uint32_t dim0, uint32_t dim1);

/* --------------------------------------------------------------------------------------- multi-GPU */
/* Row-sharded GEMM with the all-gather of C fused into the epilogue (new; the reference's multi-device
 * mode is "one queue per device", src/core/command_queue.zig:160-181).  This rank computes rows
 * [row0, row0+M_local) of C = alpha*op(A)*op(B)+beta*C and stores each finished tile into its own C and
 * into every peer's C (peer_C[i], device pointers mapped through CUDA IPC / peer access; entry `rank`
 * is this rank's own buffer).  A points at this rank's row block of op(A). */
int32_t wk_gemm_rowshard_allgather(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M_local, uint64_t N,
                                   uint64_t K, const void *alpha_or_null, const void *A, uint64_t lda, const void *B,
                                   uint64_t ldb, const void *beta_or_null, uint64_t row0, void *const *peer_C,
                                   int32_t n_peers, int32_t rank, uint64_t ldc);
/* CUDA IPC plumbing for the peer pointers above (handles are 64 bytes) */
int32_t wk_ipc_get_handle(wk_queue *q, void *dptr, void *handle64);
int32_t wk_ipc_open_handle(wk_queue *q, const void *handle64, void **dptr);
int32_t wk_ipc_close_handle(wk_queue *q, void *dptr);
int32_t wk_enable_peer_access(wk_queue *q, int32_t peer_device_ordinal);

/* number of kernel launches issued by this library since process start (bench.py's gpu_launches) */
uint64_t wk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* WEKUA_B200_H */
