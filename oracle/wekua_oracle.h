/*
 * oracle/wekua_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of kython28/wekua's dense-BLAS hot path (host launch logic in
 * src/blas, src/tensor, src/math, src/nn .zig files + the OpenCL-C kernel texts beside them), used ONLY
 * as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (libwekua_b200.so) never links or calls this.
 *
 * Parity pinning: the reference cannot be built here (no zig, no OpenCL device, zig-opencl v0.8.1
 * un-vendored -- SURVEY.md section 8c), so this oracle is pinned against the reference's own
 * known-answer tests transcribed into tests/test_oracle_*.py (gemm A*I / I*B, pack layout, axpy,
 * hadamard/sum/mean, trig, fill/identity/uniform).  The nn kernels (sigmoid, tanh_dev, bias,
 * bias_step, mse, gdm, adagrad, rmsprop), integer GEMM and large random GEMM have NO reference
 * test: for those rows parity is UNPINNED (restatement of the kernel text only).
 *
 * dtype ids follow src/core/types.zig:60-87 (real types only): 0 i8, 1 u8, 2 i16, 3 u16, 4 i32,
 * 5 u32, 6 i64, 7 u64, 8 f32, 9 f64.
 */
#ifndef WEKUA_ORACLE_H
#define WEKUA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WKO_MAX_DIMS 8

enum { WKO_MEM_GLOBAL = 0, WKO_MEM_LOCAL = 1 }; /* cl.device.LocalMemType as used by work_configuration.zig:143 */

/* the subset of src/core/command_queue.zig:10-28 that changes arithmetic or layout */
typedef struct {
    uint64_t vector_widths[10]; /* per type id */
    int32_t local_mem_type;
    uint64_t local_mem_size;
    uint64_t max_work_group_size;
} wko_device;

/* src/tensor/main.zig:41-60 Dimensions + MemoryLayout + the gemm part of WorkConfiguration */
typedef struct {
    uint64_t ndim;
    uint64_t shape[WKO_MAX_DIMS];
    uint64_t vl_shape[WKO_MAX_DIMS];
    uint64_t pitches[WKO_MAX_DIMS];
    uint64_t number_of_elements;
    uint64_t number_of_elements_without_padding;
    uint64_t row_pitch, row_pitch_for_vectors;
    uint64_t slice_pitch, slice_pitch_for_vectors;
    uint64_t number_of_vectors;
    uint64_t depth, rows, rows_padded, cols;
    uint64_t vector_width;
    int32_t vectors_enabled;
    int32_t gemm_algorithm; /* 0..5 = 2x2..64x64, work_configuration.zig:9-16 */
} wko_layout;

/* PoCL-like CPU device (.global local memory, 16 KiB tile rule) and NVIDIA-OpenCL-like GPU device */
void wko_device_cpu(wko_device *d, uint64_t vw_f32);
void wko_device_gpu(wko_device *d);
/* the device our CUDA CommandQueue reports: vector width 1 everywhere, .local memory */
void wko_device_b200(wko_device *d);

/* src/utils/utils.zig:6-32 */
void wko_calculate_work_items(const uint64_t *global, uint64_t *local, uint64_t n, uint64_t max_wg);

/* src/tensor/main.zig:113-251 (layout math) + work_configuration.zig:110-193 (tile choice) */
int32_t wko_layout_init(wko_layout *l, const wko_device *dev, int32_t dtype, const uint64_t *shape, uint64_t ndim,
                        int32_t vectors_enabled_cfg);

/* src/blas/gemm.zig:42-58 */
int32_t wko_get_algorithm(int32_t default_algorithm, uint64_t k_size);

/* src/blas/gemm.zig:117-186 geometry of PackedTensors */
typedef struct {
    wko_layout a, b;
    int32_t algorithm;
    int32_t vectors_enabled;
    uint64_t m_size, n_size, k_size;
} wko_packed_geom;
int32_t wko_packed_init(wko_packed_geom *g, const wko_device *dev, int32_t dtype, uint64_t n_size, uint64_t m_size,
                        uint64_t k_size, int32_t default_algorithm, int32_t vectors_enabled);

/* gemm_pack.cl through PackedTensors.pack (gemm.zig:272-357); dst buffers must be zero-initialised */
int32_t wko_pack(const wko_packed_geom *g, int32_t dtype, const void *a, const wko_layout *la, int32_t op_a,
                 const void *b, const wko_layout *lb, int32_t op_b, void *packed_a, void *packed_b);

/* blas.gemm (gemm.zig:834-874).  alpha/beta: pointer to one element or NULL (Zig `null`).
 * use_packing != 0 follows gemmWithPacking with PackedTensors.init(pipeline, c, K, pack_vectors).
 * returns 0, or -1 InvalidValue. */
int32_t wko_gemm(const wko_device *dev, int32_t dtype, const void *alpha, const void *a, const wko_layout *la,
                 int32_t op_a, const void *b, const wko_layout *lb, int32_t op_b, const void *beta, void *c,
                 const wko_layout *lc, int32_t use_packing, int32_t pack_vectors);

/* blas.axpy (axpy.zig:93-169) */
int32_t wko_axpy(int32_t dtype, const void *x, const wko_layout *lx, const void *alpha, void *y, const wko_layout *ly);
/* math.dot (basic.zig:17-76), math.sum (:131-203), math.mean (:206-240) */
int32_t wko_hadamard(int32_t dtype, void *x, const wko_layout *lx, const void *y, const wko_layout *ly);
int32_t wko_sum(const wko_device *dev, int32_t dtype, const void *x, const wko_layout *lx, void *out);
int32_t wko_mean(const wko_device *dev, int32_t dtype, const void *x, const wko_layout *lx, void *out);

/* 1-D whole-buffer kernels; n = number_of_elements (padding included), f32/f64 only.
 * op: 0 sin 1 cos 2 tan 3 sinh 4 cosh 5 tanh 6 sigmoid */
int32_t wko_unary(int32_t dtype, void *x, uint64_t n, int32_t op);
int32_t wko_sigmoid_dev(int32_t dtype, const void *out, void *dev, uint64_t n);
int32_t wko_tanh_dev(int32_t dtype, const void *in, void *dev, uint64_t n);
int32_t wko_bias(int32_t dtype, void *out, const void *bias, uint64_t row_pitch, uint64_t n);
int32_t wko_bias_step(int32_t dtype, const void *dev, void *bias_grad, uint64_t dev_row_pitch, uint64_t dev_rows,
                      uint64_t n);
int32_t wko_mse(int32_t dtype, const void *out, const void *expected, void *err, void *dev_or_null, uint64_t n);
int32_t wko_gdm(int32_t dtype, void *x, const void *g, void *v, const void *lr, const void *beta, uint64_t n);
int32_t wko_adagrad(int32_t dtype, void *x, const void *g, void *h, const void *lr, uint64_t n);
int32_t wko_rmsprop(int32_t dtype, void *x, const void *g, void *h, const void *lr, const void *gamma, uint64_t n);

/* tensor utilities */
int32_t wko_fill(int32_t dtype, void *buf, const wko_layout *l, const void *scalar);
int32_t wko_identity(int32_t dtype, void *buf, const wko_layout *l);
int32_t wko_transpose(int32_t dtype, const void *a, const wko_layout *la, void *b, const wko_layout *lb, uint64_t dim0,
                      uint64_t dim1);
uint64_t wko_xxhash64(uint64_t index, uint64_t seed); /* uniform.cl:32-54 */
/* random.uniform (uniform.zig:60-123, uniform.cl:56-185); min/max NULL = not given */
int32_t wko_uniform(int32_t dtype, void *buf, const wko_layout *l, uint64_t seed, const void *min_or_null,
                    const void *max_or_null);

/* host <-> padded buffer (memory/read_from_buffer.zig, write_to_buffer.zig) */
int32_t wko_read_from_buffer(int32_t dtype, void *tensor_buf, const wko_layout *l, const void *host);
int32_t wko_write_to_buffer(int32_t dtype, const void *tensor_buf, const wko_layout *l, void *host);

int32_t wko_num_threads(void);
void wko_set_num_threads(int32_t n);

#ifdef __cplusplus
}
#endif
#endif
