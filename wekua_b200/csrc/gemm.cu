// gemm.cu -- blas.gemm entry points (src/blas/gemm.zig:834-874) and back-end selection.
//
//   f32  -> 3xTF32 on tcgen05 tensor cores with TMEM accumulators, TMA-fed   (gemm_f32_tc.cu)
//   f64  -> FP64 tensor-core MMA (DMMA)                                      (gemm_f64_tc.cu)
//   ints -> tcgen05.mma kind::i8 over byte planes, exact mod 2^bits          (gemm_i8_tc.cu)
//           small problems: SIMT, wrap-around arithmetic, bit-exact          (gemm_simt.cu)
//   launch-bound float problems: small-tile SIMT kernel below the measured cross-over             (gemm_simt.cu)
// A float operand whose layout cannot feed the tensor-core loaders (pointer / pitch not 16-byte aligned) is first copied
// into an aligned scratch (stage_aligned).  There is no CPU path.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

// cross-over of the small-tile SIMT kernel against the tensor-core kernels, in multiply-adds (measured: profiles/small_gemm_f32_r02.txt)
#ifndef WK_SMALL_SIMT_F32
#define WK_SMALL_SIMT_F32 7.0e7  /* ~400^3: 256^3 7.9 us against 15.7 on the tensor cores; twice that (512^3: 14.9 us against 16.5) where the
                                    mid-tile cp.async kernel can take the operands (profiles/small_gemm_r02c.txt) */
#endif
#ifndef WK_SMALL_SIMT_F64
#define WK_SMALL_SIMT_F64 1.4e8  /* ~512^3: 256^3 12.7 us against 42.5 on DMMA (four 128 x 128 tiles), 768^3 101 against 49 */
#endif

namespace wk {
static int g_gemm_path = 0;  // 0 auto, 1 SIMT, 2 tensor-core (fail if not eligible)

static int g_linear_backward_mode = -1;  // -1: environment / default
static int linear_backward_mode() {
    static const int env = [] { const char *e = getenv("WK_LINEAR_BACKWARD_FUSED"); return e && *e ? atoi(e) : 1; }();
    return g_linear_backward_mode >= 0 ? g_linear_backward_mode : env;
}

static int gemm_path() {
    static int env = -1;
    if (env < 0) {
        const char *e = getenv("WK_GEMM_PATH");
        env = 0;
        if (e && (e[0] == 's' || e[0] == '1')) env = 1;
        if (e && (e[0] == 't' || e[0] == '2')) env = 2;
    }
    return g_gemm_path ? g_gemm_path : env;
}

// Float operands the tensor-core loaders cannot address -- TMA (f32) and 16-byte cp.async (f64) need 16-byte aligned bases
// and row pitches, and the reference's layout law with vector width 1 gives an f32 tensor of `cols % 4 == 2` columns a
// pitch of cols floats (src/tensor/main.zig:174-187) -- are first copied into an aligned scratch: one pitched
// device-to-device copy per operand (2 N^2 bytes against N^3 flops; measured at N = 4098: 37.5 -> see DESIGN.md).
static int32_t stage_aligned(wk_queue *q, size_t es, const void **ptr, uint64_t *ld, uint64_t rows, uint64_t cols, size_t *offset) {
    const uint64_t per16 = 16 / es;
    if (aligned16(*ptr) && *ld % per16 == 0) return WK_OK;
    const uint64_t new_ld = (cols + per16 - 1) / per16 * per16;
    char *dst = (char *)q->align_ws + *offset;
    WK_CUDA(cudaMemcpy2DAsync(dst, new_ld * es, *ptr, *ld * es, cols * es, rows, cudaMemcpyDeviceToDevice, q->stream));
    *offset += ((size_t)rows * new_ld * es + 255) / 256 * 256;
    *ptr = dst;
    *ld = new_ld;
    return WK_OK;
}

static int32_t gemm_any(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K,
                        const void *alpha, const void *A, uint64_t lda, const void *B, uint64_t ldb, const void *beta, void *C,
                        uint64_t ldc, const void *bias, int32_t act, const GemmPeers *peers) {
    if (!A || !B || !C) {
        set_error("gemm: null buffer");
        return WK_ERR_INVALID_BUFFER;
    }
    if (M == 0 || N == 0 || K == 0 || (op_a & ~1) || (op_b & ~1)) {
        set_error("gemm: invalid shape/op");
        return WK_ERR_INVALID_VALUE;
    }
    const uint64_t a_cols = op_a ? M : K, b_cols = op_b ? K : N;
    if (lda < a_cols || ldb < b_cols || ldc < N) {
        set_error("gemm: pitch smaller than row length");
        return WK_ERR_INVALID_VALUE;
    }
    const int path = gemm_path();
    // Launch-bound float problems: the tensor-core kernels carry ~15 us of fixed cost (cluster launch, TMEM, pipeline fill, split-K
    // exchange), the small-tile SIMT kernel ~3 us -- below the measured cross-over it is simply faster, and exact fp32 / fp64 FMA
    // arithmetic on top (thresholds in multiply-adds; WK_GEMM_SMALL_SIMT_F32 / _F64 override, 0 disables)
    static const double small_f32 = [] { const char *e = getenv("WK_GEMM_SMALL_SIMT_F32"); return e && *e ? atof(e) : WK_SMALL_SIMT_F32; }();
    static const double small_f64 = [] { const char *e = getenv("WK_GEMM_SMALL_SIMT_F64"); return e && *e ? atof(e) : WK_SMALL_SIMT_F64; }();
    const double macs = (double)M * (double)N * (double)K;
    // (the SIMT kernel has no split-K: long-K / few-tile shapes such as 64 x 512 x 4096 stay on the tensor cores)
    const bool small_float = path == 0 && !(peers && peers->n > 1) &&
                             ((dtype == 8 && K <= 512 &&
                               (macs <= small_f32 || (macs <= 2 * small_f32 && gemm_simt_mid_ok(q, dtype, op_a, op_b, M, N, K, A, lda, B, ldb)))) ||
                              (dtype == 9 && macs <= small_f64 && K <= 1024));
    if (path != 1 && !small_float && (dtype == 8 || dtype == 9)) {
        const size_t es = dtype_size(dtype), per16 = 16 / es;
        const uint64_t a_rows = op_a ? K : M, b_rows = op_b ? N : K;
        const bool a_bad = !aligned16(A) || lda % per16, b_bad = !aligned16(B) || ldb % per16;
        static const int stage_env = [] { const char *e = getenv("WK_GEMM_STAGE_UNALIGNED"); return e && *e ? atoi(e) : 1; }();
        // (tiny problems -- the XOR network's 4 x 10 x 2 -- are launch-bound either way and keep the one-launch SIMT kernel)
        if ((a_bad || b_bad) && stage_env && (double)M * (double)N * (double)K >= (double)(1 << 21)) {
            const size_t need = (a_bad ? ((size_t)a_rows * (a_cols + per16) * es + 256) : 0) + (b_bad ? ((size_t)b_rows * (b_cols + per16) * es + 256) : 0);
            if (q->align_ws_bytes < need) {
                int32_t rc = grow_buffer(q, &q->align_ws, &q->align_ws_bytes, need);
                if (rc != WK_OK) return rc;
            }
            size_t off = 0;
            int32_t rc = stage_aligned(q, es, &A, &lda, a_rows, a_cols, &off);
            if (rc == WK_OK) rc = stage_aligned(q, es, &B, &ldb, b_rows, b_cols, &off);
            if (rc != WK_OK) return rc;
        }
        int32_t rc = dtype == 8
                         ? gemm_f32_tc(q, op_a, op_b, M, N, K, (const float *)alpha, (const float *)A, lda, (const float *)B, ldb,
                                       (const float *)beta, (float *)C, ldc, (const float *)bias, act, peers)
                         : gemm_f64_tc(q, op_a, op_b, M, N, K, (const double *)alpha, (const double *)A, lda, (const double *)B,
                                       ldb, (const double *)beta, (double *)C, ldc, (const double *)bias, act, peers);
        if (rc != -1) return rc;
        if (path == 2) {
            set_error("gemm: tensor-core path forced but problem not eligible (alignment)");
            return WK_ERR_INVALID_VALUE;
        }
    }
    if (path != 1 && dtype >= 0 && dtype <= 7 && !bias && act == WK_ACT_NONE && !(peers && peers->n > 1)) {
        // The byte-plane passes cost W staging/GEMM launches with a fixed ~20 us each: below ~1 G multiply-adds the one-launch
        // SIMT kernel wins.  WK_GEMM_INT_TC=0 never, =2 always (also what wk_gemm_set_path(2) does).
        static const int int_tc_env = [] { const char *e = getenv("WK_GEMM_INT_TC"); return e && *e ? atoi(e) : 1; }();
        const double macs = (double)M * (double)N * (double)K;
        if (int_tc_env != 0 && (path == 2 || int_tc_env == 2 || macs >= (double)(1ull << 29))) {
            int32_t rc = gemm_int_tc(q, dtype, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
            if (rc != -1) return rc;
            if (path == 2) {
                set_error("gemm: tensor-core path forced but problem not eligible");
                return WK_ERR_INVALID_VALUE;
            }
        }
    }
    int32_t rc = gemm_simt(q, dtype, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act);
    if (rc != WK_OK || !peers || peers->n <= 1) return rc;
    // operands the tensor-core kernels cannot address (e.g. a transposed A block starting at a row that is not a multiple of
    // 16 bytes): the SIMT kernel has no fused epilogue, so the finished block follows to the peers as pitched copies
    const size_t es = dtype_size(dtype);
    for (int i = 0; i < peers->n; i++) {
        if (i == peers->self) continue;
        WK_CUDA(cudaMemcpy2DAsync(peers->ptrs[i], ldc * es, C, ldc * es, N * es, M, cudaMemcpyDeviceToDevice, q->stream));
    }
    return WK_OK;
}
// Complex(T) GEMM (ids 10-19; the reference's complex branches: gemm_2x2.cl:121-236, gemm_nxn.cl:383-398) as ONE real
// GEMM of shape M x 2N x 2K on the real back-ends -- derivation and operand layouts in complex.cu.
//   alpha: imag == 0 -> the real epilogue's alpha;  otherwise folded into the expansion pass of B
//   beta : imag == 0 -> the real epilogue's beta;   otherwise C *= beta first (one streaming pass), then beta' = 1
static bool imag_is_zero(int32_t base, const void *cx) {
    const size_t s = real_dtype_size(base);
    const unsigned char *p = (const unsigned char *)cx + s;
    if (base == 8) return load_host<float>(p) == 0.0f;   // -0.0 counts as zero
    if (base == 9) return load_host<double>(p) == 0.0;
    for (size_t i = 0; i < s; i++)
        if (p[i]) return false;
    return true;
}

static int32_t gemm_complex(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K,
                            const void *alpha, const void *A, uint64_t lda, const void *B, uint64_t ldb, const void *beta, void *C,
                            uint64_t ldc) {
    if (!A || !B || !C) {
        set_error("gemm: null buffer");
        return WK_ERR_INVALID_BUFFER;
    }
    if (M == 0 || N == 0 || K == 0 || (op_a & ~1) || (op_b & ~1)) {
        set_error("gemm: invalid shape/op");
        return WK_ERR_INVALID_VALUE;
    }
    const uint64_t a_cols = op_a ? M : K, b_cols = op_b ? K : N, a_rows = op_a ? K : M, b_rows = op_b ? N : K;
    if (lda < a_cols || ldb < b_cols || ldc < N) {
        set_error("gemm: pitch smaller than row length");
        return WK_ERR_INVALID_VALUE;
    }
    const int32_t base = dtype - 10;
    const size_t es = real_dtype_size(base);
    const uint64_t align_el = 16 / es;  // workspace pitches keep every row 16-byte aligned (TMA / cp.async)
    auto round_up = [&](uint64_t v) { return (v + align_el - 1) / align_el * align_el; };
    const uint64_t ldb2 = round_up(2 * b_cols);                  // expanded B: [2*b_rows, 2*b_cols] reals
    const uint64_t lda2 = op_a ? round_up(a_cols) : 2 * lda;     // split A: [2*K, M] reals; op_a = N: storage in place
    const size_t b_bytes = (size_t)(2 * b_rows) * ldb2 * es;
    const size_t b_bytes_al = (b_bytes + 255) / 256 * 256;
    const size_t a_bytes = op_a ? (size_t)(2 * a_rows) * lda2 * es : 0;
    int32_t rc = ensure_workspace(q, b_bytes_al + a_bytes);
    if (rc != WK_OK) return rc;
    void *B2 = q->ws;
    void *A2 = op_a ? (void *)((char *)q->ws + b_bytes_al) : const_cast<void *>(A);

    const bool fold_alpha = alpha && !imag_is_zero(base, alpha);
    rc = cx_expand_b(q, base, op_b, b_rows, b_cols, B, ldb, B2, ldb2, fold_alpha ? alpha : nullptr);
    if (rc != WK_OK) return rc;
    if (op_a) {
        rc = cx_split_a(q, base, a_rows, a_cols, A, lda, A2, lda2);
        if (rc != WK_OK) return rc;
    }
    // real scalars for the epilogue: one host element of the base type (the .re component when imag == 0)
    unsigned char one[8] = {0};
    if (base == 8) { const float f = 1.0f; memcpy(one, &f, 4); }
    else if (base == 9) { const double d = 1.0; memcpy(one, &d, 8); }
    else one[0] = 1;  // little-endian integer 1 of any width
    const void *r_alpha = alpha ? (fold_alpha ? (const void *)one : alpha) : nullptr;
    const void *r_beta = beta;
    if (beta && !imag_is_zero(base, beta)) {
        rc = wk_scal(q, dtype, 1, M, N, beta, C, ldc, ldc * M);
        if (rc != WK_OK) return rc;
        r_beta = one;
    }
    return gemm_any(q, base, op_a, op_b, M, 2 * N, 2 * K, r_alpha, A2, lda2, B2, ldb2, r_beta, C, 2 * ldc, nullptr, WK_ACT_NONE,
                    nullptr);
}
}  // namespace wk

using namespace wk;

WK_API int32_t wk_gemm_set_path(int32_t path) {
    if (path < 0 || path > 2) return WK_ERR_INVALID_VALUE;
    g_gemm_path = path;
    return WK_OK;
}

WK_API int32_t wk_gemm(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K,
                       const void *alpha, const void *A, uint64_t lda, const void *B, uint64_t ldb, const void *beta, void *C,
                       uint64_t ldc) {
    WK_CHECK_QUEUE(q);
    if (dtype < 0 || dtype > 19) {
        set_error("gemm: dtype %d not supported", dtype);
        return WK_ERR_TYPE_NOT_SUPPORTED;
    }
    if (dtype >= 10) return gemm_complex(q, dtype, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    return gemm_any(q, dtype, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, nullptr, WK_ACT_NONE, nullptr);
}

WK_API int32_t wk_gemm_bias_act(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K,
                                const void *A, uint64_t lda, const void *B, uint64_t ldb, void *C, uint64_t ldc,
                                const void *bias, int32_t act) {
    WK_CHECK_QUEUE(q);
    if (dtype != 8 && dtype != 9) {
        set_error("gemm_bias_act: f32/f64 only");
        return WK_ERR_TYPE_NOT_SUPPORTED;
    }
    if (act < WK_ACT_NONE || act > WK_ACT_TANH) return WK_ERR_INVALID_VALUE;
    // f32: the tensor-core kernel's epilogue runs on four warps per SM.  With several tiles per SM it hides behind the next tile's
    // MMAs and bias + activation ride along for free; with at most ~one tile per CTA pair it is exposed, and 256 expf / tanhf
    // per thread cost more (measured: 20 us at 256 x 4096) than one more streaming pass over a dense C that still sits in L2.
    const uint64_t tiles = ((M + 255) / 256) * ((N + 255) / 256);
    if (dtype == 8 && (bias || act != WK_ACT_NONE) && ldc == N && tiles <= 2 * 74 && gemm_path() != 1) {
        int32_t rc = gemm_any(q, dtype, op_a, op_b, M, N, K, nullptr, A, lda, B, ldb, nullptr, C, ldc, nullptr, WK_ACT_NONE, nullptr);
        if (rc != WK_OK) return rc;
        if (bias && act != WK_ACT_NONE) return bias_act(q, dtype, C, bias, ldc, M * ldc, act);
        if (bias) return wk_bias_add(q, dtype, C, bias, ldc, M * ldc);
        return wk_unary(q, dtype, act == WK_ACT_SIGMOID ? WK_OP_SIGMOID : WK_OP_TANH, C, M * ldc);
    }
    return gemm_any(q, dtype, op_a, op_b, M, N, K, nullptr, A, lda, B, ldb, nullptr, C, ldc, bias, act, nullptr);
}

// Linear.backward of one (sub-)layer, src/nn/layer/linear.zig:579-678:
//   d      = act'(output)                 (sigmoid.cl:22-38 / tanh.cl:3-20: from the layer OUTPUT)
//   s      = sensitivity * d              (math.dot)
//   grad   = s^T . prev_output            (gemm TN)          [n_out, n_in]
//   bgrad  = column sums of s             (bias_step.cl)     [n_out]
//   next   = s . weight                   (gemm NN)          [batch, n_in]      (skipped when next == NULL, linear.zig:648)
// The reference issues 5 launches and moves [batch, n_out] through memory three extra times.  Modes (wk_linear_backward_set_mode
// / WK_LINEAR_BACKWARD_FUSED):
//   1 (default)  THREE launches: one pass writes s over `sensitivity` and sums its columns (the bias gradient) -- act', the
//                Hadamard product and bias_step fused; then the two GEMMs at full speed.
//   2            TWO launches (f32, tensor-core path): both GEMMs read `sensitivity` and `output` and form s in the converter
//                stage; the TN GEMM's converters also add up the columns of s.  Measured SLOWER than mode 1 at every size
//                (B = N = 8192: 16.7 vs 13.3 ms per layer step): the converter warps are the co-bottleneck of the 3xTF32
//                pipeline, so work added there stretches every k-block, and the variant has two stages and no split-K.  Kept
//                for the record and for the parity test; see DESIGN.md.
//   0            the reference's op-by-op sequence.
// `sensitivity` is CONSUMED: on return it holds s (modes 0, 1) or its old contents (mode 2) -- the reference overwrites it
// before it is read again (mse.cl / the next backward GEMM).
WK_API int32_t wk_linear_backward(wk_queue *q, int32_t dtype, int32_t activation, uint64_t batch, uint64_t n_out, uint64_t n_in,
                                  void *sensitivity, uint64_t ld_s, const void *output, uint64_t ld_o, const void *prev_output,
                                  uint64_t ld_p, const void *weight, uint64_t ld_w, void *gradient, uint64_t ld_g,
                                  void *bias_gradient_or_null, void *next_sensitivity_or_null, uint64_t ld_n) {
    WK_CHECK_QUEUE(q);
    if (dtype != 8 && dtype != 9) {
        set_error("linear_backward: f32/f64 only");
        return WK_ERR_TYPE_NOT_SUPPORTED;
    }
    if (!sensitivity || !output || !prev_output || !weight || !gradient) return WK_ERR_INVALID_BUFFER;
    if (activation < WK_ACT_NONE || activation > WK_ACT_TANH || batch == 0 || n_out == 0 || n_in == 0) return WK_ERR_INVALID_VALUE;
    if (ld_s < n_out || ld_o < n_out || ld_p < n_in || ld_w < n_in || ld_g < n_in || (next_sensitivity_or_null && ld_n < n_in))
        return WK_ERR_INVALID_VALUE;
    const int mode = linear_backward_mode();
    const bool big = (double)batch * (double)n_out * (double)n_in >= (double)(1 << 21);  // tiny layers stay launch-bound either way
    if (mode == 2 && dtype == 8 && activation != WK_ACT_NONE && big && gemm_path() != 1 && ld_s == ld_o) {
        GemmProlog pg{(const float *)output, ld_o, activation, (float *)bias_gradient_or_null};
        int32_t rc = gemm_f32_tc(q, 1, 0, n_out, n_in, batch, nullptr, (const float *)sensitivity, ld_s, (const float *)prev_output, ld_p,
                                 nullptr, (float *)gradient, ld_g, nullptr, WK_ACT_NONE, nullptr, &pg);
        if (rc != -1) {
            if (rc != WK_OK || !next_sensitivity_or_null) return rc;
            GemmProlog pn{(const float *)output, ld_o, activation, nullptr};
            rc = gemm_f32_tc(q, 0, 0, batch, n_in, n_out, nullptr, (const float *)sensitivity, ld_s, (const float *)weight, ld_w, nullptr,
                             (float *)next_sensitivity_or_null, ld_n, nullptr, WK_ACT_NONE, nullptr, &pn);
            if (rc != -1) return rc;
            // (cannot happen: the same operands were addressable a moment ago) -- finish op by op from here
            rc = wk_act_backward(q, dtype, activation, output, nullptr, sensitivity, batch * ld_s);
            if (rc != WK_OK) return rc;
            return gemm_any(q, dtype, 0, 0, batch, n_in, n_out, nullptr, sensitivity, ld_s, weight, ld_w, nullptr, next_sensitivity_or_null,
                            ld_n, nullptr, WK_ACT_NONE, nullptr);
        }
    }
    int32_t rc = WK_OK;
    bool bias_done = false;
    if (activation != WK_ACT_NONE) {
        if (mode >= 1 && bias_gradient_or_null) {  // ONE pass: s = sensitivity * act'(output) written back, its columns summed
            rc = act_backward_colsum(q, dtype, activation, output, ld_o, sensitivity, ld_s, batch, n_out, bias_gradient_or_null);
            bias_done = true;
        } else {  // whole rows (pad columns included), like the reference's 1-D launches over the buffer
            rc = wk_act_backward(q, dtype, activation, output, nullptr, sensitivity, batch * ld_s);
        }
        if (rc != WK_OK) return rc;
    }
    rc = gemm_any(q, dtype, 1, 0, n_out, n_in, batch, nullptr, sensitivity, ld_s, prev_output, ld_p, nullptr, gradient, ld_g, nullptr,
                  WK_ACT_NONE, nullptr);
    if (rc != WK_OK) return rc;
    if (bias_gradient_or_null && !bias_done) {
        rc = wk_bias_step(q, dtype, sensitivity, bias_gradient_or_null, ld_s, batch, n_out);
        if (rc != WK_OK) return rc;
    }
    if (!next_sensitivity_or_null) return WK_OK;
    return gemm_any(q, dtype, 0, 0, batch, n_in, n_out, nullptr, sensitivity, ld_s, weight, ld_w, nullptr, next_sensitivity_or_null, ld_n,
                    nullptr, WK_ACT_NONE, nullptr);
}

WK_API int32_t wk_linear_backward_set_mode(int32_t mode) {
    if (mode < -1 || mode > 2) return WK_ERR_INVALID_VALUE;
    g_linear_backward_mode = mode;
    return WK_OK;
}

WK_API int32_t wk_gemm_rowshard_allgather(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M_local, uint64_t N,
                                          uint64_t K, const void *alpha, const void *A, uint64_t lda, const void *B,
                                          uint64_t ldb, const void *beta, uint64_t row0, void *const *peer_C, int32_t n_peers,
                                          int32_t rank, uint64_t ldc) {
    WK_CHECK_QUEUE(q);
    if (!peer_C || n_peers < 1 || n_peers > 16 || rank < 0 || rank >= n_peers) return WK_ERR_INVALID_VALUE;
    if (dtype != 8 && dtype != 9) return WK_ERR_TYPE_NOT_SUPPORTED;
    const size_t es = dtype_size(dtype);
    // every buffer is addressed at this rank's row block
    void *shifted[16];
    for (int i = 0; i < n_peers; i++) {
        if (!peer_C[i]) return WK_ERR_INVALID_BUFFER;
        shifted[i] = (char *)peer_C[i] + row0 * ldc * es;
    }
    GemmPeers peers{shifted, n_peers, rank};
    return gemm_any(q, dtype, op_a, op_b, M_local, N, K, alpha, A, lda, B, ldb, beta, shifted[rank], ldc, nullptr, WK_ACT_NONE,
                    n_peers > 1 ? &peers : nullptr);
}
