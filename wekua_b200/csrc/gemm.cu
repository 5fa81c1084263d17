// gemm.cu -- blas.gemm entry points (src/blas/gemm.zig:834-874) and back-end selection.
//
//   f32  -> 3xTF32 on tcgen05 tensor cores with TMEM accumulators, TMA-fed   (gemm_f32_tc.cu)
//   f64  -> FP64 tensor-core MMA (DMMA)                                      (gemm_f64_tc.cu)
//   ints -> SIMT, wrap-around arithmetic, bit-exact                          (gemm_simt.cu)
// A float problem whose layout cannot feed the tensor-core loaders (pointer / pitch not 16-byte aligned) takes the
// SIMT kernel as well.  There is no CPU path.
#include <stdlib.h>

#include "common.cuh"

namespace wk {
static int g_gemm_path = 0;  // 0 auto, 1 SIMT, 2 tensor-core (fail if not eligible)

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
    if (path != 1 && (dtype == 8 || dtype == 9)) {
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
    if (peers && peers->n > 1) {
        set_error("gemm: fused all-gather epilogue needs the tensor-core path (f32/f64, 16-byte aligned pitches)");
        return WK_ERR_INVALID_VALUE;
    }
    return gemm_simt(q, dtype, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act);
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
    if (dtype < 0 || dtype > 9) {
        set_error("gemm: dtype %d not supported", dtype);
        return WK_ERR_TYPE_NOT_SUPPORTED;
    }
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
    return gemm_any(q, dtype, op_a, op_b, M, N, K, nullptr, A, lda, B, ldb, nullptr, C, ldc, bias, act, nullptr);
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
