// gemm_simt.cu -- generic SIMT GEMM: the integer path of blas.gemm (bit-exact wrap-around semantics, mod 2^bits),
// and the path for float problems whose layout cannot feed TMA (row pitch not a multiple of 16 bytes, tiny shapes).
//
// Replaces the arithmetic of src/blas/kernels/gemm_{2x2,nxn,nxn_gpu}.cl (and the pack + packed variants, which
// compute the same product from a re-tiled copy): C[M,N] = alpha * op(A) * op(B) + beta * C, row-major, K is the
// LOGICAL inner dimension (the reference runs its k-loop to the zero padding, which contributes nothing).
//
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread split into four 4x4 quadrants so shared-memory
// reads are 128-bit and conflict-free; global->register prefetch of the next k-slab overlaps the FMAs.
#include "common.cuh"

namespace wk {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == WK_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
    if (act == WK_ACT_TANH) return tanhf(v);
    return v;
}
__device__ __forceinline__ double act_apply(double v, int act) {
    if (act == WK_ACT_SIGMOID) return wk_sigmoid_f64(v);
    if (act == WK_ACT_TANH) return wk_tanh_f64(v);
    return v;
}
template <typename A> __device__ __forceinline__ A act_apply(A v, int) { return v; }

template <typename T>
__global__ void __launch_bounds__(GT) gemm_simt_kernel(const T *__restrict__ A, const T *__restrict__ B, T *__restrict__ C,
                                                       uint64_t M, uint64_t N, uint64_t K, uint64_t lda, uint64_t ldb,
                                                       uint64_t ldc, int op_a, int op_b, int has_alpha, int has_beta,
                                                       typename Acc<T>::type alpha, typename Acc<T>::type beta,
                                                       const T *__restrict__ bias, int act) {
    using Ac = typename Acc<T>::type;
    __shared__ T As[BK][BM + 4];
    __shared__ T Bs[BK][BN + 4];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const uint64_t m0 = (uint64_t)blockIdx.y * BM, n0 = (uint64_t)blockIdx.x * BN;

    Ac acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = (Ac)0;

    T ra[8], rb[8];
    auto load_slab = [&](uint64_t k0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = t + i * GT;
            int m, k;
            if (op_a == 0) { k = e & (BK - 1); m = e >> 4; }      // A[m][k]: k contiguous
            else { m = e & (BM - 1); k = e >> 7; }                // A[k][m]: m contiguous
            const uint64_t gm = m0 + m, gk = k0 + k;
            T v = (T)0;
            if (gm < M && gk < K) v = op_a == 0 ? A[gm * lda + gk] : A[gk * lda + gm];
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = t + i * GT;
            int n, k;
            if (op_b == 0) { n = e & (BN - 1); k = e >> 7; }      // B[k][n]: n contiguous
            else { k = e & (BK - 1); n = e >> 4; }                // B[n][k]: k contiguous
            const uint64_t gn = n0 + n, gk = k0 + k;
            T v = (T)0;
            if (gn < N && gk < K) v = op_b == 0 ? B[gk * ldb + gn] : B[gn * ldb + gk];
            rb[i] = v;
        }
    };
    auto store_slab = [&]() {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = t + i * GT;
            int m, k;
            if (op_a == 0) { k = e & (BK - 1); m = e >> 4; } else { m = e & (BM - 1); k = e >> 7; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = t + i * GT;
            int n, k;
            if (op_b == 0) { n = e & (BN - 1); k = e >> 7; } else { k = e & (BK - 1); n = e >> 4; }
            Bs[k][n] = rb[i];
        }
    };

    load_slab(0);
    for (uint64_t k0 = 0; k0 < K; k0 += BK) {
        store_slab();
        __syncthreads();
        if (k0 + BK < K) load_slab(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            Ac a[8], b[8];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                a[i] = to_acc<T>(As[k][ty * 4 + i]);
                a[i + 4] = to_acc<T>(As[k][64 + ty * 4 + i]);
                b[i] = to_acc<T>(Bs[k][tx * 4 + i]);
                b[i + 4] = to_acc<T>(Bs[k][64 + tx * 4 + i]);
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = acc[i][j] + a[i] * b[j];
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint64_t gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn >= N) continue;
            Ac v = acc[i][j];
            // gemm_2x2.cl:238-256: alpha*acc + beta*C_old | alpha*acc | acc
            if (has_alpha) {
                if (has_beta) v = alpha * v + beta * to_acc<T>(C[gm * ldc + gn]);
                else v = alpha * v;
            }
            if (bias) v = v + to_acc<T>(bias[gn]);
            v = act_apply(v, act);
            C[gm * ldc + gn] = from_acc<T>(v);
        }
    }
}

// Small-tile variant for problems that would leave most SMs idle with 128 x 128 tiles (a 256^3 product is four of them):
// (16 MR) x (16 MR) tiles, MR x MR register tile per thread, MR = 4 or 2 -- 16x / 64x more CTAs for the same output.
template <typename T, int MR>
__global__ void __launch_bounds__(GT) gemm_simt_small_kernel(const T *__restrict__ A, const T *__restrict__ B, T *__restrict__ C,
                                                             uint64_t M, uint64_t N, uint64_t K, uint64_t lda, uint64_t ldb,
                                                             uint64_t ldc, int op_a, int op_b, int has_alpha, int has_beta,
                                                             typename Acc<T>::type alpha, typename Acc<T>::type beta,
                                                             const T *__restrict__ bias, int act) {
    using Ac = typename Acc<T>::type;
    constexpr int TM = 16 * MR, TK = 32, LOADS = TM * TK / GT;  // k-slab of 32: these problems are bound by the latency of a
    // slab's global loads (one CTA per SM, one slab of prefetch), so fewer, longer slabs; 2 MR elements of each tile per thread
    __shared__ T As[TK][TM + 4];
    __shared__ T Bs[TK][TM + 4];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const uint64_t m0 = (uint64_t)blockIdx.y * TM, n0 = (uint64_t)blockIdx.x * TM;
    Ac acc[MR][MR];
#pragma unroll
    for (int i = 0; i < MR; i++)
#pragma unroll
        for (int j = 0; j < MR; j++) acc[i][j] = (Ac)0;
    T ra[LOADS], rb[LOADS];
    auto load_slab = [&](uint64_t k0) {
#pragma unroll
        for (int i = 0; i < LOADS; i++) {
            const int e = t + i * GT;
            int m, k;
            if (op_a == 0) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
            const uint64_t gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < M && gk < K) ? (op_a == 0 ? A[gm * lda + gk] : A[gk * lda + gm]) : (T)0;
            int n, kk;
            if (op_b == 0) { n = e % TM; kk = e / TM; } else { kk = e % TK; n = e / TK; }
            const uint64_t gn = n0 + n, gk2 = k0 + kk;
            rb[i] = (gn < N && gk2 < K) ? (op_b == 0 ? B[gk2 * ldb + gn] : B[gn * ldb + gk2]) : (T)0;
        }
    };
    auto store_slab = [&]() {
#pragma unroll
        for (int i = 0; i < LOADS; i++) {
            const int e = t + i * GT;
            int m, k;
            if (op_a == 0) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
            As[k][m] = ra[i];
            int n, kk;
            if (op_b == 0) { n = e % TM; kk = e / TM; } else { kk = e % TK; n = e / TK; }
            Bs[kk][n] = rb[i];
        }
    };
    load_slab(0);
    for (uint64_t k0 = 0; k0 < K; k0 += TK) {
        store_slab();
        __syncthreads();
        if (k0 + TK < K) load_slab(k0 + TK);
#pragma unroll
        for (int k = 0; k < TK; k++) {
            Ac a[MR], b[MR];
#pragma unroll
            for (int i = 0; i < MR; i++) {
                a[i] = to_acc<T>(As[k][ty * MR + i]);
                b[i] = to_acc<T>(Bs[k][tx * MR + i]);
            }
#pragma unroll
            for (int i = 0; i < MR; i++)
#pragma unroll
                for (int j = 0; j < MR; j++) acc[i][j] = acc[i][j] + a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < MR; i++) {
        const uint64_t gm = m0 + ty * MR + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < MR; j++) {
            const uint64_t gn = n0 + tx * MR + j;
            if (gn >= N) continue;
            Ac v = acc[i][j];
            if (has_alpha) {  // gemm_2x2.cl:238-256
                if (has_beta) v = alpha * v + beta * to_acc<T>(C[gm * ldc + gn]);
                else v = alpha * v;
            }
            if (bias) v = v + to_acc<T>(bias[gn]);
            v = act_apply(v, act);
            C[gm * ldc + gn] = from_acc<T>(v);
        }
    }
}

int32_t gemm_simt(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K,
                  const void *alpha, const void *A, uint64_t lda, const void *B, uint64_t ldb, const void *beta, void *C,
                  uint64_t ldc, const void *bias, int32_t act) {
    return WK_DISPATCH_REAL(dtype, [&]() -> int32_t {
        using Ac = typename Acc<scalar_t>::type;
        const int has_alpha = (alpha != nullptr || beta != nullptr), has_beta = (beta != nullptr);
        const Ac al = alpha ? load_scalar<scalar_t>(alpha) : (Ac)1;  // gemm.zig:592-596
        const Ac be = beta ? load_scalar<scalar_t>(beta) : (Ac)0;
        // few 128 x 128 tiles: take 64 x 64 or 32 x 32 ones so the product spreads over the chip (summation order per element is
        // the same k-ascending one: integer results identical, float results identical to the big-tile kernel's)
        const uint64_t big_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
        if (big_tiles < (uint64_t)q->sm_count / 2) {
            const uint64_t t64 = ((M + 63) / 64) * ((N + 63) / 64);
            const int mr = t64 >= (uint64_t)q->sm_count / 2 ? 4 : 2;
            const uint64_t tm = 16 * mr, sy = (M + tm - 1) / tm, sx = (N + tm - 1) / tm;
            if (sy <= 65535) {
                const dim3 grid((unsigned)sx, (unsigned)sy);
                if (mr == 4)
                    gemm_simt_small_kernel<scalar_t, 4><<<grid, GT, 0, q->stream>>>((const scalar_t *)A, (const scalar_t *)B, (scalar_t *)C, M, N, K,
                                                                                   lda, ldb, ldc, op_a, op_b, has_alpha, has_beta, al, be,
                                                                                   (const scalar_t *)bias, act);
                else
                    gemm_simt_small_kernel<scalar_t, 2><<<grid, GT, 0, q->stream>>>((const scalar_t *)A, (const scalar_t *)B, (scalar_t *)C, M, N, K,
                                                                                   lda, ldb, ldc, op_a, op_b, has_alpha, has_beta, al, be,
                                                                                   (const scalar_t *)bias, act);
                WK_CHECK_LAUNCH();
                return WK_OK;
            }
        }
        const uint64_t gy = (M + BM - 1) / BM, gx = (N + BN - 1) / BN;
        if (gy > 65535) {
            set_error("gemm_simt: M too large");
            return WK_ERR_INVALID_VALUE;
        }
        gemm_simt_kernel<scalar_t><<<dim3((unsigned)gx, (unsigned)gy), GT, 0, q->stream>>>(
            (const scalar_t *)A, (const scalar_t *)B, (scalar_t *)C, M, N, K, lda, ldb, ldc, op_a, op_b, has_alpha, has_beta,
            al, be, (const scalar_t *)bias, act);
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}

}  // namespace wk
