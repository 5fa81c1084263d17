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

// Mid-size float problems (a 512^3 product is 16 tiles of 128 x 128 and 64 of 64 x 64 for 148 SMs): 32 x 64 tiles, 4 x 4 outputs
// per thread, 256 threads = two groups that each multiply one half of every k-slab (one 128-thread group per SM leaves each
// scheduler with a single warp and every shared-memory latency exposed).  With one or two such CTAs per SM the kernel is bound by the LATENCY of the operand loads, not by
// the FMA pipe (a register-staged version with one k-slab of prefetch took 0.7 us per slab: 24.8 us for 512^3), so the
// operands arrive through a 4-stage ring of 16-byte cp.async copies (zero-filled at the edges): three slabs of 128 bytes
// of k per row are in flight while one is multiplied.  A tile stays in the layout memory has -- [mn][k] when k is the
// contiguous dimension, [k][mn] otherwise (OPA / OPB) -- and the fragment loads adapt: along k a thread reads VW
// consecutive k of each of its rows / columns with one 128-bit load (row pitch = 9 x 16 bytes and columns 16 apart keep
// those loads conflict-free), along mn it reads VW consecutive rows / columns of one k.  An output element is the sum of two
// k-ascending FMA chains (its slab halves), added in a fixed order: deterministic, but not bit-identical to the single chain
// of the other SIMT kernels (float tolerance as for the tensor-core paths).  Needs every 16-byte vector along an
// operand's contiguous dimension to be whole (aligned base and pitch, the dimension a multiple of the vector width) --
// otherwise the element-wise kernels above run.
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
template <typename T> struct MidCfg {
    static constexpr int TM = 32, TN = 64, NT = 256, VW = 16 / (int)sizeof(T), TK = 128 / (int)sizeof(T), STAGES = 4;
    static constexpr int PK = TK + VW, PM = TM + VW, PN = TN + VW;  // row pitches (elements) of [mn][k], [k][m], [k][n] tiles
    static constexpr int a_elems(int opa) { return opa == 0 ? TM * PK : TK * PM; }
    static constexpr int b_elems(int opb) { return opb == 0 ? TK * PN : TN * PK; }
    static constexpr int smem_bytes(int opa, int opb) { return STAGES * (a_elems(opa) + b_elems(opb)) * (int)sizeof(T); }
};
template <typename T, int OPA, int OPB>
__global__ void __launch_bounds__(256) gemm_simt_mid_kernel(const T *__restrict__ A, const T *__restrict__ B, T *__restrict__ C,
                                                            uint64_t M, uint64_t N, uint64_t K, uint64_t lda, uint64_t ldb,
                                                            uint64_t ldc, int has_alpha, int has_beta, T alpha, T beta,
                                                            const T *__restrict__ bias, int act) {
    using Cf = MidCfg<T>;
    constexpr int TM = Cf::TM, TN = Cf::TN, NT = Cf::NT, VW = Cf::VW, TK = Cf::TK, STAGES = Cf::STAGES;
    constexpr int PK = Cf::PK, PM = Cf::PM, PN = Cf::PN, A_ELEMS = Cf::a_elems(OPA), STAGE_ELEMS = A_ELEMS + Cf::b_elems(OPB);
    constexpr int KV = TK / VW;  // 16-byte vectors along k per tile row: 8
    union Vec { uint4 u; T e[VW]; };
    extern __shared__ __align__(16) unsigned char mid_smem[];
    T *smem = reinterpret_cast<T *>(mid_smem);
    // two groups of 128 threads share a tile: group g multiplies the g-th half of every k-slab, the halves meet at the end
    const int t = threadIdx.x, grp = t >> 7, tx = t & 15, ty = (t >> 4) & 7;
    const uint64_t m0 = (uint64_t)blockIdx.y * TM, n0 = (uint64_t)blockIdx.x * TN;
    // rows of a thread: 4 / VW runs of VW (f32 one run of 4, f64 two runs of 2 half a tile apart); columns: the same when B is
    // [k][n] (16 lanes read 16 consecutive vectors of a row), tx + 16 j when B is [n][k] (16 lanes read 16 consecutive rows)
    auto row_of = [&](int i) { return (i / VW) * (8 * VW) + ty * VW + (i % VW); };
    auto col_of = [&](int j) { return OPB == 0 ? (j / VW) * (16 * VW) + tx * VW + (j % VW) : tx + 16 * j; };

    auto issue = [&](int stage, uint64_t k0) {
        T *As = smem + stage * STAGE_ELEMS, *Bs = As + A_ELEMS;
#pragma unroll
        for (int i = 0; i < TM * KV / NT; i++) {  // A: 256 vectors
            const int v = t + i * NT;
            uint64_t gm, gk;
            T *dst;
            if (OPA == 0) { gm = m0 + v / KV; gk = k0 + (v % KV) * VW; dst = As + (v / KV) * PK + (v % KV) * VW; }
            else { constexpr int MV = TM / VW; gk = k0 + v / MV; gm = m0 + (v % MV) * VW; dst = As + (v / MV) * PM + (v % MV) * VW; }
            const bool ok = gm < M && gk < K;
            cp_async16_zfill(dst, ok ? (OPA == 0 ? A + gm * lda + gk : A + gk * lda + gm) : A, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < TN * KV / NT; i++) {  // B: 512 vectors
            const int v = t + i * NT;
            uint64_t gn, gk;
            T *dst;
            if (OPB == 0) { constexpr int NV = TN / VW; gk = k0 + v / NV; gn = n0 + (v % NV) * VW; dst = Bs + (v / NV) * PN + (v % NV) * VW; }
            else { gn = n0 + v / KV; gk = k0 + (v % KV) * VW; dst = Bs + (v / KV) * PK + (v % KV) * VW; }
            const bool ok = gn < N && gk < K;
            cp_async16_zfill(dst, ok ? (OPB == 0 ? B + gk * ldb + gn : B + gn * ldb + gk) : B, ok ? 16 : 0);
        }
    };

    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = (T)0;

    const uint64_t n_slabs = (K + TK - 1) / TK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {  // prologue: STAGES - 1 slabs in flight (empty groups keep the count uniform)
        if ((uint64_t)s < n_slabs) issue(s, (uint64_t)s * TK);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (uint64_t sl = 0; sl < n_slabs; sl++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");  // slab sl has landed (this thread's copies) ...
        __syncthreads();  // ... and everybody's; everybody is also done with slab sl - 1, whose buffer is refilled next
        if (sl + STAGES - 1 < n_slabs) issue((int)((sl + STAGES - 1) % STAGES), (sl + STAGES - 1) * TK);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const T *As = smem + (sl % STAGES) * STAGE_ELEMS, *Bs = As + A_ELEMS;
#pragma unroll
        for (int kc = 0; kc < TK / 2; kc += VW) {
            const int kk = grp * (TK / 2) + kc;
            T a[4][VW], b[VW][4];  // [row][k], [k][col]
            if (OPA == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    Vec v;
                    v.u = *reinterpret_cast<const uint4 *>(As + row_of(i) * PK + kk);
#pragma unroll
                    for (int e = 0; e < VW; e++) a[i][e] = v.e[e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < VW; e++)
#pragma unroll
                    for (int h = 0; h < 4 / VW; h++) {
                        Vec v;
                        v.u = *reinterpret_cast<const uint4 *>(As + (kk + e) * PM + row_of(h * VW));
#pragma unroll
                        for (int r = 0; r < VW; r++) a[h * VW + r][e] = v.e[r];
                    }
            }
            if (OPB == 0) {
#pragma unroll
                for (int e = 0; e < VW; e++)
#pragma unroll
                    for (int h = 0; h < 4 / VW; h++) {
                        Vec v;
                        v.u = *reinterpret_cast<const uint4 *>(Bs + (kk + e) * PN + col_of(h * VW));
#pragma unroll
                        for (int r = 0; r < VW; r++) b[e][h * VW + r] = v.e[r];
                    }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    Vec v;
                    v.u = *reinterpret_cast<const uint4 *>(Bs + col_of(j) * PK + kk);
#pragma unroll
                    for (int e = 0; e < VW; e++) b[e][j] = v.e[e];
                }
            }
#pragma unroll
            for (int e = 0; e < VW; e++)
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = acc[i][j] + a[i][e] * b[e][j];
        }
    }

    // the second group's partial sums cross over through the (now idle) stage memory: 16 values per thread, thread-major with
    // an odd pitch; the first group adds them in a fixed order and runs the epilogue
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    T *xch = smem + (t & 127) * 17;
    if (grp == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) xch[i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (grp == 1) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint64_t gm = m0 + row_of(i);
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t gn = n0 + col_of(j);
            if (gn >= N) continue;
            T v = acc[i][j] + xch[i * 4 + j];
            if (has_alpha) {  // gemm_2x2.cl:238-256
                if (has_beta) v = alpha * v + beta * C[gm * ldc + gn];
                else v = alpha * v;
            }
            if (bias) v = v + bias[gn];
            v = act_apply(v, act);
            C[gm * ldc + gn] = v;
        }
    }
}

// can the 128-bit loader of gemm_simt_mid_kernel take these operands?
template <typename T>
static bool mid_loadable(int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const void *A, uint64_t lda, const void *B,
                         uint64_t ldb) {
    constexpr uint64_t VW = 16 / sizeof(T);
    if (!aligned16(A) || !aligned16(B) || lda % VW || ldb % VW) return false;
    if ((op_a == 0 ? K : M) % VW) return false;
    if ((op_b == 0 ? N : K) % VW) return false;
    return true;
}

// enough 32 x 64 tiles to occupy the chip (at least one for every second SM; 384^3 = 72 tiles measured slower than the 144
// tiles of 32 x 32 of the small kernel: 11.9 against 11.2 us) and operands the 16-byte copies can address
template <typename T>
static bool mid_eligible(wk_queue *q, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const void *A, uint64_t lda,
                         const void *B, uint64_t ldb) {
    static const int mid_env = getenv("WK_SIMT_MID") ? atoi(getenv("WK_SIMT_MID")) : 1;
    const uint64_t mid_y = (M + 31) / 32, mid_x = (N + 63) / 64;
    return mid_env && mid_x * mid_y >= (uint64_t)q->sm_count / 2 && mid_y <= 65535 && mid_loadable<T>(op_a, op_b, M, N, K, A, lda, B, ldb);
}
bool gemm_simt_mid_ok(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const void *A,
                      uint64_t lda, const void *B, uint64_t ldb) {
    if (dtype == 8) return mid_eligible<float>(q, op_a, op_b, M, N, K, A, lda, B, ldb);
    if (dtype == 9) return mid_eligible<double>(q, op_a, op_b, M, N, K, A, lda, B, ldb);
    return false;
}
// 32 x 64 tiles (f32 / f64) when eligible and 128 x 128 ones would leave the chip mostly empty
template <typename T, typename Ac>
static bool try_mid(wk_queue *q, uint64_t big_tiles, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const void *A,
                    uint64_t lda, const void *B, uint64_t ldb, void *C, uint64_t ldc, int has_alpha, int has_beta, Ac al, Ac be,
                    const void *bias, int32_t act) {
    if constexpr (std::is_floating_point<T>::value) {
        const uint64_t mid_y = (M + 31) / 32, mid_x = (N + 63) / 64;
        if (big_tiles >= 2 * (uint64_t)q->sm_count || !mid_eligible<T>(q, op_a, op_b, M, N, K, A, lda, B, ldb)) return false;
        static bool attr_set[2][2] = {{false, false}, {false, false}};  // per T (this function) and transpose pair
        auto launch = [&](auto kern, int smem_bytes) {
            if (!attr_set[op_a != 0][op_b != 0]) {  // more than the 48 KiB a kernel gets without asking
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
                attr_set[op_a != 0][op_b != 0] = true;
            }
            kern<<<dim3((unsigned)mid_x, (unsigned)mid_y), 256, smem_bytes, q->stream>>>((const T *)A, (const T *)B, (T *)C, M, N, K, lda, ldb, ldc,
                                                                                    has_alpha, has_beta, al, be, (const T *)bias, act);
        };
        if (op_a == 0 && op_b == 0) launch(gemm_simt_mid_kernel<T, 0, 0>, MidCfg<T>::smem_bytes(0, 0));
        else if (op_a == 0) launch(gemm_simt_mid_kernel<T, 0, 1>, MidCfg<T>::smem_bytes(0, 1));
        else if (op_b == 0) launch(gemm_simt_mid_kernel<T, 1, 0>, MidCfg<T>::smem_bytes(1, 0));
        else launch(gemm_simt_mid_kernel<T, 1, 1>, MidCfg<T>::smem_bytes(1, 1));
        return true;
    }
    return false;
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
        if (try_mid<scalar_t>(q, big_tiles, op_a, op_b, M, N, K, A, lda, B, ldb, C, ldc, has_alpha, has_beta, al, be, bias, act)) {
            WK_CHECK_LAUNCH();
            return WK_OK;
        }
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
