// gemm_f64_tc.cu -- f64 blas.gemm on the FP64 tensor-core instruction (DMMA, mma.sync.m16n8k4.f64).
//
//   C[M,N] = alpha * op(A) * op(B) + beta * C        (src/blas/gemm.zig:834-874, all four transpose pairs)
//
// Blackwell has no tcgen05 kind for f64: FP64 tensor work is the warp-level mma.sync with register accumulators.
// Measured on this B200 (profiles/mma_peak_fp64_r01.txt) DMMA peaks at 37.0 TFLOP/s against 36.0 for plain DFMA,
// i.e. the FP64 pipe itself is the ceiling and one m16n8k4 occupies a sub-partition's pipe for 32 cycles.  What DMMA
// buys is issue slots: one instruction per 512 FMAs instead of one per 32, so the loads, address math and barriers
// of a tiled kernel hide completely behind the pipe (the SIMT kernel, which needs every other issue slot for a DFMA,
// stalls at 20 TFLOP/s).  The kernel is therefore a plain, deep, conflict-free pipeline:
//
//   CTA tile 128 x 128, k-block 16, 512 threads = 16 warps (4 x 4), warp tile 32 x 32 = 2 x 4 m16n8k4 tiles,
//   64 accumulator registers per thread;
//   global -> shared with 16-byte cp.async (zero fill at the edges), 4-buffer ring filled one k-block ahead (measured best of the depth sweep), full/empty
//   mbarriers per buffer instead of a CTA-wide barrier (warps drift by up to a k-block without stalling each other);
//   shared tiles padded to a pitch of 4 (mod 16) doubles so every fragment load (8 rows x 4 k) is conflict-free:
//     k-contiguous operand (A not transposed / B transposed):  tile[mn][16 + 4]
//     mn-contiguous operand (A transposed / B not transposed): tile[k][128 + 4]
//   epilogue straight from registers: alpha/beta/bias/activation, 16-byte stores (and peer GPUs for the fused
//   all-gather).
#include <stdlib.h>

#include "common.cuh"

namespace wk {
namespace dmma {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int THREADS = 512;
#ifndef WK_F64_STAGES
#define WK_F64_STAGES 4
#endif
#ifndef WK_F64_PREFETCH
#define WK_F64_PREFETCH 1
#endif
constexpr int STAGES = WK_F64_STAGES;
constexpr int PREFETCH = WK_F64_PREFETCH;  // k-blocks in flight ahead of the one being multiplied
constexpr int PK = BK + 4;     // pitch of a k-contiguous tile row (doubles)
constexpr int PMN = BM + 4;    // pitch of an mn-contiguous tile row (doubles)
constexpr int TILE_DOUBLES = BM * PK > BK * PMN ? BM * PK : BK * PMN;  // 2560
constexpr int STAGE_DOUBLES = 2 * TILE_DOUBLES;
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;  // 160 KiB

struct Params {
    const double *A, *B;
    double *C;
    uint64_t M, N, K, lda, ldb, ldc;
    double alpha, beta;
    int has_alpha, has_beta;
    const double *bias;
    int act;
    uint32_t tiles_m, tiles_n;
    // split-K (blockIdx.y = split): every split multiplies kb_per_split k-blocks, parks its accumulator fragments in
    // `ws` and takes a ticket; the last CTA of a tile folds the partials in split order and runs the epilogue
    uint32_t splits, kb_per_split;
    double *ws;
    unsigned *tickets;
    int n_peers, self;
    double *peers[16];
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {  // bounded: a protocol bug traps instead of hanging
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > 4096) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 6000000000LL) __trap();
        }
    }
}

__device__ __forceinline__ void dmma_m16n8k4(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a0), "d"(a1), "d"(b0));
}

__device__ __forceinline__ double apply_act(double v, int act) {
    if (act == WK_ACT_SIGMOID) return wk_sigmoid_f64(v);
    if (act == WK_ACT_TANH) return wk_tanh_f64(v);
    return v;
}

// One operand tile of a k-block into shared memory.  KCONTIG: stored [mn][k] (k contiguous in global memory) ->
// tile[mn][PK]; else stored [k][mn] -> tile[k][PMN].  1024 16-byte chunks per tile, 2 per thread.
template <bool KCONTIG>
__device__ __forceinline__ void load_tile(double *tile, const double *__restrict__ g, uint64_t ld, uint64_t mn0, uint64_t mn_dim,
                                          uint64_t k0, uint64_t k_dim) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int c = threadIdx.x + i * THREADS;
        if (KCONTIG) {
            const int r = c >> 3, kc = (c & 7) * 2;  // 128 rows x 8 chunks of 2 k
            const uint64_t gm = mn0 + r, gk = k0 + kc;
            int bytes = 0;
            if (gm < mn_dim && gk < k_dim) bytes = gk + 1 < k_dim ? 16 : 8;
            const double *src = bytes ? g + gm * ld + gk : g;
            cp_async16(tile + r * PK + kc, src, bytes);
        } else {
            const int r = c >> 6, mc = (c & 63) * 2;  // 16 k-rows x 64 chunks of 2 mn
            const uint64_t gk = k0 + r, gm = mn0 + mc;
            int bytes = 0;
            if (gk < k_dim && gm < mn_dim) bytes = gm + 1 < mn_dim ? 16 : 8;
            const double *src = bytes ? g + gk * ld + gm : g;
            cp_async16(tile + r * PMN + mc, src, bytes);
        }
    }
}

__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t tiles_m, uint32_t tiles_n, uint32_t &tm, uint32_t &tn) {
    // groups of 16 row-tiles sweep the columns together so the CTAs resident at the same time share panels in L2
    constexpr uint32_t GM = 16;
    const uint32_t per_group = GM * tiles_n;
    const uint32_t group = t / per_group, in_group = t - group * per_group;
    const uint32_t first_m = group * GM;
    const uint32_t gsize = min(GM, tiles_m - first_m);
    tm = first_m + in_group % gsize;
    tn = in_group / gsize;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(THREADS, 1) gemm_dmma_kernel(const Params p) {
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;

    uint32_t tm, tn;
    tile_coords(blockIdx.x, p.tiles_m, p.tiles_n, tm, tn);
    const uint64_t m0 = (uint64_t)tm * BM, n0 = (uint64_t)tn * BN;
    const uint32_t total_kb = (uint32_t)((p.K + BK - 1) / BK);
    const uint32_t kb_first = blockIdx.y * p.kb_per_split;
    const uint32_t num_kb = min(total_kb - kb_first, p.kb_per_split);  // k-blocks of this split, numbered from 0 below

    double acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][j][e] = 0.0;

    auto load_stage = [&](uint32_t kb) {
        double *sa = smem + (kb % STAGES) * STAGE_DOUBLES, *sb = sa + TILE_DOUBLES;
        load_tile<!TA>(sa, p.A, p.lda, m0, p.M, (uint64_t)(kb_first + kb) * BK, p.K);
        load_tile<TB>(sb, p.B, p.ldb, n0, p.N, (uint64_t)(kb_first + kb) * BK, p.K);
    };

    // full[b]: every thread's cp.async of the stage in buffer b has landed (cp.async.mbarrier.arrive.noinc, 512 arrivals)
    // empty[b]: all 16 warps are done reading buffer b.  A buffer is refilled PREFETCH k-blocks ahead, i.e. at least one
    // whole k-block after its last reader started the next one, so no warp ever waits for a straggler: there is no
    // CTA-wide barrier in the main loop (the __syncthreads version lost 14 % of its issue slots there,
    // profiles/ncu_gemm_f64_r01a.txt).
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    if (threadIdx.x == 0) {
        for (int b = 0; b < STAGES; b++) {
            mbar_init(&full_bar[b], THREADS);
            mbar_init(&empty_bar[b], THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_stage = [&](uint32_t j) {  // k-block j -> buffer j % STAGES
        const uint32_t b = j % STAGES, use = j / STAGES;
        if (use > 0) mbar_wait(&empty_bar[b], (use - 1) & 1);
        load_stage(j);
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&full_bar[b]))
                     : "memory");
    };
#pragma unroll
    for (int j = 0; j < PREFETCH; j++)
        if ((uint32_t)j < num_kb) issue_stage(j);

    for (uint32_t kb = 0; kb < num_kb; kb++) {
        if (kb + PREFETCH < num_kb) issue_stage(kb + PREFETCH);
        const uint32_t b = kb % STAGES;
        mbar_wait(&full_bar[b], (kb / STAGES) & 1);

        const double *sa = smem + b * STAGE_DOUBLES, *sb = sa + TILE_DOUBLES;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[2][2], b_[4];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                if (!TA) {  // tile[m][PK]
                    a[i][0] = sa[(wm + i * 16 + g) * PK + kk + t];
                    a[i][1] = sa[(wm + i * 16 + g + 8) * PK + kk + t];
                } else {    // tile[k][PMN]
                    a[i][0] = sa[(kk + t) * PMN + wm + i * 16 + g];
                    a[i][1] = sa[(kk + t) * PMN + wm + i * 16 + g + 8];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!TB) b_[j] = sb[(kk + t) * PMN + wn + j * 8 + g];  // tile[k][PMN]
                else b_[j] = sb[(wn + j * 8 + g) * PK + kk + t];       // tile[n][PK]
            }
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m16n8k4(acc[i][j], a[i][0], a[i][1], b_[j]);
        }
        __syncwarp();  // every lane's shared-memory reads of this buffer have been consumed by the MMAs above
        if (lane == 0) mbar_arrive(&empty_bar[b]);
    }

    if (p.splits > 1) {
        // fragments are owned by the same (warp, lane) in every CTA, so the workspace keeps them in register order:
        // ws[tile][split][pair of accumulators][thread] as double2 -> every warp store covers 512 contiguous bytes
        __shared__ uint32_t last_flag;
        double2 *mine = reinterpret_cast<double2 *>(p.ws) + ((uint64_t)blockIdx.x * p.splits + blockIdx.y) * (16 * THREADS) + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int h = 0; h < 2; h++)
                    __stcg(mine + ((i * 4 + j) * 2 + h) * THREADS, make_double2(acc[i][j][h * 2], acc[i][j][h * 2 + 1]));
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned old = atomicAdd(p.tickets + blockIdx.x, 1u);
            const bool last = (old == p.splits - 1);
            if (last) p.tickets[blockIdx.x] = 0;  // self-resetting
            last_flag = last ? 1u : 0u;
        }
        __syncthreads();
        if (!last_flag) return;
        __threadfence();
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[i][j][e] = 0.0;
        const double2 *part = reinterpret_cast<const double2 *>(p.ws) + (uint64_t)blockIdx.x * p.splits * (16 * THREADS) + threadIdx.x;
        for (uint32_t s2 = 0; s2 < p.splits; s2++, part += 16 * THREADS) {  // fixed order: independent of who is last
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const double2 v = __ldcg(part + ((i * 4 + j) * 2 + h) * THREADS);
                        acc[i][j][h * 2] += v.x;
                        acc[i][j][h * 2 + 1] += v.y;
                    }
        }
    }

    // ---------------------------------------------------------------- epilogue: c0,c1 = (row g, cols 2t,2t+1); c2,c3 = row g+8
    const bool vec_ok = (p.ldc % 2 == 0);
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint64_t row = m0 + wm + i * 16 + g + h * 8;
            if (row >= p.M) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t col = n0 + wn + j * 8 + 2 * t;
                if (col >= p.N) continue;
                double v0 = acc[i][j][h * 2], v1 = acc[i][j][h * 2 + 1];
                double *cp = p.C + row * p.ldc + col;
                const bool two = col + 1 < p.N;
                if (p.has_alpha) { v0 *= p.alpha; v1 *= p.alpha; }
                if (p.has_beta) {
                    if (two && vec_ok) {
                        const double2 o = *reinterpret_cast<const double2 *>(cp);
                        v0 += p.beta * o.x; v1 += p.beta * o.y;
                    } else {
                        v0 += p.beta * cp[0];
                        if (two) v1 += p.beta * cp[1];
                    }
                }
                if (p.bias) {
                    v0 += p.bias[col];
                    if (two) v1 += p.bias[col + 1];
                }
                if (p.act) { v0 = apply_act(v0, p.act); v1 = apply_act(v1, p.act); }
                if (two && vec_ok) {
                    *reinterpret_cast<double2 *>(cp) = make_double2(v0, v1);
                    for (int pi = 0; pi < p.n_peers; pi++)
                        if (pi != p.self) *reinterpret_cast<double2 *>(p.peers[pi] + row * p.ldc + col) = make_double2(v0, v1);
                } else {
                    cp[0] = v0;
                    if (two) cp[1] = v1;
                    for (int pi = 0; pi < p.n_peers; pi++)
                        if (pi != p.self) {
                            p.peers[pi][row * p.ldc + col] = v0;
                            if (two) p.peers[pi][row * p.ldc + col + 1] = v1;
                        }
                }
            }
        }
    }
}

}  // namespace dmma

int32_t gemm_f64_tc(wk_queue *q, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const double *alpha,
                    const double *A, uint64_t lda, const double *B, uint64_t ldb, const double *beta, double *C, uint64_t ldc,
                    const double *bias, int32_t act, const GemmPeers *peers) {
    using namespace dmma;
    // cp.async moves 16-byte chunks: bases and row starts must be 16-byte aligned
    if (!aligned16(A) || !aligned16(B) || (lda % 2) || (ldb % 2)) return -1;
    if (q->prop.major < 9) return -1;

    Params p{};
    p.A = A; p.B = B; p.C = C;
    p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.has_alpha = (alpha != nullptr || beta != nullptr);
    p.has_beta = (beta != nullptr);
    p.alpha = alpha ? *alpha : 1.0;
    p.beta = beta ? *beta : 0.0;
    p.bias = bias;
    p.act = act;
    p.tiles_m = (uint32_t)((M + BM - 1) / BM);
    p.tiles_n = (uint32_t)((N + BN - 1) / BN);
    if ((uint64_t)p.tiles_m * p.tiles_n > 0x7fffffffULL) return -1;
    p.n_peers = 0;
    p.self = 0;
    if (peers && peers->n > 1) {
        p.n_peers = peers->n;
        p.self = peers->self;
        for (int i = 0; i < peers->n; i++) {
            if (!aligned16(peers->ptrs[i])) return -1;
            p.peers[i] = (double *)peers->ptrs[i];
        }
    }
    if (!aligned16(C)) return -1;  // 16-byte epilogue accesses

    static bool attr_set[64] = {false};
    if (!attr_set[q->device & 63]) {
        WK_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set[q->device & 63] = true;
    }
    // split-K when the tiles alone leave more than half of the SMs idle: >= 16 k-blocks (256 k) per split, <= 8 splits
    const uint64_t n_tiles = (uint64_t)p.tiles_m * p.tiles_n;
    const uint32_t total_kb = (uint32_t)((K + BK - 1) / BK);
    static const char *sk_env = getenv("WK_GEMM_SPLITK");
    uint32_t splits = 1;
    if (sk_env && *sk_env) splits = (uint32_t)atoi(sk_env);
    else if (n_tiles * 2 <= (uint64_t)q->sm_count) splits = (uint32_t)((uint64_t)q->sm_count / n_tiles);
    if (splits > 8) splits = 8;
    if (splits > total_kb / 16) splits = total_kb / 16;
    if (splits < 1) splits = 1;
    p.kb_per_split = (total_kb + splits - 1) / splits;
    p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.ws = nullptr;
    p.tickets = nullptr;
    if (p.splits > 1) {
        int32_t rc = ensure_splitk(q, (size_t)n_tiles * p.splits * BM * BN * sizeof(double), (size_t)n_tiles);
        if (rc != WK_OK) return rc;
        p.ws = (double *)q->splitk_ws;
        p.tickets = q->splitk_tickets;
    }
    const dim3 grid((unsigned)n_tiles, p.splits);
    if (op_a == 0 && op_b == 0) gemm_dmma_kernel<false, false><<<grid, THREADS, SMEM_BYTES, q->stream>>>(p);
    else if (op_a == 0 && op_b == 1) gemm_dmma_kernel<false, true><<<grid, THREADS, SMEM_BYTES, q->stream>>>(p);
    else if (op_a == 1 && op_b == 0) gemm_dmma_kernel<true, false><<<grid, THREADS, SMEM_BYTES, q->stream>>>(p);
    else gemm_dmma_kernel<true, true><<<grid, THREADS, SMEM_BYTES, q->stream>>>(p);
    WK_CHECK_LAUNCH();
    return WK_OK;
}

}  // namespace wk
