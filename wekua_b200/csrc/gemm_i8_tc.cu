// gemm_i8_tc.cu -- integer blas.gemm on the 5th-generation tensor cores: tcgen05.mma.kind::i8 (u8 x u8 -> s32 in TMEM).
//
//   C[M,N] = alpha * op(A) * op(B) + beta * C   in exact arithmetic mod 2^bits  (src/blas/gemm.zig:834-874; the
//   reference's integer kernels, src/blas/kernels/gemm_nxn_gpu.cl:82-319, compute in the element type: OpenCL vector
//   lanes do not promote, so every product and sum wraps)
//
// One device kernel: D[128 x 256 per CTA] (s32, TMEM) = Abytes[M][k-range] . Bbytes[N][k-range]^T over unsigned bytes,
// followed by the epilogue  C = alpha * (D << shift) + beta * C  in W-byte integers.  What the host builds from it:
//   * i8 / u8 (W = 1): the product mod 2^8 only depends on the operands mod 2^8, so signed bytes are multiplied as
//     unsigned ones and the low byte of the s32 sum is the reference's result.  Operands are used where they lie when
//     they are K-major (A not transposed, B transposed) with 16-byte aligned rows; otherwise one streaming pass
//     re-tiles them (a transpose or a pitched copy: N^2 bytes against N^3 multiply-adds).
//   * i16 / u16, i32 / u32, i64 / u64 (W = 2, 4, 8): a = sum_i a_i 256^i with unsigned byte planes a_i, so
//     a.b mod 2^(8W) = sum_{i+j<W} 256^(i+j) a_i b_j.  With the planes of A concatenated along K in ascending order and the
//     planes of B in descending order, D_s = sum_{i+j=s} A_i B_j^T is ONE byte GEMM over the first (s+1) planes of A and
//     the last (s+1) planes of B: W launches, each adding alpha * (D_s << 8s) into C (beta' = 1 after the first).
//     W (W + 1) / 2 byte-GEMM equivalents in all -- 3, 10, 36 -- at the tensor cores' 8-bit rate.
//   The s32 accumulator cannot overflow: every launch covers at most 32768 byte products of <= 255^2 per output element
//   (longer K ranges are chunked on the host; chunks fold into C with the same beta' = 1 rule, exact mod 2^bits).
//
// Kernel structure (persistent, static tile schedule): warp 0 TMA producer (128-byte swizzled K-major boxes, zero fill
// out of bounds), warp 2 relays "stage landed" to the leader CTA of a pair, warp 1 issues 4 tcgen05.mma per k-block of
// 128 bytes, warps 4-7 read the accumulator with tcgen05.ld and store C.  CTAS = 2: cta_group::2, 256 x 256 tiles.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wk {
namespace i8tc {
using namespace tc;

constexpr int BM = 128, BN = 256, BKB = 128;  // BKB: k-bytes per stage = one 128-byte swizzle row = 4 MMAs of K = 32
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 6;
constexpr uint32_t MAX_K_PER_LAUNCH = 32768;  // 32768 * 255 * 255 < 2^31

template <int CTAS> struct Cfg {
    static constexpr int BN_LOAD = BN / CTAS;
    static constexpr int A_BYTES = BM * BKB;        // 16 KiB
    static constexpr int B_BYTES = BN_LOAD * BKB;   // 32 / 16 KiB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = CTAS == 1 ? 4 : 6;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

struct Params {
    void *C;
    uint64_t M, N, ldc;
    uint32_t num_kb;   // k-blocks of 128 bytes in this launch's k-range
    int32_t a_k0, b_k0;  // first k-byte of the range in A's / B's rows
    int has_alpha, has_beta;
    uint64_t alpha, beta;  // scalars zero-/sign-extended to the lane type (only the low 8W bits matter)
    int shift;             // D << shift before alpha (byte-plane group s: 8 s)
    uint32_t tiles_m, tiles_n, group_m;
};

struct Barriers {
    uint64_t full[MAX_STAGES];    // TMA -> relay                       (own CTA)
    uint64_t ready[MAX_STAGES];   // relays of both CTAs -> MMA         (leader CTA's copy is used)
    uint64_t empty[MAX_STAGES];   // MMA (commit, multicast) -> TMA     (own CTA)
    uint64_t acc_full[2];         // MMA (commit, multicast) -> epilogue
    uint64_t acc_empty[2];        // epilogue of both CTAs -> MMA       (leader CTA's copy is used)
    uint32_t tmem_base;
};

__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t tiles_m, uint32_t tiles_n, uint32_t GM, uint32_t &tm, uint32_t &tn) {
    const uint32_t per_group = GM * tiles_n;
    const uint32_t group = t / per_group, in_group = t - group * per_group;
    const uint32_t first_m = group * GM;
    const uint32_t gsize = min(GM, tiles_m - first_m);
    tm = first_m + in_group % gsize;
    tn = in_group / gsize;
}

template <int CTAS> __device__ __forceinline__ void arrive_on_leader(uint64_t *bar) {
    if (CTAS == 1) mbar_arrive(bar);
    else mbar_arrive_cluster(bar, 0);
}

template <int W> struct UInt;
template <> struct UInt<1> { using type = uint8_t; using lane = uint32_t; };
template <> struct UInt<2> { using type = uint16_t; using lane = uint32_t; };
template <> struct UInt<4> { using type = uint32_t; using lane = uint32_t; };
template <> struct UInt<8> { using type = uint64_t; using lane = uint64_t; };

template <int CTAS, int W>
__global__ void __launch_bounds__(THREADS, 1)
gemm_u8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
    using C = Cfg<CTAS>;
    using OutT = typename UInt<W>::type;
    using Lane = typename UInt<W>::lane;
    constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, A_BYTES = C::A_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    Barriers *bars = reinterpret_cast<Barriers *>(smem + STAGES * STAGE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CTAS == 1 ? 0u : cluster_ctarank();
    const uint32_t unit = blockIdx.x / CTAS, n_units = gridDim.x / CTAS;
    const uint32_t num_tiles = p.tiles_m * p.tiles_n;
    const uint32_t num_kb = p.num_kb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->ready[s], CTAS);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&bars->acc_full[a], 1);
            mbar_init(&bars->acc_empty[a], 4 * CTAS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CTAS == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
        else tmem_alloc_2cta(&bars->tmem_base, TMEM_COLS);
    }
    tc_fence_before();
    if (CTAS == 1) __syncthreads();
    else cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ================================================================= TMA producer (every CTA, own operand halves)
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = unit; t < num_tiles; t += n_units) {
                uint32_t tm, tn;
                tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
                const int32_t m0 = (int32_t)((tm * CTAS + rank) * BM);
                const int32_t n0 = (int32_t)(tn * BN + rank * C::BN_LOAD);
                for (uint32_t kb = 0; kb < num_kb; kb++, it++) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&bars->empty[s], ph ^ 1);
                    uint8_t *a_dst = smem + s * STAGE_BYTES, *b_dst = a_dst + A_BYTES;
                    mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
                    tma_load_2d(a_dst, &tmA, p.a_k0 + (int32_t)(kb * BKB), m0, &bars->full[s]);
                    tma_load_2d(b_dst, &tmB, p.b_k0 + (int32_t)(kb * BKB), n0, &bars->full[s]);
                }
            }
        }
    } else if (warp == 2) {
        // ================================================================= relay: this CTA's stage landed -> leader's MMA warp
        // (the data was written by the async proxy and is read by the async proxy; the barrier only orders them)
        uint32_t it = 0;
        for (uint32_t t = unit; t < num_tiles; t += n_units)
            for (uint32_t kb = 0; kb < num_kb; kb++, it++) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&bars->full[s], ph);
                __syncwarp();
                if (lane == 0) arrive_on_leader<CTAS>(&bars->ready[s]);
            }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (leader CTA only)
        if (rank == 0) {
            const uint32_t idesc = umma_idesc_u8(BM * CTAS, BN);
            const uint64_t base = umma_desc_base(16, 1024, UMMA_SW128);  // K-major, 128-byte rows, 8-row groups 1024 B apart
            uint32_t it = 0, tile_i = 0;
            for (uint32_t t = unit; t < num_tiles; t += n_units, tile_i++) {
                const uint32_t acc = tile_i & 1, acc_ph = (tile_i >> 1) & 1;
                mbar_wait(&bars->acc_empty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < num_kb; kb++, it++) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&bars->ready[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_s = smem_u32(smem + s * STAGE_BYTES), b_s = a_s + A_BYTES;
#pragma unroll
                        for (int k = 0; k < BKB / 32; k++) {  // K = 32 bytes per MMA: 32 bytes further along the swizzled row
                            const uint64_t da = umma_desc(base, a_s + k * 32), db = umma_desc(base, b_s + k * 32);
                            if (CTAS == 1) mma_i8_ss(d_tmem, da, db, idesc, (kb != 0) | (k != 0));
                            else mma_i8_ss_2cta(d_tmem, da, db, idesc, (kb != 0) | (k != 0));
                        }
                        if (CTAS == 1) {
                            mma_commit(&bars->empty[s]);
                            if (kb == num_kb - 1) mma_commit(&bars->acc_full[acc]);
                        } else {
                            mma_commit_2cta_multicast(&bars->empty[s], 3);
                            if (kb == num_kb - 1) mma_commit_2cta_multicast(&bars->acc_full[acc], 3);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue (every CTA: its 128 rows x 256 columns)
        const int q = warp & 3;
        uint32_t tile_i = 0;
        const Lane alpha = (Lane)p.alpha, beta = (Lane)p.beta;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.C) | (p.ldc * W)) & 15) == 0;
        for (uint32_t t = unit; t < num_tiles; t += n_units, tile_i++) {
            uint32_t tm, tn;
            tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
            const uint32_t acc = tile_i & 1, acc_ph = (tile_i >> 1) & 1;
            mbar_wait(&bars->acc_full[acc], acc_ph);
            tc_fence_after();
            const uint64_t row = (uint64_t)(tm * CTAS + rank) * BM + q * 32 + lane;  // TMEM lane = row of the tile
            const uint64_t col0 = (uint64_t)tn * BN;
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                tmem_ld_wait();
                if (c == BN / 32 - 1) {  // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_on_leader<CTAS>(&bars->acc_empty[acc]);
                }
                const uint64_t col = col0 + (uint64_t)c * 32;
                if (row >= p.M || col >= p.N) continue;
                OutT *cp = reinterpret_cast<OutT *>(p.C) + row * p.ldc + col;
                constexpr int PER16 = 16 / W;  // elements per 16-byte piece
                if (vec_ok && col + 32 <= p.N) {
#pragma unroll
                    for (int g = 0; g < 32 / PER16; g++) {
                        union { uint4 u; OutT e[PER16]; } o, old;
                        if (p.has_beta) old.u = *reinterpret_cast<const uint4 *>(cp + g * PER16);
#pragma unroll
                        for (int e = 0; e < PER16; e++) {
                            Lane v = (Lane)r[g * PER16 + e] << p.shift;  // the s32 sum is in [0, 2^31): zero-extension is exact
                            if (p.has_alpha) v = alpha * v;
                            if (p.has_beta) v += beta * (Lane)old.e[e];
                            o.e[e] = (OutT)v;
                        }
                        *reinterpret_cast<uint4 *>(cp + g * PER16) = o.u;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e++) {
                        if (col + e < p.N) {
                            Lane v = (Lane)r[e] << p.shift;
                            if (p.has_alpha) v = alpha * v;
                            if (p.has_beta) v += beta * (Lane)cp[e];
                            cp[e] = (OutT)v;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    if (CTAS == 1) __syncthreads();
    else cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (CTAS == 1) tmem_dealloc(tmem_base, TMEM_COLS);
        else tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------- operand staging
// Byte planes of a W-byte integer operand X[R][K] (row-major, pitch ld elements), K-major, every plane zero-padded to Kp (a
// multiple of 128) columns:   dst[r][plane_slot * Kp + k] = byte `plane` of X(r, k)
// plane_slot = plane (ascending, operand A) or W - 1 - plane (descending, operand B).  A thread takes 16 consecutive k of one
// row: 16 W bytes in (128-bit loads when the rows are 16-byte aligned), one 16-byte store per plane out -- 8 threads cover
// 128 contiguous bytes of every plane row.  Transposed operands are transposed first (the 128-bit tiled transpose kernel).
template <int W, bool VEC>
__global__ void __launch_bounds__(256) split_planes_kernel(const uint8_t *__restrict__ src, uint64_t ld, uint64_t R, uint64_t K,
                                                           uint8_t *__restrict__ dst, uint64_t Kp, int descending) {
    using ElemT = typename UInt<W>::type;
    const uint64_t r = (uint64_t)blockIdx.y * 32 + (threadIdx.x >> 3);
    const uint64_t k = ((uint64_t)blockIdx.x * 8 + (threadIdx.x & 7)) * 16;
    if (r >= R || k >= Kp) return;
    union { uint4 v[W]; ElemT e[16]; } in;
    const ElemT *s = reinterpret_cast<const ElemT *>(src) + r * ld + k;
    if (VEC && k + 16 <= K) {
#pragma unroll
        for (int i = 0; i < W; i++) in.v[i] = __ldg(reinterpret_cast<const uint4 *>(s) + i);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) in.e[i] = k + i < K ? s[i] : (ElemT)0;
    }
#pragma unroll
    for (int b = 0; b < W; b++) {
        uint32_t w[4];
#pragma unroll
        for (int jq = 0; jq < 4; jq++) {
            uint32_t x = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) x |= (uint32_t)((in.e[jq * 4 + e] >> (8 * b)) & 0xff) << (8 * e);  // little endian: byte b = plane b
            w[jq] = x;
        }
        const int slot = descending ? W - 1 - b : b;
        *reinterpret_cast<uint4 *>(dst + r * (uint64_t)W * Kp + (uint64_t)slot * Kp + k) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace i8tc

static int env_int_i8(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <int W>
static int32_t launch_u8(wk_queue *q, int ctas, const CUtensorMap &tmA, const CUtensorMap &tmB, const i8tc::Params &p) {
    using namespace i8tc;
    const uint64_t num_tiles = (uint64_t)p.tiles_m * p.tiles_n;
    const uint64_t max_units = (uint64_t)q->sm_count / ctas;
    const unsigned units = (unsigned)(num_tiles < max_units ? num_tiles : max_units);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units * ctas);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = ctas == 1 ? Cfg<1>::SMEM_BYTES : Cfg<2>::SMEM_BYTES;
    cfg.stream = q->stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (ctas == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        na++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    static bool attr_set[64] = {false};
    if (!attr_set[q->device & 63]) {
        WK_CUDA(cudaFuncSetAttribute(gemm_u8_kernel<1, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_u8_kernel<2, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES));
        attr_set[q->device & 63] = true;
    }
    if (ctas == 1) WK_CUDA(cudaLaunchKernelEx(&cfg, gemm_u8_kernel<1, W>, tmA, tmB, p));
    else WK_CUDA(cudaLaunchKernelEx(&cfg, gemm_u8_kernel<2, W>, tmA, tmB, p));
    WK_CHECK_LAUNCH();
    return WK_OK;
}

template <int W>
static int32_t split_planes(wk_queue *q, const void *src, uint64_t ld, uint64_t R, uint64_t K, void *dst, uint64_t Kp, int descending) {
    using namespace i8tc;
    const dim3 grid((unsigned)((Kp + 127) / 128), (unsigned)((R + 31) / 32));
    if (grid.y > 65535) {  // 2M rows: not a GEMM this path is for
        set_error("gemm_int_tc: operand too tall");
        return WK_ERR_INVALID_VALUE;
    }
    const bool vec = aligned16(src) && (ld * W) % 16 == 0;
    if (vec) split_planes_kernel<W, true><<<grid, 256, 0, q->stream>>>((const uint8_t *)src, ld, R, K, (uint8_t *)dst, Kp, descending);
    else split_planes_kernel<W, false><<<grid, 256, 0, q->stream>>>((const uint8_t *)src, ld, R, K, (uint8_t *)dst, Kp, descending);
    WK_CHECK_LAUNCH();
    return WK_OK;
}

// Returns -1 when the tensor-core path does not apply (the caller falls back to the SIMT kernel).
template <int W>
static int32_t gemm_int_tc_w(wk_queue *q, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const void *alpha,
                             const void *A, uint64_t lda, const void *B, uint64_t ldb, const void *beta, void *C, uint64_t ldc,
                             bool is_signed) {
    using namespace i8tc;
    using namespace tc;
    if (q->prop.major != 10) return -1;
    if (M > 0x7fffffffULL || N > 0x7fffffffULL || (K + 127) * W > 0x7fffffffULL) return -1;  // (k-byte offsets are int32 TMA coordinates)
    static const int ctas_env = env_int_i8("WK_GEMM_CTAS", 0);
    const int ctas = ctas_env == 1 || ctas_env == 2 ? ctas_env : (M > BM ? 2 : 1);

    // operands in place (W = 1, K-major, 16-byte aligned rows) or staged: byte planes [rows][W * Kp], K-major, zero-padded.
    // A transposed operand (op(X) = stored X^T) is transposed first by the 128-bit tiled transpose kernel -- straight into its
    // final place for W = 1 (one plane: the tensor map's bound K zero-fills the tail instead of padding), through a scratch
    // matrix for wider elements.
    const bool a_direct = W == 1 && op_a == 0 && aligned16(A) && lda % 16 == 0;
    const bool b_direct = W == 1 && op_b == 1 && aligned16(B) && ldb % 16 == 0;
    const bool a_trans = op_a == 1, b_trans = op_b == 0;
    const uint64_t Kp = (K + BKB - 1) / BKB * BKB;
    const uint64_t per16 = 16 / W, Kt = (K + per16 - 1) / per16 * per16;  // element pitch of a transposed scratch matrix
    auto al256 = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t a_bytes = a_direct ? 0 : al256((size_t)M * W * Kp), b_bytes = b_direct ? 0 : al256((size_t)N * W * Kp);
    const size_t ta_bytes = (!a_direct && a_trans && W > 1) ? al256((size_t)M * Kt * W) : 0;
    const size_t tb_bytes = (!b_direct && b_trans && W > 1) ? al256((size_t)N * Kt * W) : 0;
    const uint8_t *a_ptr = (const uint8_t *)A, *b_ptr = (const uint8_t *)B;
    uint64_t a_pitch = lda, b_pitch = ldb, a_inner = K, b_inner = K;
    if (a_bytes + b_bytes) {
        const size_t need = a_bytes + b_bytes + ta_bytes + tb_bytes;
        if (q->int_ws_bytes < need) {
            int32_t rc = grow_buffer(q, &q->int_ws, &q->int_ws_bytes, need);
            if (rc != WK_OK) return rc;
        }
        constexpr int32_t size_dtype = W == 1 ? 1 : W == 2 ? 3 : W == 4 ? 5 : 7;  // an unsigned dtype id of W bytes (transpose by size)
        auto stage = [&](const void *X, uint64_t ldx, uint64_t rows, bool trans, uint8_t *dst, uint8_t *scratch, int descending) -> int32_t {
            if (trans && W == 1) return wk_transpose2d(q, size_dtype, K, rows, X, ldx, dst, Kp);  // stored [K][rows] -> [rows][Kp]
            if (trans) {
                int32_t rc = wk_transpose2d(q, size_dtype, K, rows, X, ldx, scratch, Kt);
                if (rc != WK_OK) return rc;
                return split_planes<W>(q, scratch, Kt, rows, K, dst, Kp, descending);
            }
            return split_planes<W>(q, X, ldx, rows, K, dst, Kp, descending);
        };
        uint8_t *ws = (uint8_t *)q->int_ws;
        if (!a_direct) {
            int32_t rc = stage(A, lda, M, a_trans, ws, ws + a_bytes + b_bytes, 0);
            if (rc != WK_OK) return rc;
            a_ptr = ws;
            a_pitch = (uint64_t)W * Kp;
            a_inner = (W == 1 && a_trans) ? K : a_pitch;  // (transposed 8-bit operand: columns [K, Kp) were not written)
        }
        if (!b_direct) {
            uint8_t *dst = ws + a_bytes;
            int32_t rc = stage(B, ldb, N, b_trans, dst, ws + a_bytes + b_bytes + ta_bytes, 1);
            if (rc != WK_OK) return rc;
            b_ptr = dst;
            b_pitch = (uint64_t)W * Kp;
            b_inner = (W == 1 && b_trans) ? K : b_pitch;
        }
    }
    CUtensorMap tmA, tmB;
    if (!make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, a_ptr, a_inner, M, a_pitch, BKB, BM) ||
        !make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, b_ptr, b_inner, N, b_pitch, BKB, BN / ctas)) {
        set_error("gemm_int_tc: cuTensorMapEncodeTiled failed");
        return WK_ERR_CUDA;
    }

    Params p{};
    p.C = C;
    p.M = M; p.N = N; p.ldc = ldc;
    const bool has_alpha = (alpha != nullptr || beta != nullptr), has_beta = (beta != nullptr);
    auto load = [&](const void *s, uint64_t dflt) -> uint64_t {
        if (!s) return dflt;
        uint64_t v = 0;
        memcpy(&v, s, W);  // little endian: the low 8W bits; anything above them is irrelevant mod 2^(8W) ...
        if constexpr (W < 8) {
            if (is_signed && ((v >> (8 * W - 1)) & 1)) v |= ~0ULL << (8 * W);  // ... but keep the lane value the SIMT kernel uses
        }
        return v;
    };
    const uint64_t al = load(alpha, 1), be = load(beta, 0);
    p.tiles_m = (uint32_t)((M + (uint64_t)BM * ctas - 1) / ((uint64_t)BM * ctas));
    p.tiles_n = (uint32_t)((N + BN - 1) / BN);
    p.group_m = 16;

    // passes: byte-plane group s = all pairs (A_i, B_j) with i + j = s.  With A's planes ascending and B's descending in their
    // rows, the group is ONE contiguous k-range of (s + 1) * Kp bytes in both operands, cut into pieces of at most 32768
    // products (what an s32 accumulator takes).  (W = 1: one plane, slot 0, whether the operand is used in place or staged.)
    bool first = true;
    auto launch = [&](uint64_t a_off, uint64_t b_off, uint64_t len, int s) -> int32_t {
        p.a_k0 = (int32_t)a_off;
        p.b_k0 = (int32_t)b_off;
        p.num_kb = (uint32_t)((len + BKB - 1) / BKB);  // rows end at the tensor map's bound (in place) or are zero-padded (planes)
        p.shift = 8 * s;
        // C = alpha * (D << shift) + beta * C for the first launch; every later launch ADDS: beta' = 1 (exact mod 2^bits)
        p.has_alpha = first ? has_alpha : 1;
        p.has_beta = first ? has_beta : 1;
        p.alpha = al;
        p.beta = first ? be : 1;
        first = false;
        return launch_u8<W>(q, ctas, tmA, tmB, p);
    };
    for (int s = 0; s < W; s++) {
        // the group's k-range: planes 0..s of A (from byte 0) against planes s..0 of B (from slot W - 1 - s), position by position --
        // so ANY contiguous piece of it is a valid partial sum, and pieces of <= 32768 products keep the s32 accumulator exact
        const uint64_t total = ((uint64_t)s + 1) * Kp, b_base = (uint64_t)(W - 1 - s) * Kp;
        for (uint64_t k_lo = 0; k_lo < total; k_lo += MAX_K_PER_LAUNCH) {
            const uint64_t len = total - k_lo < MAX_K_PER_LAUNCH ? total - k_lo : MAX_K_PER_LAUNCH;
            int32_t rc = launch(k_lo, b_base + k_lo, len, s);
            if (rc != WK_OK) return rc;
        }
    }
    return WK_OK;
}

int32_t gemm_int_tc(wk_queue *q, int32_t dtype, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const void *alpha,
                    const void *A, uint64_t lda, const void *B, uint64_t ldb, const void *beta, void *C, uint64_t ldc) {
    const bool is_signed = (dtype % 2) == 0;  // ids 0..7: i8 u8 i16 u16 i32 u32 i64 u64
    switch (dtype) {
        case 0: case 1: return gemm_int_tc_w<1>(q, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, is_signed);
        case 2: case 3: return gemm_int_tc_w<2>(q, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, is_signed);
        case 4: case 5: return gemm_int_tc_w<4>(q, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, is_signed);
        case 6: case 7: return gemm_int_tc_w<8>(q, op_a, op_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, is_signed);
        default: return -1;
    }
}

}  // namespace wk
