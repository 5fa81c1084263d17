// gemm_f32_tc.cu -- f32 blas.gemm on the 5th-generation tensor cores: 3xTF32 split with fp32 accumulation in TMEM.
//
//   C[M,N] = alpha * op(A) * op(B) + beta * C        (src/blas/gemm.zig:834-874, all four transpose pairs)
//
// Every fp32 operand x is split on the fly into hi = tf32(x) and lo = tf32(x - hi); the product is accumulated as
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (the lo*lo term, <= 2^-22 relative, is dropped), which keeps the result at fp32
// level while all multiplies run on tcgen05.mma.kind::tf32.
//
// Structure (persistent, static tile schedule, k-block = 32 = one 128-byte swizzle row).  Template parameter CTAS:
//   CTAS = 1  one CTA per SM, 128 x 256 output tile, tcgen05.mma.cta_group::1 (M = 128)
//   CTAS = 2  CTA pairs (cluster of 2), 256 x 256 output tile, tcgen05.mma.cta_group::2 (M = 256): each CTA stages
//             its own 128 rows of A and its own 128 columns of B, so the shared-memory traffic per flop -- the
//             measured limiter of the CTAS = 1 kernel (profiles/ncu_gemm_f32_r01a.txt) -- drops by a third.
//   warp 0      TMA producer   : cp.async.bulk.tensor (128B swizzle, zero fill out of bounds) -> raw fp32 stage
//   warps 2-9   converters     : raw -> hi (in place) and lo (second buffer), same swizzled layout, 128-bit smem ops
//   warp 1      MMA issuer     : one elected thread (leader CTA) issues 12 tcgen05.mma per k-block (4 k-steps x 3 split
//                                terms) into a fp32 TMEM accumulator; tcgen05.commit frees the smem stage / publishes the tile
//   warps 10-13 epilogue       : tcgen05.ld -> alpha/beta/bias/activation -> global (and peer GPUs for the fused
//                                all-gather); two TMEM accumulators (2 x 256 columns) let it overlap the next tile
// The transpose variants only change the TMA boxes and the UMMA descriptors (K-major vs MN-major operands); the
// converters are layout-agnostic because hi/lo keep the byte layout TMA produced.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wk {
namespace tc {

constexpr int BM = 128;   // rows of A / C per CTA
constexpr int BN = 256;   // columns of the output tile (UMMA N)
constexpr int BK = 32;    // one 128-byte swizzle row of fp32
constexpr int CONV_WARPS = 8, EPI_WARPS = 4;
constexpr int THREADS = (2 + CONV_WARPS + EPI_WARPS) * 32;  // 448
constexpr int TMEM_COLS = 512;                               // two 128x256 fp32 accumulators
constexpr int MAX_STAGES = 3;
constexpr int EPI_PITCH = 36;                                // floats per row of the (peer path's) epilogue transpose tile (32 + 4)
constexpr int EPI_WARP_BYTES = 8192;                         // per epilogue warp: two 32 x 32 fp32 TMA-store staging boxes (swizzled, 4 KiB each)
constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;        // 32 KiB, 1024-byte aligned (right after the stages)

template <int CTAS, bool PROL = false> struct Cfg {
    static constexpr int BN_LOAD = BN / CTAS;          // B columns staged by each CTA
    static constexpr int A_BYTES = BM * BK * 4;        // 16 KiB
    static constexpr int B_BYTES = BN_LOAD * BK * 4;   // 32 / 16 KiB
    static constexpr int RAW_BYTES = A_BYTES + B_BYTES;  // hi tiles (TMA lands here)
    static constexpr int Y_BYTES = PROL ? A_BYTES : 0;   // PROL: the tile of Y that multiplies the A tile (same geometry)
    static constexpr int STAGE_BYTES = 2 * RAW_BYTES + Y_BYTES;  // + lo tiles (+ Y tile)
    static constexpr int STAGES = PROL ? 2 : (CTAS == 1 ? 2 : 3);
    static constexpr int COLSUM_BYTES = PROL ? 4096 : 0;  // PROL: per-warp column-sum partials of the converter warps
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/ + COLSUM_BYTES + 1024 /*align slack*/;
};

struct Params {
    float *C;
    uint64_t M, N, K, ldc;
    float alpha, beta;
    int has_alpha, has_beta;
    const float *bias;
    int act;
    int op_a, op_b;
    int a_mn3d, b_mn3d;  // MN-major operand loaded as ONE 3-D TMA box per stage instead of 4-8 32 x 32 boxes (MN % 32 == 0)
    int c_vec;  // C rows are 16-byte aligned: 128-bit epilogue accesses
    int c_tma;  // ... and C has a tensor map: finished 32 x 32 chunks leave through TMA stores
    int split;  // 0: store hi = rna_tf32(x) explicitly; 1: leave x in place, the tensor core truncates; 2: rounds
    uint32_t tiles_m, tiles_n, group_m;
    // split-K (small problems that cannot fill the chip with output tiles): a work item is (tile, split); every split
    // accumulates kb_per_split k-blocks into its own TMEM accumulator, parks the raw partial tile in `ws`, counts its
    // arrival and waits for the tile's other splits (all resident); then every split folds its share of the tile's
    // column chunks -- partials added in split order (deterministic) -- and runs the epilogue on them.
    uint32_t splits, kb_per_split;
    uint32_t n_full;  // tiles [0, n_full) are not split (whole waves); only the tail tiles [n_full, tiles) are
    float *ws;
    unsigned *tickets;
    int n_peers, self;
    int peer_bulk;  // fused all-gather: peers receive whole 128-byte row segments as bulk async copies from the staging tile
    int peer_tma;   // fused all-gather: every finished 32 x 32 chunk leaves as ONE TMA store per rank (PeerMaps), self included
    float *peers[16];
    // A-prologue (Linear.backward, src/nn/layer/linear.zig:608-613 fused into the GEMM that consumes it): the A operand is
    // used as A o act'(Y), Y = the layer output with A's shape and layout; colsum (op_a = T only) receives the column sums
    // of A o act'(Y) over K -- getBiasSensitivity, linear.zig:534-577 -- from the tiles of the first tile column
    int prol_act;
    float *colsum;
    // WK_GEMM_TRACE=1 (debugging aid for launch-bound problems): %globaltimer stamps of the pipeline's milestones, 16 slots per CTA
    unsigned long long *trace;
};

__device__ __forceinline__ void trace_stamp(const Params &p, int slot) {
    if (p.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
        p.trace[(size_t)blockIdx.x * 16 + slot] = t;
    }
}

// fused all-gather through the TMA engine: one tensor map per rank's C (this rank's row block of it), self included
constexpr int MAX_PEER_MAPS = 8;
struct PeerMaps {
    CUtensorMap m[MAX_PEER_MAPS];
};

struct Barriers {
    uint64_t raw_full[MAX_STAGES];    // TMA -> converters                      (own CTA)
    uint64_t conv_done[MAX_STAGES];   // converters of both CTAs -> MMA         (leader CTA's copy is used)
    uint64_t stage_free[MAX_STAGES];  // MMA (commit, multicast) -> TMA         (own CTA)
    uint64_t acc_full[2];             // MMA (commit, multicast) -> epilogue    (own CTA)
    uint64_t acc_empty[2];            // epilogue of both CTAs -> MMA           (leader CTA's copy is used)
    uint64_t fold_full;               // split-K: the other splits' partials of my chunks have landed in shared memory
    uint32_t tmem_base;
};

// work item -> (tile, first k-block, end k-block)
__device__ __forceinline__ void work_coords(const Params &p, uint32_t w, uint32_t num_kb, uint32_t &tile, uint32_t &sp, uint32_t &kb0,
                                            uint32_t &kb1) {
    if (w < p.n_full) {  // a whole tile
        tile = w;
        sp = 0;
        kb0 = 0;
        kb1 = num_kb;
        return;
    }
    const uint32_t r = w - p.n_full;
    const uint32_t q = r / p.splits;
    tile = p.n_full + q;
    sp = r - q * p.splits;
    kb0 = sp * p.kb_per_split;
    kb1 = min(num_kb, kb0 + p.kb_per_split);
}

__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t tiles_m, uint32_t tiles_n, uint32_t GM, uint32_t &tm, uint32_t &tn) {
    // groups of GM row-tiles sweep the columns together so concurrently resident tiles share A and B panels in L2
    const uint32_t per_group = GM * tiles_n;
    const uint32_t group = t / per_group, in_group = t - group * per_group;
    const uint32_t first_m = group * GM;
    const uint32_t gsize = min(GM, tiles_m - first_m);
    tm = first_m + in_group % gsize;
    tn = in_group / gsize;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == WK_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
    if (act == WK_ACT_TANH) return tanhf(v);
    return v;
}

// out of line: the activation's libdevice code (expf, tanhf) would otherwise be inlined at every unrolled epilogue position and
// push the hot store path out of the instruction cache
// (by value: an array passed by reference would force the caller's registers into local memory)
__device__ __noinline__ float apply_act1(float v, int act) { return apply_act(v, act); }
__device__ __noinline__ float4 apply_act4(float4 v, int act) {
    return make_float4(apply_act(v.x, act), apply_act(v.y, act), apply_act(v.z, act), apply_act(v.w, act));
}

template <int CTAS> __device__ __forceinline__ void arrive_on_leader(uint64_t *bar) {
    if (CTAS == 1) mbar_arrive(bar);
    else mbar_arrive_cluster(bar, 0);
}

// PRE: the lo planes were computed by split_lo_kernel beforehand and arrive by TMA next to the raw tiles; the converter warps
// only relay the barrier (no shared-memory traffic of their own).  Same arithmetic, bit-identical results (split mode 1).
template <int CTAS, bool PRE, bool PROL = false>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ PeerMaps pm, const Params p) {
    static_assert(!(PRE && PROL), "the pre-split experiment has no prologue variant");
    using C = Cfg<CTAS, PROL>;  // PROL: tmAlo is the tensor map of Y
    constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, RAW_BYTES = C::RAW_BYTES, A_BYTES = C::A_BYTES;
    extern __shared__ uint8_t smem_raw[];
    // the dynamic smem base has the same CTA-relative offset in both CTAs of a pair, so the aligned tiles do too
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    Barriers *bars = reinterpret_cast<Barriers *>(smem + STAGES * STAGE_BYTES + EPI_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CTAS == 1 ? 0u : cluster_ctarank();
    const uint32_t unit = blockIdx.x / CTAS, n_units = gridDim.x / CTAS;  // a unit = one CTA or one CTA pair
    const uint32_t num_tiles = p.n_full + (p.tiles_m * p.tiles_n - p.n_full) * p.splits;  // work items: tiles, then (tail tile, split)
    const uint32_t num_kb = (uint32_t)((p.K + BK - 1) / BK);
    if (threadIdx.x == 0) trace_stamp(p, 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (PRE) {
            tma_prefetch_desc(&tmAlo);
            tma_prefetch_desc(&tmBlo);
        }
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&bars->raw_full[s], 1);
            mbar_init(&bars->conv_done[s], CONV_WARPS * CTAS);
            mbar_init(&bars->stage_free[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&bars->acc_full[a], 1);
            mbar_init(&bars->acc_empty[a], EPI_WARPS * CTAS);
        }
        mbar_init(&bars->fold_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CTAS == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
        else tmem_alloc_2cta(&bars->tmem_base, TMEM_COLS);
    }
    tc_fence_before();
    if (CTAS == 1) __syncthreads();
    else cluster_sync_all();  // the peer's barriers must be initialised before anyone arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    if (threadIdx.x == 0) trace_stamp(p, 1);

    if (warp == 0) {
        // ================================================================= TMA producer (every CTA, own operand halves)
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t w = unit; w < num_tiles; w += n_units) {
                uint32_t t, sp, kb0, kb1, tm, tn;
                work_coords(p, w, num_kb, t, sp, kb0, kb1);
                tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
                const int32_t m0 = (int32_t)((tm * CTAS + rank) * BM);
                const int32_t n0 = (int32_t)(tn * BN + rank * C::BN_LOAD);
                for (uint32_t kb = kb0; kb < kb1; kb++, it++) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&bars->stage_free[s], ph ^ 1);
                    uint8_t *a_dst = smem + s * STAGE_BYTES, *b_dst = a_dst + A_BYTES;
                    mbar_arrive_expect_tx(&bars->raw_full[s], PRE ? 2 * RAW_BYTES : RAW_BYTES + C::Y_BYTES);
                    if (PROL) {  // the Y tile: A's boxes, landing behind the lo tiles
                        uint8_t *yd = a_dst + 2 * RAW_BYTES;
                        const int32_t ky = (int32_t)(kb * BK);
                        if (p.op_a == 0) tma_load_2d(yd, &tmAlo, ky, m0, &bars->raw_full[s]);
                        else
                            for (int j = 0; j < BM / 32; j++) tma_load_2d(yd + j * 4096, &tmAlo, m0 + 32 * j, ky, &bars->raw_full[s]);
                    }
                    const int32_t k0 = (int32_t)(kb * BK);
#pragma unroll
                    for (int plane = 0; plane < (PRE ? 2 : 1); plane++) {  // raw tiles, then (PRE) the lo tiles RAW_BYTES further on
                        const CUtensorMap *ma = plane ? &tmAlo : &tmA, *mb = plane ? &tmBlo : &tmB;
                        uint8_t *ad = a_dst + plane * RAW_BYTES, *bd = b_dst + plane * RAW_BYTES;
                        if (p.op_a == 0) {  // A[M][K]: one 32(k) x 128(m) box, K-major
                            tma_load_2d(ad, ma, k0, m0, &bars->raw_full[s]);
                        } else if (p.a_mn3d && !plane) {  // A[K][M], M % 32 == 0: one (32, 32 k, 4 chunks) box
                            tma_load_3d(ad, ma, 0, k0, m0 / 32, &bars->raw_full[s]);
                        } else {            // A[K][M]: four 32(m) x 32(k) boxes, MN-major
                            for (int j = 0; j < BM / 32; j++) tma_load_2d(ad + j * 4096, ma, m0 + 32 * j, k0, &bars->raw_full[s]);
                        }
                        if (p.op_b == 1) {  // B[N][K]: one 32(k) x BN_LOAD(n) box, K-major
                            tma_load_2d(bd, mb, k0, n0, &bars->raw_full[s]);
                        } else if (p.b_mn3d && !plane) {  // B[K][N], N % 32 == 0: one (32, 32 k, BN_LOAD / 32 chunks) box
                            tma_load_3d(bd, mb, 0, k0, n0 / 32, &bars->raw_full[s]);
                        } else {            // B[K][N]: 32(n) x 32(k) boxes, MN-major
                            for (int j = 0; j < C::BN_LOAD / 32; j++) tma_load_2d(bd + j * 4096, mb, n0 + 32 * j, k0, &bars->raw_full[s]);
                        }
                    }
                    if (it == 0) trace_stamp(p, 2);
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (leader CTA only)
        if (rank == 0) {
            const uint32_t idesc = umma_idesc_tf32(BM * CTAS, BN, p.op_a == 1, p.op_b == 0);
            // K-major tile (SW128): 128-byte rows of 32 k, 8-row groups 1024 B apart (SBO).
            // MN-major tile (SW128 with 32-byte atoms): 128-byte rows of 32 mn, one row per k, atoms of 4 k-rows = 512 B
            // (SBO), 32-wide MN chunks 4096 B apart (LBO).
            const uint64_t a_base = p.op_a == 0 ? umma_desc_base(16, 1024, UMMA_SW128) : umma_desc_base(4096, 512, UMMA_SW128_32B);
            const uint64_t b_base = p.op_b == 1 ? umma_desc_base(16, 1024, UMMA_SW128) : umma_desc_base(4096, 512, UMMA_SW128_32B);
            const uint32_t a_kstep = p.op_a == 0 ? 32 : 1024;  // bytes per k-step of 8
            const uint32_t b_kstep = p.op_b == 1 ? 32 : 1024;
            uint32_t it = 0, tile_i = 0;
            for (uint32_t w = unit; w < num_tiles; w += n_units, tile_i++) {
                uint32_t t, sp, kb0, kb1;
                work_coords(p, w, num_kb, t, sp, kb0, kb1);
                const uint32_t acc = tile_i & 1, acc_ph = (tile_i >> 1) & 1;
                mbar_wait(&bars->acc_empty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = kb0; kb < kb1; kb++, it++) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&bars->conv_done[s], ph);
                    tc_fence_after();
                    if (it == 0 && lane == 0) trace_stamp(p, 4);
                    if (elect_one()) {
                        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), b_hi = a_hi + A_BYTES;
                        const uint32_t a_lo = a_hi + RAW_BYTES, b_lo = b_hi + RAW_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 8; k++) {
                            const uint64_t dah = umma_desc(a_base, a_hi + k * a_kstep), dal = umma_desc(a_base, a_lo + k * a_kstep);
                            const uint64_t dbh = umma_desc(b_base, b_hi + k * b_kstep), dbl = umma_desc(b_base, b_lo + k * b_kstep);
                            if (CTAS == 1) {
                                mma_tf32_ss(d_tmem, dal, dbh, idesc, (kb != kb0) | (k != 0));
                                mma_tf32_ss(d_tmem, dah, dbl, idesc, 1);
                                mma_tf32_ss(d_tmem, dah, dbh, idesc, 1);
                            } else {
                                mma_tf32_ss_2cta(d_tmem, dal, dbh, idesc, (kb != kb0) | (k != 0));
                                mma_tf32_ss_2cta(d_tmem, dah, dbl, idesc, 1);
                                mma_tf32_ss_2cta(d_tmem, dah, dbh, idesc, 1);
                            }
                        }
                        // smem stage reusable (in both CTAs) once these MMAs retire; last k-block: accumulator complete
                        if (CTAS == 1) {
                            mma_commit(&bars->stage_free[s]);
                            if (kb == kb1 - 1) mma_commit(&bars->acc_full[acc]);
                        } else {
                            mma_commit_2cta_multicast(&bars->stage_free[s], 3);
                            if (kb == kb1 - 1) mma_commit_2cta_multicast(&bars->acc_full[acc], 3);
                        }
                    }
                    __syncwarp();
                }
                if (lane == 0) trace_stamp(p, 5);
            }
        }
    } else if (warp < 2 + CONV_WARPS) {
        // ================================================================= converters: raw -> (hi, lo)
        const int ct = threadIdx.x - 64;  // 0..255
        const int split = p.split;
        uint32_t it = 0;
        for (uint32_t w = unit; w < num_tiles; w += n_units) {
            uint32_t t, sp, kb0, kb1;
            work_coords(p, w, num_kb, t, sp, kb0, kb1);
            // PROL: column sums of A o act'(Y) over K (the bias gradient), taken by the tiles of the first tile column while
            // the operand passes through the converters' registers anyway
            bool do_colsum = false;
            uint32_t cs_m0 = 0;
            float4 cs[4];
            if (PROL) {
                uint32_t tm, tn;
                tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
                do_colsum = p.colsum != nullptr && tn == 0;
                cs_m0 = (tm * CTAS + rank) * BM;
#pragma unroll
                for (int j = 0; j < 4; j++) cs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (uint32_t kb = kb0; kb < kb1; kb++, it++) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&bars->raw_full[s], ph);
                if (it == 0 && ct == 0) trace_stamp(p, 3);
                if (PRE) {  // both planes came by TMA (async proxy -> async proxy): nothing to convert, relay the barrier
                    __syncwarp();
                    if (lane == 0) arrive_on_leader<CTAS>(&bars->conv_done[s]);
                    continue;
                }
                float4 *hi = reinterpret_cast<float4 *>(smem + s * STAGE_BYTES);
                float4 *lo = reinterpret_cast<float4 *>(smem + s * STAGE_BYTES + RAW_BYTES);
                const float4 *yt = reinterpret_cast<const float4 *>(smem + s * STAGE_BYTES + 2 * RAW_BYTES);
#pragma unroll
                for (int j = 0; j < RAW_BYTES / 16 / (CONV_WARPS * 32); j++) {
                    const int i = ct + j * (CONV_WARPS * 32);
                    const float4 x = hi[i];
                    float xs[4] = {x.x, x.y, x.z, x.w};
                    const bool a_part = PROL && j < A_BYTES / 16 / (CONV_WARPS * 32);  // the first 4 pieces of a thread lie in A
                    if (a_part) {
                        // a' = a * act'(y), the arithmetic of wk_act_backward (ActBackwardF) without FMA contraction, so the
                        // fused operand is bit-identical to the tensor the unfused sequence stores
                        const float4 y4 = yt[i];
                        const float ys[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float d = p.prol_act == WK_ACT_SIGMOID ? __fmul_rn(ys[e], __fsub_rn(1.0f, ys[e]))
                                                                          : __fsub_rn(1.0f, __fmul_rn(ys[e], ys[e]));
                            xs[e] = __fmul_rn(xs[e], d);
                        }
                        if (do_colsum) {
                            cs[j].x += xs[0]; cs[j].y += xs[1]; cs[j].z += xs[2]; cs[j].w += xs[3];
                        }
                    }
                    float hs[4], ls[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        uint32_t hb, lb;
                        if (split == 1) hb = __float_as_uint(xs[e]) & 0xffffe000u;
                        else asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(xs[e]));
                        const float hf = __uint_as_float(hb);
                        // inf - inf would poison the product: a non-finite hi carries the value alone
                        const float rem = (hb & 0x7f800000u) == 0x7f800000u ? 0.0f : xs[e] - hf;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
                        hs[e] = hf;
                        ls[e] = __uint_as_float(lb);
                    }
                    if (split == 0) hi[i] = make_float4(hs[0], hs[1], hs[2], hs[3]);
                    else if (a_part) hi[i] = make_float4(xs[0], xs[1], xs[2], xs[3]);  // the product replaces the raw tile
                    lo[i] = make_float4(ls[0], ls[1], ls[2], ls[3]);
                }
                // generic-proxy smem writes -> visible to the tensor core (async proxy), then release to the MMA warp
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) arrive_on_leader<CTAS>(&bars->conv_done[s]);
            }
            if (PROL && p.colsum != nullptr) {  // (uniform over the CTA's converter warps: named barrier 2)
                // A is MN-major here (op_a = T): thread ct holds, in box j (32 m x 32 k), row k = ct >> 3 and the physical 16-byte
                // piece ct & 7; the 32-byte-atom swizzle (Swizzle<2,5,2>) puts logical piece lc at lc ^ ((k & 3) << 1).  A warp
                // covers 4 rows: two shuffles add them per logical piece, the 8 warps' results meet in shared memory and are
                // added in warp order -- a fixed order, no atomics.
                float4 *part = reinterpret_cast<float4 *>(smem + STAGES * STAGE_BYTES + EPI_BYTES + 256);
                const int cw = ct >> 5, r_loc = (ct >> 3) & 3, lc = (ct & 7) ^ (r_loc << 1);
                if (do_colsum) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float4 v = cs[j];
                        v.x += __shfl_xor_sync(0xffffffffu, v.x, 10); v.y += __shfl_xor_sync(0xffffffffu, v.y, 10);
                        v.z += __shfl_xor_sync(0xffffffffu, v.z, 10); v.w += __shfl_xor_sync(0xffffffffu, v.w, 10);
                        v.x += __shfl_xor_sync(0xffffffffu, v.x, 20); v.y += __shfl_xor_sync(0xffffffffu, v.y, 20);
                        v.z += __shfl_xor_sync(0xffffffffu, v.z, 20); v.w += __shfl_xor_sync(0xffffffffu, v.w, 20);
                        if (r_loc == 0) part[(cw * 4 + j) * 8 + lc] = v;
                    }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(CONV_WARPS * 32) : "memory");
                if (do_colsum && ct < BM) {
                    const int j = ct >> 5, l2 = (ct >> 2) & 7, e = ct & 3;
                    float sum = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < CONV_WARPS; w8++) {
                        const float4 v = part[(w8 * 4 + j) * 8 + l2];
                        sum += e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w;
                    }
                    if (cs_m0 + ct < p.M) p.colsum[cs_m0 + ct] = sum;
                }
                asm volatile("bar.sync 2, %0;" ::"n"(CONV_WARPS * 32) : "memory");
            }
        }
    } else {
        // ================================================================= epilogue (every CTA: its 128 rows x 256 columns)
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        uint8_t *const epi_raw = smem + STAGES * STAGE_BYTES + q * EPI_WARP_BYTES;  // 1024-byte aligned
        float *epi = reinterpret_cast<float *>(epi_raw);
        uint32_t tile_i = 0;
        const bool vec_ok = p.c_vec != 0;
        for (uint32_t w = unit; w < num_tiles; w += n_units, tile_i++) {
            uint32_t t, sp, kb0, kb1, tm, tn;
            work_coords(p, w, num_kb, t, sp, kb0, kb1);
            tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
            const uint32_t acc = tile_i & 1, acc_ph = (tile_i >> 1) & 1;
            mbar_wait(&bars->acc_full[acc], acc_ph);
            tc_fence_after();
            if (q == 0 && lane == 0) trace_stamp(p, 6);
            const uint64_t row0 = (uint64_t)(tm * CTAS + rank) * BM + q * 32;
            const uint64_t col0 = (uint64_t)tn * BN;
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);

            // r = 32 consecutive columns [c*32, c*32+32) of this lane's row -> alpha/beta/bias/activation -> C (and peers).
            // TMEM hands every lane one ROW (32 columns); global memory wants every warp instruction to cover whole
            // 128-byte row segments (local HBM sectors and, for the fused all-gather, NVLink packets).  Transpose the
            // 32 x 32 chunk through a padded per-warp shared-memory tile: 16-byte writes at a 144-byte row pitch and
            // 16-byte reads of 8 lanes per row are both conflict-free.
            auto store_chunk = [&](const uint32_t (&r)[32], int c) {
                float4 *my_row = reinterpret_cast<float4 *>(epi + lane * EPI_PITCH);
#pragma unroll
                for (int g = 0; g < 8; g++)
                    my_row[g] = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]),
                                            __uint_as_float(r[g * 4 + 3]));
                __syncwarp();
                const uint64_t col = col0 + (uint64_t)c * 32 + (lane & 7) * 4;
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int rl = it * 4 + (lane >> 3);
                    const uint64_t row = row0 + rl;
                    if (row >= p.M || col >= p.N) continue;
                    const float4 a4 = *reinterpret_cast<const float4 *>(epi + rl * EPI_PITCH + (lane & 7) * 4);
                    float v[4] = {a4.x, a4.y, a4.z, a4.w};
                    float *cp = p.C + row * p.ldc + col;
                    const bool full = vec_ok && (col + 4 <= p.N);
                    if (p.has_alpha) {
#pragma unroll
                        for (int e = 0; e < 4; e++) v[e] *= p.alpha;
                    }
                    if (full) {
                        if (p.has_beta) {
                            const float4 o = *reinterpret_cast<const float4 *>(cp);
                            v[0] += p.beta * o.x; v[1] += p.beta * o.y; v[2] += p.beta * o.z; v[3] += p.beta * o.w;
                        }
                        if (p.bias) {
                            const float4 bv = *reinterpret_cast<const float4 *>(p.bias + col);
                            v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
                        }
                        float4 out = make_float4(v[0], v[1], v[2], v[3]);
                        if (p.act) out = apply_act4(out, p.act);
                        *reinterpret_cast<float4 *>(cp) = out;
                        if (p.peer_bulk) {  // the finished values go back into the staging tile; the copy engine ships them
                            *reinterpret_cast<float4 *>(epi + rl * EPI_PITCH + (lane & 7) * 4) = out;
                        } else {
                            for (int pi = 0; pi < p.n_peers; pi++)
                                if (pi != p.self) *reinterpret_cast<float4 *>(p.peers[pi] + row * p.ldc + col) = out;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (col + e < p.N) {
                                float o = v[e];
                                if (p.has_beta) o += p.beta * cp[e];
                                if (p.bias) o += p.bias[col + e];
                                if (p.act) o = apply_act1(o, p.act);
                                cp[e] = o;
                                for (int pi = 0; pi < p.n_peers; pi++)
                                    if (pi != p.self) p.peers[pi][row * p.ldc + col + e] = o;
                            }
                        }
                    }
                }
                if (p.peer_bulk) {
                    // Register stores to a peer hold the warp for an NVLink round trip each (448 per warp and tile with 7
                    // peers: the epilogue took about as long as the tile's MMAs).  Instead every lane hands ITS row of the
                    // finished chunk -- one 128-byte segment -- to the bulk-copy engine once per peer and moves on; only the
                    // shared-memory read has to be over before the staging tile is reused.
                    fence_proxy_async();
                    __syncwarp();
                    const uint64_t row = row0 + lane, colc = col0 + (uint64_t)c * 32;
                    if (row < p.M && colc < p.N) {
                        const uint32_t bytes = (uint32_t)((p.N - colc < 32 ? p.N - colc : 32) * 4);
                        const uint32_t src = smem_u32(epi + lane * EPI_PITCH);
                        for (int pi = 0; pi < p.n_peers; pi++)
                            if (pi != p.self)
                                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                                                 p.peers[pi] + row * p.ldc + colc),
                                             "r"(src), "r"(bytes)
                                             : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncwarp();  // the tile is overwritten by the next chunk
            };

            // Local-only stores (no peer GPUs): every lane writes ITS row of the chunk -- one whole 128-byte line -- straight from
            // the registers tcgen05.ld filled.  The epilogue runs ONE warp per scheduler, so it is bound by instruction latency:
            // the transposing path above costs ~1200 dependent instructions per chunk and warp (2.5 us: 20 us per tile, measured
            // with WK_GEMM_TRACE), this one ~60.  The partial lines a warp instruction leaves are merged in L2 before write-back.
            float *my_c = p.C + (row0 + lane) * p.ldc + col0;
            asm volatile("" : "+l"(my_c));  // (not rematerialised from row * ldc at every use)
            const bool my_row_ok = row0 + lane < p.M;
            auto store_rows = [&](const uint32_t (&r)[32], int c) {
                const uint64_t col = col0 + (uint64_t)c * 32;
                if (!my_row_ok || col >= p.N) return;
                float *cp = my_c + c * 32;
                asm volatile("" : "+l"(cp));  // keep the row pointer in a register pair: the 8 accesses below are [cp + imm]
                if (vec_ok && col + 32 <= p.N) {
#pragma unroll
                    for (int g = 0; g < 8; g++) {
                        float4 v = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]),
                                               __uint_as_float(r[g * 4 + 3]));
                        if (p.has_alpha) {
                            v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha;
                        }
                        if (p.has_beta) {
                            const float4 o = *reinterpret_cast<const float4 *>(cp + g * 4);
                            v.x += p.beta * o.x; v.y += p.beta * o.y; v.z += p.beta * o.z; v.w += p.beta * o.w;
                        }
                        if (p.bias) {
                            const float4 bv = *reinterpret_cast<const float4 *>(p.bias + col + g * 4);
                            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                        }
                        if (p.act) v = apply_act4(v, p.act);
                        *reinterpret_cast<float4 *>(cp + g * 4) = v;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e++) {  // fully unrolled: a dynamic index would put r[] in local memory
                        if (col + e < p.N) {
                            float o = __uint_as_float(r[e]);
                            if (p.has_alpha) o *= p.alpha;
                            if (p.has_beta) o += p.beta * cp[e];
                            if (p.bias) o += p.bias[col + e];
                            cp[e] = p.act ? apply_act1(o, p.act) : o;
                        }
                    }
                }
            };
            // TMA-store path (C has a tensor map, no beta, no peers): the lane puts its row of the chunk into a 128-byte-swizzled
            // 32 x 32 staging box (16-byte piece g of row r at piece g ^ (r & 7): conflict-free) and ONE thread hands the box to
            // the TMA engine, which clips it against C's bounds.  ~40 instructions per chunk and warp, and the global write is
            // asynchronous: two boxes per warp alternate, a box is reused once its previous store has finished READING it.
            // Everything the chunk loop needs is pinned in registers first: with ONE warp per scheduler the epilogue is bound by
            // instruction latency, and a constant-bank read or an address re-derivation in front of every store (what the compiler
            // emits when it may rematerialise) costs 20-40 cycles each -- measured 0.64 us per chunk against 0.1 us like this.
            uint32_t box_lane = smem_u32(epi_raw) + lane * 128;      // this lane's row in staging box 0
            uint32_t swz = (uint32_t)(lane & 7) << 4;                // 16-byte piece g sits at piece g ^ (row & 7)
            uint64_t tmc_addr = reinterpret_cast<uint64_t>(&tmC);
            int f_alpha = p.has_alpha, f_act = p.act;
            float alpha_v = p.alpha;
            const float *bias_p = p.bias;
            asm volatile("" : "+r"(box_lane), "+r"(swz), "+l"(tmc_addr), "+r"(f_alpha), "+r"(f_act), "+f"(alpha_v), "+l"(bias_p));
            const bool plain = !f_alpha && !f_act && bias_p == nullptr;
            const int n_maps = p.peer_tma ? p.n_peers : 0;
            auto store_tma = [&](const uint32_t (&r)[32], int c, int n_chunk) {
                const uint32_t box = box_lane + (n_chunk & 1) * 4096;
                const bool tr = p.trace && c == 1 && q == 0 && lane == 0;
                if (tr) trace_stamp(p, 12);
                if (n_chunk >= 2) {  // the store issued two chunks ago has read this box
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                if (plain) {
#pragma unroll
                    for (int g = 0; g < 8; g++)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + (((uint32_t)g << 4) ^ swz)), "r"(r[g * 4]),
                                     "r"(r[g * 4 + 1]), "r"(r[g * 4 + 2]), "r"(r[g * 4 + 3])
                                     : "memory");
                } else {
                    const uint64_t col = col0 + (uint64_t)c * 32;
#pragma unroll
                    for (int g = 0; g < 8; g++) {
                        float4 v = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]),
                                               __uint_as_float(r[g * 4 + 3]));
                        if (f_alpha) {
                            v.x *= alpha_v; v.y *= alpha_v; v.z *= alpha_v; v.w *= alpha_v;
                        }
                        if (bias_p) {
                            if (col + g * 4 + 4 <= p.N) {
                                const float4 bv = *reinterpret_cast<const float4 *>(bias_p + col + g * 4);
                                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                            } else {  // ragged right edge: the store clips these columns, the bias read must not run past N
                                if (col + g * 4 + 0 < p.N) v.x += bias_p[col + g * 4 + 0];
                                if (col + g * 4 + 1 < p.N) v.y += bias_p[col + g * 4 + 1];
                                if (col + g * 4 + 2 < p.N) v.z += bias_p[col + g * 4 + 2];
                            }
                        }
                        if (f_act) v = apply_act4(v, f_act);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(box + (((uint32_t)g << 4) ^ swz)), "f"(v.x), "f"(v.y),
                                     "f"(v.z), "f"(v.w)
                                     : "memory");
                    }
                }
                if (tr) trace_stamp(p, 13);
                fence_proxy_async();  // generic-proxy writes -> visible to the TMA engine
                __syncwarp();
                if (tr) trace_stamp(p, 14);
                if (lane == 0) {
                    const int32_t ccol = (int32_t)(col0 + (uint64_t)c * 32), crow = (int32_t)row0;
                    if ((uint64_t)ccol < p.N && row0 < p.M) {
                        if (n_maps == 0) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmc_addr),
                                         "r"(ccol), "r"(crow), "r"(box - lane * 128)
                                         : "memory");
                        } else {
                            // fused all-gather: the same staged box goes to every rank's C (own copy first), one TMA store each --
                            // the NVLink transfer is issued by the copy engine while this warp moves on to the next chunk
                            for (int i = 0; i < n_maps; i++) {
                                const int dst = p.self + i < n_maps ? p.self + i : p.self + i - n_maps;
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                                 reinterpret_cast<uint64_t>(&pm.m[dst])),
                                             "r"(ccol), "r"(crow), "r"(box - lane * 128)
                                             : "memory");
                            }
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (tr) trace_stamp(p, 15);
            };
            const bool local_only = p.n_peers == 0;
            const bool use_tma_store = (local_only || p.peer_tma) && p.c_tma && !p.has_beta;
            int n_chunk = 0;

            if (p.splits == 1 || w < p.n_full) {
                if (use_tma_store) {
                    // two register buffers: the TMEM load of chunk c + 1 is in flight while chunk c is staged and stored
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32(taddr, ra);
#pragma unroll 1
                    for (int c = 0; c < BN / 32; c += 2) {
                        tmem_ld_wait();
                        tmem_ld_32x32(taddr + (c + 1) * 32, rb);
                        if (c == 0 && q == 0 && lane == 0) trace_stamp(p, 9);
                        store_tma(ra, c, n_chunk++);
                        if (q == 0 && lane == 0 && c == 0) trace_stamp(p, 10);
                        tmem_ld_wait();
                        if (c + 2 < BN / 32) {
                            tmem_ld_32x32(taddr + (c + 2) * 32, ra);
                        } else {  // accumulator fully read: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) arrive_on_leader<CTAS>(&bars->acc_empty[acc]);
                        }
                        store_tma(rb, c + 1, n_chunk++);
                        if (q == 0 && lane == 0 && c == 2) trace_stamp(p, 11);
                    }
                } else {
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
                        if (local_only) store_rows(r, c);
                        else store_chunk(r, c);
                    }
                }
            } else {
                // split-K: every split of the tile folds ITS share of the tile's 32-column chunks (c = sp, sp + splits, ...).  The
                // chunks the OTHER splits fold are parked in the workspace in register order -- float4 index
                // ((c * 8 + g) * 128 + thread): every warp access covers 512 contiguous bytes and any CTA of the tile reads a
                // chunk back with the same (warp, lane) -> row mapping; a split's own chunks stay in its TMEM accumulator.
                const uint64_t half = (uint64_t)(t - p.n_full) * CTAS + rank;  // which 128-row block of which split tile
                const int et = q * 32 + lane;                                  // 0..127: row of the block
                constexpr uint64_t PART_F4 = (uint64_t)BM * BN / 4;
                uint32_t splits = p.splits;
                float4 *mine = reinterpret_cast<float4 *>(p.ws) + (half * splits + sp) * PART_F4 + et;
                const float4 *part0 = reinterpret_cast<const float4 *>(p.ws) + half * splits * PART_F4 + et;
                unsigned *arrived = p.tickets + 2 * half;
                asm volatile("" : "+l"(mine), "+l"(part0), "+l"(arrived), "+r"(splits));  // pinned: see store_tma
                auto next_parked = [&](uint32_t c) {  // next chunk >= c that another split folds
                    while (c < BN / 32 && c % splits == sp) c++;
                    return c;
                };
                auto park = [&](const uint32_t (&r)[32], uint32_t c) {
                    float4 *mc = mine + c * (8 * 128);
#pragma unroll
                    for (int g = 0; g < 8; g++)
                        __stcg(mc + g * 128, make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]),
                                                         __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3])));
                };
                {  // two register buffers: the TMEM load of the next chunk is in flight while this one is written out
                    uint32_t ra[32], rb[32];
                    uint32_t ca = next_parked(0), cb;
                    if (ca < BN / 32) tmem_ld_32x32(taddr + ca * 32, ra);
#pragma unroll 1
                    while (ca < BN / 32) {
                        tmem_ld_wait();
                        cb = next_parked(ca + 1);
                        if (cb < BN / 32) tmem_ld_32x32(taddr + cb * 32, rb);
                        park(ra, ca);
                        if (cb >= BN / 32) break;
                        tmem_ld_wait();
                        ca = next_parked(cb + 1);
                        if (ca < BN / 32) tmem_ld_32x32(taddr + ca * 32, ra);
                        park(rb, cb);
                    }
                }
                if (q == 0 && lane == 0) trace_stamp(p, 9);
                // publish and wait for the other splits of the tile: all of them are resident (one work item per unit,
                // host-checked), so this is a short spin, not a scheduling dependency.  ONE thread releases for the whole
                // warpgroup (the CTA barrier orders everybody's stores before its release: cumulativity) and acquires for it.
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                if (threadIdx.x == (2 + CONV_WARPS) * 32) {
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(arrived) : "memory");
                    unsigned seen;
                    long long t0 = 0;
                    for (uint32_t spins = 0;; spins++) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrived) : "memory");
                        if (seen >= splits) break;
                        if (spins > 4096) {  // bounded like mbar_wait: a protocol bug traps instead of hanging the GPU
                            if (t0 == 0) t0 = clock64();
                            else if (clock64() - t0 > 6000000000LL) __trap();
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                if (q == 0 && lane == 0) trace_stamp(p, 10);
                // fold: the own partial (from TMEM) first, then the other splits' in ascending order.  Which split folds a chunk
                // is a function of the chunk alone, so the summation order -- and every bit of C -- is the same on every launch.
                // The other splits' partials of my chunks (16 KiB each, contiguous) are fetched by the bulk-copy engine into the
                // pipeline's stage memory, idle now that the accumulator is complete: one thread issues <= 12 copies, everybody
                // waits once and folds from shared memory instead of paying an L2 round trip per half chunk.
                const uint32_t n_mine = (BN / 32 - sp + splits - 1) / splits;  // chunks sp, sp + splits, ...
                if (threadIdx.x == (2 + CONV_WARPS) * 32) {
                    fence_proxy_async_all();  // the partials were written (and acquired) through the generic proxy; the copy engine reads them
                    mbar_arrive_expect_tx(&bars->fold_full, n_mine * (splits - 1) * 16384u);
                    uint32_t slot = 0;
                    for (uint32_t c = sp; c < BN / 32; c += splits)
                        for (uint32_t s2 = 0; s2 < splits; s2++) {
                            if (s2 == sp) continue;
                            const float4 *src = part0 - et + (uint64_t)s2 * PART_F4 + c * (8 * 128);
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                             smem_u32(smem + slot * 16384)),
                                         "l"(src), "r"(16384u), "r"(smem_u32(&bars->fold_full))
                                         : "memory");
                            slot++;
                        }
                }
                mbar_wait(&bars->fold_full, 0);
                {
                    uint32_t slot = 0;
#pragma unroll 1
                    for (uint32_t c = sp; c < BN / 32; c += splits) {
                        uint32_t r[32];
                        tmem_ld_32x32(taddr + c * 32, r);
                        tmem_ld_wait();
#pragma unroll 1
                        for (uint32_t s2 = 0; s2 < splits; s2++) {
                            if (s2 == sp) continue;
                            const float4 *ps = reinterpret_cast<const float4 *>(smem + slot * 16384) + et;
                            slot++;
#pragma unroll
                            for (int g = 0; g < 8; g++) {
                                const float4 v = ps[g * 128];
                                r[g * 4] = __float_as_uint(__uint_as_float(r[g * 4]) + v.x);
                                r[g * 4 + 1] = __float_as_uint(__uint_as_float(r[g * 4 + 1]) + v.y);
                                r[g * 4 + 2] = __float_as_uint(__uint_as_float(r[g * 4 + 2]) + v.z);
                                r[g * 4 + 3] = __float_as_uint(__uint_as_float(r[g * 4 + 3]) + v.w);
                            }
                        }
                        if (use_tma_store) store_tma(r, (int)c, n_chunk++);
                        else if (local_only) store_rows(r, (int)c);
                        else store_chunk(r, (int)c);
                        if (q == 0 && lane == 0 && c == sp) trace_stamp(p, 11);
                    }
                }
                // accumulator fully read: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_on_leader<CTAS>(&bars->acc_empty[acc]);
                unsigned *done = arrived + 1;
                // the last split to finish re-arms the counters for the next launch
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                if (threadIdx.x == (2 + CONV_WARPS) * 32) {
                    if (atomicAdd(done, 1u) == p.splits - 1) {
                        *arrived = 0;
                        *done = 0;
                    }
                }
            }
        }
        // bulk copies to the peers / TMA stores of C: complete (written, not only read) before this CTA is allowed to finish
        if (p.peer_bulk || p.c_tma) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (q == 0 && lane == 0) trace_stamp(p, 7);
    }

    tc_fence_before();
    if (CTAS == 1) __syncthreads();
    else cluster_sync_all();  // no CTA of a pair may exit (or free TMEM) while its partner can still touch it
    if (threadIdx.x == 0) trace_stamp(p, 8);
    if (warp == 1) {
        tc_fence_after();
        if (CTAS == 1) tmem_dealloc(tmem_base, TMEM_COLS);
        else tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    }
}

// lo plane of a stored operand: lo = rna_tf32(x - trunc_tf32(x)), the converter warps' arithmetic (split mode 1), one pass
// over HBM.  Rows of `cols` floats at pitch ld (a multiple of 4, 16-byte aligned rows): float4 accesses may run into the
// row's padding but never past the pitch.
__global__ void __launch_bounds__(256) split_lo_kernel(const float *__restrict__ x, float *__restrict__ lo, uint64_t rows,
                                                       uint64_t cols, uint64_t ld) {
    const uint64_t c4n = (cols + 3) / 4, total = rows * c4n;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (uint64_t)gridDim.x * 256) {
        const uint64_t r = i / c4n, c = (i - r * c4n) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * ld + c));
        const float xs[4] = {v.x, v.y, v.z, v.w};
        float ls[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t hb = __float_as_uint(xs[e]) & 0xffffe000u;
            const float rem = (hb & 0x7f800000u) == 0x7f800000u ? 0.0f : xs[e] - __uint_as_float(hb);
            uint32_t lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
            ls[e] = __uint_as_float(lb);
        }
        *reinterpret_cast<float4 *>(lo + r * ld + c) = make_float4(ls[0], ls[1], ls[2], ls[3]);
    }
}

}  // namespace tc

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

int32_t gemm_f32_tc(wk_queue *q, int32_t op_a, int32_t op_b, uint64_t M, uint64_t N, uint64_t K, const float *alpha,
                    const float *A, uint64_t lda, const float *B, uint64_t ldb, const float *beta, float *C, uint64_t ldc,
                    const float *bias, int32_t act, const GemmPeers *peers, const GemmProlog *prolog) {
    using namespace tc;
    if (prolog) {
        if (!aligned16(prolog->Y) || (prolog->ldy % 4) || (prolog->colsum && op_a != 1) || (peers && peers->n > 1)) return -1;
        if (prolog->act != WK_ACT_SIGMOID && prolog->act != WK_ACT_TANH) return -1;
    }
    // TMA needs 16-byte aligned bases and pitches; the epilogue wants 16-byte aligned C rows
    // (gemm.cu stages unaligned A / B into an aligned scratch first; an unaligned C takes the epilogue's element-wise stores)
    if (!aligned16(A) || !aligned16(B) || (lda % 4) || (ldb % 4)) return -1;
    if (bias && !aligned16(bias)) return -1;
    if (q->prop.major != 10) return -1;
    if (M > 0x7fffffffULL || N > 0x7fffffffULL || K > 0x7fffffffULL) return -1;
    if (peers)
        for (int i = 0; i < peers->n; i++)
            if (!aligned16(peers->ptrs[i])) return -1;

    // CTA pairs (256 x 256 tiles) when the problem has at least two row tiles per pair to share; WK_GEMM_CTAS overrides
    static const int ctas_env = env_int("WK_GEMM_CTAS", 0);
    static const int split_env = env_int("WK_GEMM_SPLIT", 1);
    // (the prologue variant exists for CTA pairs only: a single CTA's 128 x 256 stage plus the Y tile leaves room for one stage)
    const int ctas = prolog ? 2 : (ctas_env == 1 || ctas_env == 2 ? ctas_env : (M > BM ? 2 : 1));
    const uint32_t bn_load = BN / ctas;

    CUtensorMap tmA, tmB;
    bool ok;
    if (op_a == 0) ok = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, A, K, M, lda * 4, BK, BM);
    else ok = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, A, M, K, lda * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (op_b == 1) ok = ok && make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, B, K, N, ldb * 4, BK, bn_load);
    else ok = ok && make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, B, N, K, ldb * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (!ok) {
        set_error("gemm_f32_tc: cuTensorMapEncodeTiled failed");
        return WK_ERR_CUDA;
    }
    // MN-major operands whose MN extent is a multiple of 32: one 3-D box per stage instead of BM / 32 (+ BN_LOAD / 32) 2-D ones
    static const int mn3d_env = env_int("WK_GEMM_MN3D", 1);
    int a_mn3d = 0, b_mn3d = 0;
    if (mn3d_env && !prolog) {
        CUtensorMap t3;
        if (op_a == 1 && M % 32 == 0 && make_tmap_mn_chunks(&t3, A, M, K, lda * 4, BK, BM / 32)) { tmA = t3; a_mn3d = 1; }
        if (op_b == 0 && N % 32 == 0 && make_tmap_mn_chunks(&t3, B, N, K, ldb * 4, BK, bn_load / 32)) { tmB = t3; b_mn3d = 1; }
    }
    // WK_GEMM_PRESPLIT=1 (experimental, off by default): lo planes computed in one streaming pass per operand and loaded by
    // TMA, so the converter warps move no shared memory.  Costs an operand-sized workspace and 2x the L2 -> SM traffic.
    static const int presplit_env = env_int("WK_GEMM_PRESPLIT", 0);
    const bool presplit = presplit_env && split_env == 1 && !prolog;
    CUtensorMap tmAlo = tmA, tmBlo = tmB;
    if (presplit) {
        const uint64_t a_rows = op_a ? K : M, a_cols = op_a ? M : K, b_rows = op_b ? N : K, b_cols = op_b ? K : N;
        const size_t a_bytes = (a_rows * lda * 4 + 255) & ~(size_t)255, b_bytes = b_rows * ldb * 4;
        if (q->presplit_bytes < a_bytes + b_bytes) {
            int32_t rc = grow_buffer(q, &q->presplit_ws, &q->presplit_bytes, a_bytes + b_bytes);
            if (rc != WK_OK) return rc;
        }
        float *a_lo = (float *)q->presplit_ws, *b_lo = (float *)((char *)q->presplit_ws + a_bytes);
        const unsigned grid = (unsigned)q->sm_count * 8;
        split_lo_kernel<<<grid, 256, 0, q->stream>>>(A, a_lo, a_rows, a_cols, lda);
        WK_CHECK_LAUNCH();
        split_lo_kernel<<<grid, 256, 0, q->stream>>>(B, b_lo, b_rows, b_cols, ldb);
        WK_CHECK_LAUNCH();
        if (op_a == 0) ok = make_tmap_2d(&tmAlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a_lo, K, M, lda * 4, BK, BM);
        else ok = make_tmap_2d(&tmAlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a_lo, M, K, lda * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (op_b == 1) ok = ok && make_tmap_2d(&tmBlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, b_lo, K, N, ldb * 4, BK, bn_load);
        else ok = ok && make_tmap_2d(&tmBlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, b_lo, N, K, ldb * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (!ok) {
            set_error("gemm_f32_tc: cuTensorMapEncodeTiled failed (lo planes)");
            return WK_ERR_CUDA;
        }
    }

    if (prolog) {  // Y travels in the tmAlo slot: A's geometry over Y's memory
        if (op_a == 0) ok = make_tmap_2d(&tmAlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, prolog->Y, K, M, prolog->ldy * 4, BK, BM);
        else ok = make_tmap_2d(&tmAlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, prolog->Y, M, K, prolog->ldy * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (!ok) {
            set_error("gemm_f32_tc: cuTensorMapEncodeTiled failed (prologue operand)");
            return WK_ERR_CUDA;
        }
    }
    Params p{};
    p.prol_act = prolog ? prolog->act : 0;
    p.colsum = prolog ? prolog->colsum : nullptr;
    p.C = C;
    p.M = M; p.N = N; p.K = K; p.ldc = ldc;
    p.c_vec = aligned16(C) && ldc % 4 == 0;
    static const int tma_store_env = env_int("WK_GEMM_TMA_STORE", 1);
    CUtensorMap tmC = tmA;
    p.c_tma = 0;
    static PeerMaps pm_zero{};
    PeerMaps pm = pm_zero;
    p.peer_tma = 0;
    if (p.c_vec && tma_store_env && !(peers && peers->n > 1)) {
        // finished 32 x 32 chunks of C: 128-byte rows, swizzled staging box, clipped against [M, N] by the engine
        p.c_tma = make_tmap_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, C, N, M, ldc * 4, 32, 32) ? 1 : 0;
    } else if (p.c_vec && tma_store_env && peers && peers->n <= MAX_PEER_MAPS) {
        // measured on 8 B200s in one session (profiles/bench_r02d_g8*.json): 1817 TF/s with one TMA store per (chunk, rank) against
        // 1866 with the per-row bulk copies of the transposing path (compute-only 2006-2054) -- not a gain, so it stays opt-in
        static const int peer_tma_env = env_int("WK_GEMM_PEER_TMA", 0);
        bool all = peer_tma_env != 0;
        for (int i = 0; i < peers->n && all; i++)
            all = make_tmap_2d(&pm.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, peers->ptrs[i], N, M, ldc * 4, 32, 32);
        if (all) p.c_tma = p.peer_tma = 1;
    }
    p.has_alpha = (alpha != nullptr || beta != nullptr);
    p.has_beta = (beta != nullptr);
    p.alpha = alpha ? *alpha : 1.0f;
    p.beta = beta ? *beta : 0.0f;
    p.bias = bias;
    p.act = act;
    p.op_a = op_a; p.op_b = op_b;
    p.a_mn3d = a_mn3d; p.b_mn3d = b_mn3d;
    p.split = split_env;
    p.tiles_m = (uint32_t)((M + (uint64_t)BM * ctas - 1) / ((uint64_t)BM * ctas));
    p.tiles_n = (uint32_t)((N + BN - 1) / BN);
    static const int gm_env = env_int("WK_GEMM_GM", 16);
    p.group_m = gm_env > 0 ? gm_env : 16;
    // split-K when the output tiles alone leave more than half of the chip idle (small M*N, long K): at least 4
    // k-blocks per split, at most 8 splits; WK_GEMM_SPLITK=1 disables, =n forces n
    const uint64_t n_out_tiles = (uint64_t)p.tiles_m * p.tiles_n;
    static bool attr_set[64] = {false};
    if (!attr_set[q->device & 63]) {
        WK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES));
        WK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2, true>::SMEM_BYTES));
        attr_set[q->device & 63] = true;
    }
    static int max_pairs[64] = {0};  // co-resident CTA pairs (GPCs with an odd SM count strand one SM), per device
    if (ctas == 2 && max_pairs[q->device & 63] == 0) {
        cudaLaunchConfig_t occ{};
        occ.gridDim = dim3(2 * (unsigned)q->sm_count);
        occ.blockDim = dim3(THREADS);
        occ.dynamicSmemBytes = Cfg<2>::SMEM_BYTES;
        cudaLaunchAttribute oa[1];
        oa[0].id = cudaLaunchAttributeClusterDimension;
        oa[0].val.clusterDim.x = 2;
        oa[0].val.clusterDim.y = 1;
        oa[0].val.clusterDim.z = 1;
        occ.attrs = oa;
        occ.numAttrs = 1;
        int n_cl = 0;
        if (cudaOccupancyMaxActiveClusters(&n_cl, gemm_tf32x3_kernel<2, false>, &occ) != cudaSuccess || n_cl <= 0) {
            cudaGetLastError();
            n_cl = q->sm_count / 2 - 4;
        }
        max_pairs[q->device & 63] = n_cl;
    }
    const uint64_t units_avail = ctas == 2 ? (uint64_t)max_pairs[q->device & 63] : (uint64_t)q->sm_count;
    const uint32_t num_kb = (uint32_t)((K + BK - 1) / BK);
    static const int splitk_env = env_int("WK_GEMM_SPLITK", 0);   // 1: never split, n: force n splits (whole problem)
    static const int tail_env = env_int("WK_GEMM_TAILSPLIT", 1);  // 0: no tail split
    static const int debug_env = env_int("WK_DEBUG", 0);
    uint32_t splits = 1;
    uint64_t n_full = 0;  // unsplit tiles
    if (prolog) splits = 1;  // (the prologue variant keeps one k-range per tile: its column sums cover all of K)
    else if (splitk_env > 0) splits = (uint32_t)splitk_env;
    else if (n_out_tiles * 2 <= units_avail) splits = (uint32_t)(units_avail / n_out_tiles);
    else if (tail_env && num_kb >= 16 && !prolog) {
        // wave quantisation: the persistent schedule runs ceil(tiles / units) rounds; when the last round is at most half
        // full (N = 4096: 256 tiles = 3 rounds of 74 + 34), its tiles are split along K so the round costs 1/splits
        const uint64_t rem = n_out_tiles % units_avail;
        if (rem > 0 && rem * 2 <= units_avail) {
            n_full = n_out_tiles - rem;
            splits = (uint32_t)(units_avail / rem);
            if (splits > 4) splits = 4;
        }
    }
    const uint64_t n_split_tiles = n_out_tiles - n_full;
    if (splits > 8) splits = 8;
    if (splits > num_kb / 4) splits = num_kb / 4;
    if (n_split_tiles * splits > units_avail) splits = (uint32_t)(units_avail / n_split_tiles);  // co-residency of a tile's splits
    if (splits < 1) splits = 1;
    p.kb_per_split = (num_kb + splits - 1) / splits;
    p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
    if (p.splits == 1) n_full = 0;
    p.n_full = (uint32_t)n_full;
    p.ws = nullptr;
    p.tickets = nullptr;
    if (p.splits > 1) {
        const size_t halves = (size_t)n_split_tiles * ctas;
        const size_t ws_bytes = halves * p.splits * BM * BN * sizeof(float);
        int32_t rc = ensure_splitk(q, ws_bytes, 2 * halves);
        if (rc != WK_OK) return rc;
        p.ws = (float *)q->splitk_ws;
        p.tickets = q->splitk_tickets;
    }
    if (debug_env)
        fprintf(stderr, "[wk] gemm_f32_tc %llux%llux%llu ctas=%d tiles=%llu units_avail=%llu n_full=%llu splits=%u kb/split=%u\n",
                (unsigned long long)M, (unsigned long long)N, (unsigned long long)K, ctas, (unsigned long long)n_out_tiles,
                (unsigned long long)units_avail, (unsigned long long)n_full, p.splits, p.kb_per_split);
    p.n_peers = 0;
    p.self = 0;
    p.peer_bulk = 0;
    if (peers && peers->n > 1) {
        p.n_peers = peers->n;
        p.self = peers->self;
        for (int i = 0; i < peers->n; i++) p.peers[i] = (float *)peers->ptrs[i];
        static const int bulk_env = env_int("WK_GEMM_PEER_BULK", 1);
        p.peer_bulk = bulk_env && peers->n > 1 && N % 4 == 0 && p.c_vec;  // whole 16-byte pieces per row segment
    }

    static const int trace_env = env_int("WK_GEMM_TRACE", 0);
    static unsigned long long *trace_buf = nullptr;
    p.trace = nullptr;
    if (trace_env) {
        if (!trace_buf) WK_CUDA(cudaMalloc((void **)&trace_buf, 512 * 16 * sizeof(unsigned long long)));
        WK_CUDA(cudaMemsetAsync(trace_buf, 0, 512 * 16 * sizeof(unsigned long long), q->stream));
        p.trace = trace_buf;
    }
    const uint64_t num_tiles = n_full + n_split_tiles * p.splits;
    // split launches are cooperative: the grid must be what can be resident at once
    const uint64_t max_units = p.splits > 1 ? units_avail : (uint64_t)q->sm_count / ctas;
    const unsigned units = (unsigned)(num_tiles < max_units ? num_tiles : max_units);
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(units * ctas);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = prolog ? Cfg<2, true>::SMEM_BYTES : (ctas == 1 ? Cfg<1>::SMEM_BYTES : Cfg<2>::SMEM_BYTES);
        cfg.stream = q->stream;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (ctas == 2) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = 2;
            attr[na].val.clusterDim.y = 1;
            attr[na].val.clusterDim.z = 1;
            na++;
        }
        static const int coop_env = env_int("WK_GEMM_SPLITK_COOP", 1);  // 0: plain launch (ncu cannot replay cooperative cluster launches)
        if (p.splits > 1 && coop_env) {
            // the splits of a tile wait for each other inside the kernel: a cooperative launch makes the driver place the
            // whole grid at once (or not at all), so that wait can never depend on a CTA that has no SM yet -- also
            // when another queue's kernel holds part of the chip
            attr[na].id = cudaLaunchAttributeCooperative;
            attr[na].val.cooperative = 1;
            na++;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        if (prolog) {
            WK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<2, false, true>, tmA, tmB, tmAlo, tmBlo, tmC, pm, p));
        } else if (presplit) {
            if (ctas == 1) WK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<1, true>, tmA, tmB, tmAlo, tmBlo, tmC, pm, p));
            else WK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<2, true>, tmA, tmB, tmAlo, tmBlo, tmC, pm, p));
        } else {
            if (ctas == 1) WK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<1, false>, tmA, tmB, tmAlo, tmBlo, tmC, pm, p));
            else WK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<2, false>, tmA, tmB, tmAlo, tmBlo, tmC, pm, p));
        }
    }
    WK_CHECK_LAUNCH();
    if (trace_env) {  // debugging aid: per-CTA milestones in ns relative to the earliest kernel entry
        static unsigned long long host[512 * 16];
        WK_CUDA(cudaStreamSynchronize(q->stream));
        WK_CUDA(cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost));
        const unsigned n_cta = units * ctas < 512 ? units * ctas : 512;
        unsigned long long t0 = ~0ull, t_end = 0;
        for (unsigned c = 0; c < n_cta; c++) {
            if (host[c * 16] && host[c * 16] < t0) t0 = host[c * 16];
            if (host[c * 16 + 8] > t_end) t_end = host[c * 16 + 8];
        }
        fprintf(stderr, "[wk trace] %llux%llux%llu ctas=%d units=%u splits=%u: kernel span %.2f us; per CTA (us since first entry): entry setup tma0 raw0 mma0 mmaL accF epiD exit ld0|parked st0|synced st3|fold1 | chunk1: begin sts fence tma\n",
                (unsigned long long)M, (unsigned long long)N, (unsigned long long)K, ctas, units, p.splits, (t_end - t0) * 1e-3);
        for (unsigned c = 0; c < n_cta; c += (n_cta > 8 ? n_cta / 4 : 1)) {
            fprintf(stderr, "[wk trace]  cta %3u:", c);
            for (int s2 = 0; s2 <= 15; s2++) fprintf(stderr, " %6.2f", host[c * 16 + s2] ? (host[c * 16 + s2] - t0) * 1e-3 : -1.0);
            fprintf(stderr, "\n");
        }
    }
    return WK_OK;
}

}  // namespace wk
