// reduce.cu -- reductions of the hot path: math.sum (src/math/basic.zig:131-203 + kernels/sum.cl), the
// column sum of Linear.getBiasSensitivity (src/nn/layer/linear.zig:534-577 + bias_step.cl) and the BLAS-style
// reduction dot named by the north star.  Deterministic, no floating-point atomics: a grid-strided 128-bit streaming
// read with warp-shuffle + shared-memory block reduction into per-block partials; for sum / dot the block that arrives
// last (an integer ticket) folds the partials in index order inside the same launch, for the column sums a second small
// kernel does.  The scalar of the blocking forms travels through pinned memory, and the host polls a sequence number
// posted beside it (the reference blocks on a mapped read at the same point, basic.zig:154-171).
#include <atomic>

#include "common.cuh"

namespace wk {

constexpr int kRThreads = 256;

template <typename A> __device__ __forceinline__ A shfl_xor(A v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ uint64_t shfl_xor<uint64_t>(uint64_t v, int m) {
    return (uint64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)v, m);
}
// complex: the two components reduce independently (sum.cl:17-24)
#define WK_SHFL_CX(AT)                                                                         \
    template <> __device__ __forceinline__ CxAcc<AT> shfl_xor<CxAcc<AT>>(CxAcc<AT> v, int m) { \
        return CxAcc<AT>{shfl_xor<AT>(v.re, m), shfl_xor<AT>(v.im, m)};                        \
    }
WK_SHFL_CX(uint32_t)
WK_SHFL_CX(uint64_t)
WK_SHFL_CX(float)
WK_SHFL_CX(double)
#undef WK_SHFL_CX
template <typename A> struct AccZero { __device__ __forceinline__ static A get() { return (A)0; } };
template <typename AB> struct AccZero<CxAcc<AB>> { __device__ __forceinline__ static CxAcc<AB> get() { return CxAcc<AB>{(AB)0, (AB)0}; } };

template <typename A> __device__ __forceinline__ A block_reduce_sum(A v) {
    __shared__ A warp_part[kRThreads / 32];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = v + shfl_xor<A>(v, m);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    A r = AccZero<A>::get();
    if (warp == 0) {
        r = lane < kRThreads / 32 ? warp_part[lane] : AccZero<A>::get();
#pragma unroll
        for (int m = 4; m >= 1; m >>= 1) r = r + shfl_xor<A>(r, m);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

template <typename A> __device__ __forceinline__ A ld_cg_acc(const A *p) { return __ldcg(p); }  // L2: written by other SMs
template <typename AB> __device__ __forceinline__ CxAcc<AB> ld_cg_acc(const CxAcc<AB> *p) {
    return CxAcc<AB>{__ldcg(&p->re), __ldcg(&p->im)};
}

// Runs: n_runs contiguous runs of run_len elements; run r starts at (r / rows) * sp + (r % rows) * rp.
// NIN = 1: sum of x; NIN = 2: sum of x*y (y has its own pitches).
template <typename T, int NIN, bool VECTOR>
__global__ void __launch_bounds__(kRThreads) reduce_runs_kernel(const T *__restrict__ x, const T *__restrict__ y,
                                                                uint64_t run_len, uint64_t n_runs, uint64_t rows,
                                                                uint64_t xrp, uint64_t xsp, uint64_t yrp, uint64_t ysp,
                                                                typename Acc<T>::type *partial, unsigned *ticket,
                                                                T *__restrict__ out, unsigned *posted, unsigned seq) {
    using A = typename Acc<T>::type;
    constexpr int VEC = 16 / (int)sizeof(T);
    union Pack { uint4 u; T e[VEC]; };
    A acc[4] = {AccZero<A>::get(), AccZero<A>::get(), AccZero<A>::get(), AccZero<A>::get()};
    for (uint64_t r = blockIdx.y; r < n_runs; r += gridDim.y) {
        const uint64_t d = r / rows, j = r - d * rows;
        const T *xr = x + d * xsp + j * xrp;
        const T *yr = NIN == 2 ? y + d * ysp + j * yrp : nullptr;
        if (VECTOR) {
            const uint64_t nv = run_len / VEC;
            const uint4 *xv = reinterpret_cast<const uint4 *>(xr);
            const uint4 *yv = reinterpret_cast<const uint4 *>(yr);
            const uint64_t stride = (uint64_t)gridDim.x * kRThreads;
            uint64_t i = (uint64_t)blockIdx.x * kRThreads + threadIdx.x;
            // 4 independent 128-bit loads in flight per operand
            for (; i + 3 * stride < nv; i += 4 * stride) {
                Pack px[4], py[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    px[u].u = __ldg(xv + i + u * stride);
                    if (NIN == 2) py[u].u = __ldg(yv + i + u * stride);
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int e = 0; e < VEC; e++)
                        acc[u] = acc[u] + (NIN == 2 ? to_acc<T>(px[u].e[e]) * to_acc<T>(py[u].e[e]) : to_acc<T>(px[u].e[e]));
            }
            for (; i < nv; i += stride) {
                Pack px, py;
                px.u = __ldg(xv + i);
                if (NIN == 2) py.u = __ldg(yv + i);
#pragma unroll
                for (int e = 0; e < VEC; e++)
                    acc[0] = acc[0] + (NIN == 2 ? to_acc<T>(px.e[e]) * to_acc<T>(py.e[e]) : to_acc<T>(px.e[e]));
            }
        } else {
            for (uint64_t i = (uint64_t)blockIdx.x * kRThreads + threadIdx.x; i < run_len; i += (uint64_t)gridDim.x * kRThreads)
                acc[0] = acc[0] + (NIN == 2 ? to_acc<T>(xr[i]) * to_acc<T>(yr[i]) : to_acc<T>(xr[i]));
        }
    }
    A v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    v = block_reduce_sum<A>(v);
    // One launch: the block that arrives last folds the partials -- in index order, by the same thread/shuffle tree whoever
    // it is, so the sum does not depend on which block that was -- and re-arms the ticket for the next launch.
    __shared__ bool is_last;
    const unsigned n_part = gridDim.x * gridDim.y;
    if (threadIdx.x == 0) {
        partial[(uint64_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == n_part - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    A f = AccZero<A>::get();
    for (unsigned i = threadIdx.x; i < n_part; i += kRThreads) f = f + ld_cg_acc(partial + i);
    f = block_reduce_sum<A>(f);
    if (threadIdx.x == 0) {
        *out = from_acc<T>(f);
        *ticket = 0;
        if (posted) {  // blocking form: `out` is pinned host memory; tell the polling host thread the scalar has landed
            __threadfence_system();
            *reinterpret_cast<volatile unsigned *>(posted) = seq;
        }
    }
}

// Dense, vectorisable, LARGE inputs: the same one-launch reduction, but the unit of work is a 512 KiB chunk that persistent
// CTAs fetch from a counter as they go, and the partial sums are indexed by CHUNK, not by CTA.  A statically partitioned
// persistent grid ends with stragglers (the SMs do not get equal shares of the HBM bandwidth: the streaming maps gained 4-9 %
// from finer scheduling, profiles/sweep_stream_grid_r02.md); with chunk-indexed partials the value still does not depend on
// which CTA summed which chunk -- every chunk is summed by the same thread/shuffle tree, the last CTA folds the chunk partials
// in index order.
constexpr int kDynChunkVecs = 32768;  // 16-byte vectors per chunk (512 KiB): 128 per thread
template <typename T, int NIN>
__global__ void __launch_bounds__(kRThreads) reduce_dense_dyn_kernel(const T *__restrict__ x, const T *__restrict__ y, uint64_t nv,
                                                                     unsigned n_chunks, typename Acc<T>::type *partial,
                                                                     unsigned *ticket, T *__restrict__ out, unsigned *posted,
                                                                     unsigned seq) {
    using A = typename Acc<T>::type;
    constexpr int VEC = 16 / (int)sizeof(T);
    union Pack { uint4 u; T e[VEC]; };
    const uint4 *xv = reinterpret_cast<const uint4 *>(x);
    const uint4 *yv = reinterpret_cast<const uint4 *>(y);
    __shared__ unsigned s_next;
    unsigned c = blockIdx.x;  // the first chunk is static, the rest come from ticket[1]
    while (c < n_chunks) {
        if (threadIdx.x == 0) s_next = gridDim.x + atomicAdd(ticket + 1, 1u);
        A acc[4] = {AccZero<A>::get(), AccZero<A>::get(), AccZero<A>::get(), AccZero<A>::get()};
        const uint64_t v0 = (uint64_t)c * kDynChunkVecs + threadIdx.x;
#pragma unroll 2
        for (int k = 0; k < kDynChunkVecs / kRThreads; k += 4) {
            Pack px[4], py[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint64_t i = v0 + (uint64_t)(k + u) * kRThreads;
                px[u].u = make_uint4(0u, 0u, 0u, 0u);
                if (NIN == 2) py[u].u = make_uint4(0u, 0u, 0u, 0u);
                if (i < nv) {
                    px[u].u = __ldg(xv + i);
                    if (NIN == 2) py[u].u = __ldg(yv + i);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int e = 0; e < VEC; e++)
                    acc[u] = acc[u] + (NIN == 2 ? to_acc<T>(px[u].e[e]) * to_acc<T>(py[u].e[e]) : to_acc<T>(px[u].e[e]));
        }
        A v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        v = block_reduce_sum<A>(v);
        if (threadIdx.x == 0) partial[c] = v;
        c = s_next;
        __syncthreads();  // everybody has read s_next before thread 0 draws the next one
    }
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    A f = AccZero<A>::get();
    for (unsigned i = threadIdx.x; i < n_chunks; i += kRThreads) f = f + ld_cg_acc(partial + i);
    f = block_reduce_sum<A>(f);
    if (threadIdx.x == 0) {
        *out = from_acc<T>(f);
        ticket[0] = 0;
        ticket[1] = 0;
        if (posted) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned *>(posted) = seq;
        }
    }
}

template <typename T, int NIN>
static int32_t reduce_runs(wk_queue *q, const T *x, const T *y, uint64_t depth, uint64_t rows, uint64_t run_len, uint64_t xrp,
                           uint64_t xsp, uint64_t yrp, uint64_t ysp, void *host_out, void *device_out = nullptr) {
    using A = typename Acc<T>::type;
    if (!x || (NIN == 2 && !y) || (!host_out && !device_out)) return WK_ERR_INVALID_BUFFER;
    if (depth == 0 || rows == 0 || run_len == 0) return WK_ERR_INVALID_VALUE;
    // collapse contiguous runs
    if (rows > 1 && xrp == run_len && (NIN == 1 || yrp == run_len)) { run_len *= rows; rows = 1; }
    if (rows == 1 && depth > 1 && xsp == run_len && (NIN == 1 || ysp == run_len)) { run_len *= depth; depth = 1; }
    const uint64_t n_runs = depth * rows;
    constexpr int VEC = 16 / (int)sizeof(T);
    bool vec = aligned16(x) && run_len % VEC == 0 && (n_runs == 1 || (xrp % VEC == 0 && xsp % VEC == 0));
    if (NIN == 2) vec = vec && aligned16(y) && (n_runs == 1 || (yrp % VEC == 0 && ysp % VEC == 0));
    const uint64_t per_block = (uint64_t)kRThreads * (vec ? VEC * 4 : 4);
    uint64_t gx = (run_len + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t)q->sm_count * 8;
    if (gx > cap) gx = cap;
    uint64_t gy = n_runs;
    const uint64_t cap_y = (cap + gx - 1) / gx;
    if (gy > cap_y) gy = cap_y;
    if (gy > 65535) gy = 65535;
    const uint64_t n_part = gx * gy;
    int32_t rc = ensure_scratch(q, n_part * sizeof(A) + 64);
    if (rc != WK_OK) return rc;
    A *partial = reinterpret_cast<A *>((char *)q->scratch + 64);
    // blocking form: the fold kernel stores the scalar straight into the queue's pinned host word (device-addressable
    // under UVA), so the only thing after it is the stream synchronisation the reference's mapped read implies
    T *result = device_out ? reinterpret_cast<T *>(device_out) : reinterpret_cast<T *>(q->pinned);
    unsigned *posted = device_out ? nullptr : reinterpret_cast<unsigned *>((char *)q->pinned + 128);
    const unsigned seq = device_out ? 0 : ++q->reduce_seq;
    dim3 grid((unsigned)gx, (unsigned)gy);
    const uint64_t nv = run_len / VEC, n_chunks = (nv + kDynChunkVecs - 1) / kDynChunkVecs;
    static const int dyn_env = getenv("WK_REDUCE_DYNAMIC") ? atoi(getenv("WK_REDUCE_DYNAMIC")) : 1;
    if (dyn_env && vec && n_runs == 1 && n_chunks >= 2 * (uint64_t)q->sm_count && n_chunks < 0x7fffffffull) {
        rc = ensure_scratch(q, n_chunks * sizeof(A) + 64);
        if (rc != WK_OK) return rc;
        partial = reinterpret_cast<A *>((char *)q->scratch + 64);
        reduce_dense_dyn_kernel<T, NIN><<<(unsigned)cap, kRThreads, 0, q->stream>>>(x, y, nv, (unsigned)n_chunks, partial, q->reduce_ticket,
                                                                                   result, posted, seq);
    } else if (vec)
        reduce_runs_kernel<T, NIN, true><<<grid, kRThreads, 0, q->stream>>>(x, y, run_len, n_runs, rows, xrp, xsp, yrp, ysp, partial,
                                                                           q->reduce_ticket, result, posted, seq);
    else
        reduce_runs_kernel<T, NIN, false><<<grid, kRThreads, 0, q->stream>>>(x, y, run_len, n_runs, rows, xrp, xsp, yrp, ysp, partial,
                                                                            q->reduce_ticket, result, posted, seq);
    WK_CHECK_LAUNCH();
    if (device_out) return WK_OK;  // async form: the scalar stays on the device, stream-ordered
    // The reference blocks on a mapped read here (basic.zig:154-171).  Waking up from cudaStreamSynchronize costs 5-8 us
    // after the kernel has finished; the last block posts a sequence number next to the scalar instead and this thread
    // polls the pinned word (checking the stream every ~50 us so that a failed launch surfaces as an error, not a hang).
    volatile unsigned *flag = reinterpret_cast<volatile unsigned *>((char *)q->pinned + 128);
    for (unsigned spins = 0; *flag != seq; spins++) {
        if ((spins & 0x3ff) == 0x3ff) {
            const cudaError_t st = cudaStreamQuery(q->stream);
            if (st == cudaSuccess) break;  // everything enqueued has run: the scalar is there
            if (st != cudaErrorNotReady) return cuda_fail(st, "cudaStreamQuery (blocking reduction)", __FILE__, __LINE__);
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(host_out, q->pinned, sizeof(T));
    return WK_OK;
}

// bias_step.cl:24-37.  Stage 1: thread (c, chunk) sums rows [chunk*rows_per, ...) of column c in ascending row
// order (coalesced across c); stage 2 folds the chunks in ascending order.  With one chunk the summation order
// is exactly the reference's.
// VECTOR: a thread owns one 128-bit group of columns (row pitch a multiple of the vector width, aligned rows) and keeps
// 8 row vectors in flight; per column the additions are the same ones in the same order.
// ACT != 0 (fused first step of Linear.backward, linear.zig:608-613 + :534-577 in ONE pass): the summed value is
// s * act'(y) -- the arithmetic of ActBackwardF, no FMA contraction in this translation unit -- and it is also written back
// over s, so the two GEMMs that follow read the finished sensitivity.
template <typename T, bool VECTOR, int ACT = 0>
__global__ void __launch_bounds__(256) colsum_stage1(const T *__restrict__ s, uint64_t rp, uint64_t rows, uint64_t n_cols,
                                                     uint64_t rows_per, T *__restrict__ out, uint64_t out_pitch,
                                                     const T *__restrict__ y = nullptr, uint64_t y_rp = 0, T *s_out = nullptr) {
    constexpr int VEC = VECTOR ? 16 / (int)sizeof(T) : 1;
    union Pack { uint4 u; T e[16 / sizeof(T)]; };
    const uint64_t c = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * VEC;
    if (c >= n_cols) return;
    const uint64_t r0 = (uint64_t)blockIdx.y * rows_per;
    uint64_t r1 = r0 + rows_per;
    if (r1 > rows) r1 = rows;
    T acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e++) acc[e] = (T)0;
    const T *p = s + r0 * rp + c;
    auto load = [](const T *q) {
        Pack v;
        if (VECTOR) v.u = ACT ? *reinterpret_cast<const uint4 *>(q) : __ldg(reinterpret_cast<const uint4 *>(q));  // (ACT: s is rewritten)
        else v.e[0] = *q;
        return v;
    };
    auto finish = [&](Pack &a, uint64_t row) {  // ACT: a = s * act'(y), stored back
        if (ACT) {
            const Pack yv = load(y + row * y_rp + c);
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const T yy = yv.e[e];
                const T d = ACT == WK_ACT_SIGMOID ? yy * ((T)1 - yy) : (T)1 - yy * yy;
                a.e[e] = a.e[e] * d;
            }
            T *o = s_out + row * rp + c;
            if (VECTOR) *reinterpret_cast<uint4 *>(o) = a.u;
            else *o = a.e[0];
        }
    };
    uint64_t r = r0;
    constexpr int FLY = ACT ? 4 : 8;  // row vectors in flight (ACT carries the y vectors as well), summed in row order
    for (; r + FLY <= r1; r += FLY, p += FLY * rp) {
        Pack a[FLY];
#pragma unroll
        for (int i = 0; i < FLY; i++) a[i] = load(p + i * rp);
#pragma unroll
        for (int i = 0; i < FLY; i++) finish(a[i], r + i);
#pragma unroll
        for (int i = 0; i < FLY; i++)
#pragma unroll
            for (int e = 0; e < VEC; e++) acc[e] += a[i].e[e];
    }
    for (; r < r1; r++, p += rp) {
        Pack a = load(p);
        finish(a, r);
#pragma unroll
        for (int e = 0; e < VEC; e++) acc[e] += a.e[e];
    }
    T *o = out + (uint64_t)blockIdx.y * out_pitch + c;
#pragma unroll
    for (int e = 0; e < VEC; e++)
        if (c + e < n_cols) o[e] = acc[e];
}

// Stage 2: out[c] = sum over chunks.  32 columns x 8 chunk-lanes per block: lane j adds chunks j, j+8, ... (4 loads in
// flight), the 8 lane sums are folded in lane order through shared memory -- fixed association, so the result is the
// same on every launch.  (One thread per column walking all chunks serially cost ~15 us for 296 chunks.)
template <typename T>
__global__ void __launch_bounds__(256) colsum_stage2(const T *__restrict__ part, uint64_t n_chunks, uint64_t pitch,
                                                     uint64_t n_cols, T *__restrict__ out) {
    __shared__ T lane_sum[8][33];
    const uint64_t c = (uint64_t)blockIdx.x * 32 + threadIdx.x;
    const int j = threadIdx.y;
    T acc = (T)0;
    if (c < n_cols) {
        uint64_t k = j;
        for (; k + 24 < n_chunks; k += 32) {
            const T a0 = part[k * pitch + c], a1 = part[(k + 8) * pitch + c], a2 = part[(k + 16) * pitch + c],
                    a3 = part[(k + 24) * pitch + c];
            acc += a0; acc += a1; acc += a2; acc += a3;
        }
        for (; k < n_chunks; k += 8) acc += part[k * pitch + c];
    }
    lane_sum[j][threadIdx.x] = acc;
    __syncthreads();
    if (j == 0 && c < n_cols) {
        T r = lane_sum[0][threadIdx.x];
#pragma unroll
        for (int i = 1; i < 8; i++) r += lane_sum[i][threadIdx.x];
        out[c] = r;
    }
}

}  // namespace wk

using namespace wk;

WK_API int32_t wk_sum(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t row_pitch, uint64_t slice_pitch,
                      const void *x, void *host_out) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        return reduce_runs<scalar_t, 1>(q, (const scalar_t *)x, nullptr, depth, rows, row_pitch, row_pitch, slice_pitch, 0, 0,
                                        host_out);
    });
}

WK_API int32_t wk_dot_reduce(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *x,
                             uint64_t xrp, uint64_t xsp, const void *y, uint64_t yrp, uint64_t ysp, void *host_out) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        return reduce_runs<scalar_t, 2>(q, (const scalar_t *)x, (const scalar_t *)y, depth, rows, cols, xrp, xsp, yrp, ysp,
                                        host_out);
    });
}

WK_API int32_t wk_sum_async(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t row_pitch, uint64_t slice_pitch,
                            const void *x, void *device_out) {
    WK_CHECK_QUEUE(q);
    if (!device_out) return WK_ERR_INVALID_BUFFER;
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        return reduce_runs<scalar_t, 1>(q, (const scalar_t *)x, nullptr, depth, rows, row_pitch, row_pitch, slice_pitch, 0, 0,
                                        nullptr, device_out);
    });
}

WK_API int32_t wk_dot_reduce_async(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *x,
                                   uint64_t xrp, uint64_t xsp, const void *y, uint64_t yrp, uint64_t ysp, void *device_out) {
    WK_CHECK_QUEUE(q);
    if (!device_out) return WK_ERR_INVALID_BUFFER;
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        return reduce_runs<scalar_t, 2>(q, (const scalar_t *)x, (const scalar_t *)y, depth, rows, cols, xrp, xsp, yrp, ysp,
                                        nullptr, device_out);
    });
}

namespace wk {
// column sums of sens (optionally of sens * act'(y), written back over sens): the launcher behind wk_bias_step and the fused
// first step of wk_linear_backward
template <typename scalar_t, int ACT>
static int32_t colsum_launch(wk_queue *q, scalar_t *sens, scalar_t *bias_grad, uint64_t row_pitch, uint64_t rows, uint64_t n_cols,
                             const scalar_t *y, uint64_t y_rp) {
    constexpr uint64_t VEC = 16 / sizeof(scalar_t);
    // the vector kernel reads whole 128-bit column groups: the last group may reach into the row's padding columns
    const bool vec = aligned16(sens) && row_pitch % VEC == 0 && (n_cols + VEC - 1) / VEC * VEC <= row_pitch &&
                     (!ACT || (aligned16(y) && y_rp % VEC == 0 && (n_cols + VEC - 1) / VEC * VEC <= y_rp));
    const uint64_t col_units = vec ? (n_cols + VEC - 1) / VEC : n_cols;
    const uint64_t gx = (col_units + 255) / 256;
    // enough chunks to fill the machine, but never split short columns (keeps the reference's order)
    uint64_t chunks = 1;
    if (rows > 256 || (ACT && rows >= 32)) {  // (the fused pass is a new op: no reference summation order to keep for short columns)
        // exactly one resident wave: (blocks that fit on an SM) x SMs, so no partial second wave trails behind
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vec ? colsum_stage1<scalar_t, true, ACT> : colsum_stage1<scalar_t, false, ACT>,
                                                          256, 0) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 4;
        }
        chunks = ((uint64_t)q->sm_count * per_sm) / gx;
        const uint64_t max_chunks = ACT ? (rows + 7) / 8 : (rows + 63) / 64;
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks > 65535) chunks = 65535;
        if (chunks < 1) chunks = 1;
    }
    const uint64_t rows_per = (rows + chunks - 1) / chunks;
    auto stage1 = [&](scalar_t *out, uint64_t out_pitch) {
        const dim3 grid((unsigned)gx, (unsigned)chunks);
        if (vec)
            colsum_stage1<scalar_t, true, ACT><<<grid, 256, 0, q->stream>>>(sens, row_pitch, rows, n_cols, rows_per, out, out_pitch, y, y_rp, sens);
        else
            colsum_stage1<scalar_t, false, ACT><<<grid, 256, 0, q->stream>>>(sens, row_pitch, rows, n_cols, rows_per, out, out_pitch, y, y_rp, sens);
    };
    if (chunks == 1) {
        stage1(bias_grad, 0);
        WK_CHECK_LAUNCH();
        return WK_OK;
    }
    int32_t rc = ensure_scratch(q, chunks * n_cols * sizeof(scalar_t));
    if (rc != WK_OK) return rc;
    stage1((scalar_t *)q->scratch, n_cols);
    WK_CHECK_LAUNCH();
    colsum_stage2<scalar_t><<<(unsigned)((n_cols + 31) / 32), dim3(32, 8), 0, q->stream>>>((const scalar_t *)q->scratch, chunks, n_cols, n_cols,
                                                                                         bias_grad);
    WK_CHECK_LAUNCH();
    return WK_OK;
}

// sens *= act'(output) and bias_grad = column sums of the result, one pass over [rows, n_cols] (f32 / f64)
int32_t act_backward_colsum(wk_queue *q, int32_t dtype, int32_t act, const void *output, uint64_t out_pitch, void *sens,
                            uint64_t row_pitch, uint64_t rows, uint64_t n_cols, void *bias_grad) {
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        if (act == WK_ACT_SIGMOID)
            return colsum_launch<scalar_t, WK_ACT_SIGMOID>(q, (scalar_t *)sens, (scalar_t *)bias_grad, row_pitch, rows, n_cols,
                                                           (const scalar_t *)output, out_pitch);
        return colsum_launch<scalar_t, WK_ACT_TANH>(q, (scalar_t *)sens, (scalar_t *)bias_grad, row_pitch, rows, n_cols,
                                                    (const scalar_t *)output, out_pitch);
    });
}
}  // namespace wk

WK_API int32_t wk_bias_step(wk_queue *q, int32_t dtype, const void *sens, void *bias_grad, uint64_t row_pitch, uint64_t rows,
                            uint64_t n_cols) {
    WK_CHECK_QUEUE(q);
    if (!sens || !bias_grad) return WK_ERR_INVALID_BUFFER;
    if (n_cols == 0) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_REAL(dtype, [&]() -> int32_t {
        return wk::colsum_launch<scalar_t, 0>(q, (scalar_t *)const_cast<void *>(sens), (scalar_t *)bias_grad, row_pitch, rows, n_cols, nullptr, 0);
    });
}
