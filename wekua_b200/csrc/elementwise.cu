// elementwise.cu -- the HBM-bound streaming kernels: axpy / scal / hadamard / unary math / activations and
// their derivatives / bias / mse / optimizers / fill / uniform / identity / transpose.
//
// One generic kernel template does the streaming: every thread moves 128-bit vectors, UNROLL independent
// vectors per pointer are in flight before the first use (memory-level parallelism, guideline 7/13), the grid
// is a multiple of the SM count and grid-strides over the buffer.  Read-only operands use the non-coherent
// path with L1 no-allocate; in/out operands use L1 no-allocate loads and plain stores.
//
// Reference kernels replaced: src/blas/kernels/axpy.cl, src/math/kernels/{dot,trig}.cl,
// src/nn/activation/kernels/{sigmoid,tanh}.cl, src/nn/layer/kernels/bias.cl, src/nn/loss/kernels/mse.cl,
// src/nn/optimizers/kernels/{gdm,adagrad,rmsprop}.cl, src/tensor/kernels/{fill,identity,transpose}.cl,
// src/tensor/random/kernels/uniform.cl.
#include <float.h>
#include <math.h>

#include <limits>

#include "common.cuh"

namespace wk {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kCtasPerSm = 8;
static int env_knob(const char *name, int def) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : def;
}

template <int NP> struct Ptrs { void *p[NP]; };

__device__ __forceinline__ uint4 ld_ro(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_rw(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_na(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// F: struct with  static constexpr unsigned kRead, kWrite (bit i = pointer i);
//                 __device__ void operator()(T (&v)[NP], uint64_t idx) const   (idx = element index)
// vectors in flight per thread and pointer: kernels that read ONE operand from DRAM (scal, unary maps, activation
// derivatives, fill-like) need twice the depth of the two- and three-operand kernels for the same bytes in flight
// Grid size of the streaming maps.  One chunk (kThreads x MapUnroll vectors per pointer: 16 KiB) per CTA and as many CTAs as
// the tensor has chunks beats a persistent
// grid of 8 CTAs per SM looping over the tensor by 4-8 % on every light map (profiles/sweep_stream_grid_r02.md: f32 axpy
// 0.97 -> 1.06 of the copy bandwidth, f64 1.08): the block scheduler evens out the SMs' progress, a persistent grid ends
// with stragglers.  kCtasPerSm > 0 would cap the grid at that many CTAs per SM for a functor (none needs it: the f64 adagrad /
// rmsprop maps that lost without a cap at 4 vectors per thread gain at 2: 0.94 / 0.97 capped, 1.07 / 1.06 uncapped).
// WK_MAP_CTAS_PER_SM overrides.
template <typename F> struct MapGridCap { static constexpr int kCtasPerSm = 0; };
// the multi-tensor kernel keeps 4 vectors per thread for every functor, so there the two f64 maps keep their cap (below)
template <typename F> struct MtGridCap { static constexpr int kCtasPerSm = 0; };
// Vectors in flight per thread and pointer (a CTA's chunk is kThreads x this many 16-byte vectors).  With one chunk per CTA
// the memory system is kept busy by the NUMBER of CTAs, so depth beyond 4 only costs registers (round 1 doubled it for
// one-pointer maps under its persistent grid): 4 instead of 8 lifts the unary maps by up to 12 % (f32 sin 0.94 -> 1.02,
// tan 0.92 -> 1.01; f64 cosh 0.94 -> 1.05, sigmoid 0.92 -> 1.03).  The register-hungriest multi-pointer maps take 2
// (specialisations next to the functors; profiles/sweep_stream_grid_r02.md).
template <typename F> struct MapUnroll { static constexpr int value = kUnroll; };
template <typename F> constexpr int unroll_for() { return MapUnroll<F>::value; }

template <typename T, int NP, typename F>
__global__ void __launch_bounds__(kThreads) map_vec_kernel(Ptrs<NP> ptrs, uint64_t n, F f) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int kUnroll = unroll_for<F>();  // shadows the file-scope default
    union Pack { uint4 u; T e[VEC]; };
    const uint64_t n_vec = n / VEC;
    const uint64_t chunk = (uint64_t)kThreads * kUnroll;
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < n_vec; base += (uint64_t)gridDim.x * chunk) {
        Pack reg[NP][kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const uint64_t vi = base + (uint64_t)u * kThreads + threadIdx.x;
            if (vi < n_vec) {
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    if (F::kRead & (1u << p)) {
                        const uint4 *src = reinterpret_cast<const uint4 *>(ptrs.p[p]) + vi;
                        reg[p][u].u = (F::kWrite & (1u << p)) ? ld_rw(src) : ld_ro(src);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const uint64_t vi = base + (uint64_t)u * kThreads + threadIdx.x;
            if (vi < n_vec) {
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    T v[NP];
#pragma unroll
                    for (int p = 0; p < NP; p++) v[p] = reg[p][u].e[e];
                    f(v, vi * VEC + e);
#pragma unroll
                    for (int p = 0; p < NP; p++) reg[p][u].e[e] = v[p];
                }
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (F::kWrite & (1u << p)) st_na(reinterpret_cast<uint4 *>(ptrs.p[p]) + vi, reg[p][u].u);
            }
        }
    }
    // tail: fewer than VEC elements
    if (blockIdx.x == 0) {
        const uint64_t i = n_vec * VEC + threadIdx.x;
        if (i < n) {
            T v[NP];
#pragma unroll
            for (int p = 0; p < NP; p++)
                if (F::kRead & (1u << p)) v[p] = reinterpret_cast<const T *>(ptrs.p[p])[i];
            f(v, i);
#pragma unroll
            for (int p = 0; p < NP; p++)
                if (F::kWrite & (1u << p)) reinterpret_cast<T *>(ptrs.p[p])[i] = v[p];
        }
    }
}

// scalar fallback (unaligned pointers)
template <typename T, int NP, typename F>
__global__ void __launch_bounds__(kThreads) map_scalar_kernel(Ptrs<NP> ptrs, uint64_t n, F f) {
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
        T v[NP];
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (F::kRead & (1u << p)) v[p] = reinterpret_cast<const T *>(ptrs.p[p])[i];
        f(v, i);
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (F::kWrite & (1u << p)) reinterpret_cast<T *>(ptrs.p[p])[i] = v[p];
    }
}

// pitched [depth, rows, cols] variant (logical region of padded tensors): pointer p uses pitches rp[p]/sp[p]
template <int NP> struct Pitches { uint64_t rp[NP], sp[NP]; };
template <typename T, int NP, typename F>
__global__ void __launch_bounds__(kThreads) map_pitched_kernel(Ptrs<NP> ptrs, Pitches<NP> pit, uint64_t depth,
                                                                uint64_t rows, uint64_t cols, F f) {
    const uint64_t n_rows = depth * rows;
    for (uint64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        const uint64_t d = r / rows, j = r - d * rows;
        for (uint64_t k = (uint64_t)blockIdx.x * kThreads + threadIdx.x; k < cols; k += (uint64_t)gridDim.x * kThreads) {
            T v[NP];
            uint64_t off[NP];
#pragma unroll
            for (int p = 0; p < NP; p++) {
                off[p] = d * pit.sp[p] + j * pit.rp[p] + k;
                if (F::kRead & (1u << p)) v[p] = reinterpret_cast<const T *>(ptrs.p[p])[off[p]];
            }
            f(v, off[0]);
#pragma unroll
            for (int p = 0; p < NP; p++)
                if (F::kWrite & (1u << p)) reinterpret_cast<T *>(ptrs.p[p])[off[p]] = v[p];
        }
    }
}

template <typename T, int NP, typename F>
static int32_t launch_map(wk_queue *q, Ptrs<NP> ptrs, uint64_t n, F f) {
    if (n == 0) return WK_OK;
    bool aligned = true;
    for (int p = 0; p < NP; p++) {
        if (!ptrs.p[p]) {
            set_error("null buffer");
            return WK_ERR_INVALID_BUFFER;
        }
        aligned &= aligned16(ptrs.p[p]);
    }
    static const int cap_env = env_knob("WK_MAP_CTAS_PER_SM", -1);  // (tuning knob: CTAs per SM, 0 = no cap)
    const int ctas_per_sm = cap_env >= 0 ? cap_env : MapGridCap<F>::kCtasPerSm;
    const uint64_t cap = ctas_per_sm > 0 ? (uint64_t)q->sm_count * ctas_per_sm : 0x7fffffffull;
    if (aligned) {
        constexpr int VEC = 16 / (int)sizeof(T);
        constexpr int kUnroll = unroll_for<F>();
        uint64_t blocks = (n / VEC + (uint64_t)kThreads * kUnroll - 1) / ((uint64_t)kThreads * kUnroll);
        if (blocks == 0) blocks = 1;
        if (blocks > cap) blocks = cap;
        map_vec_kernel<T, NP, F><<<(unsigned)blocks, kThreads, 0, q->stream>>>(ptrs, n, f);
    } else {
        uint64_t blocks = (n + kThreads - 1) / kThreads;
        if (blocks > cap) blocks = cap;
        map_scalar_kernel<T, NP, F><<<(unsigned)blocks, kThreads, 0, q->stream>>>(ptrs, n, f);
    }
    WK_CHECK_LAUNCH();
    return WK_OK;
}

// [depth, rows, cols] with pitches: collapses to the dense 1-D kernel when every operand is contiguous
template <typename T, int NP, typename F>
static int32_t launch_map3d(wk_queue *q, Ptrs<NP> ptrs, Pitches<NP> pit, uint64_t depth, uint64_t rows, uint64_t cols,
                            F f) {
    if (depth == 0 || rows == 0 || cols == 0) return WK_ERR_INVALID_VALUE;
    bool dense = true;
    for (int p = 0; p < NP; p++) {
        if (!ptrs.p[p]) {
            set_error("null buffer");
            return WK_ERR_INVALID_BUFFER;
        }
        const bool row_dense = (rows == 1) || (pit.rp[p] == cols);
        const bool slice_dense = (depth == 1) || (pit.sp[p] == rows * cols && row_dense);
        dense &= row_dense && slice_dense;
    }
    if (dense) return launch_map<T, NP, F>(q, ptrs, depth * rows * cols, f);
    uint64_t gx = (cols + kThreads - 1) / kThreads;
    if (gx > 1024) gx = 1024;
    uint64_t gy = depth * rows;
    if (gy > 65535) gy = 65535;
    map_pitched_kernel<T, NP, F><<<dim3((unsigned)gx, (unsigned)gy), kThreads, 0, q->stream>>>(ptrs, pit, depth, rows,
                                                                                              cols, f);
    WK_CHECK_LAUNCH();
    return WK_OK;
}

// ------------------------------------------------------------------------------------------ functors
template <typename T, int MODE>  // MODE 0: y += x, 1: y += alpha*x, 2: y -= x  (axpy.cl:55-65)
struct AxpyF {
    static constexpr unsigned kRead = 3, kWrite = 2;
    typename Acc<T>::type alpha;
    __device__ __forceinline__ void operator()(T (&v)[2], uint64_t) const {
        using A = typename Acc<T>::type;
        const A x = to_acc<T>(v[0]), y = to_acc<T>(v[1]);
        if (MODE == 1) v[1] = from_acc<T>((A)(y + x * alpha));  // complex: COMPLEX_MUL(x, alpha), axpy.cl:41-44
        else if (MODE == 2) v[1] = from_acc<T>((A)(y - x));
        else v[1] = from_acc<T>((A)(y + x));
    }
};

template <typename T> struct ScalF {  // old_kernels/scal.cl:3-20
    static constexpr unsigned kRead = 1, kWrite = 1;
    typename Acc<T>::type alpha;
    __device__ __forceinline__ void operator()(T (&v)[1], uint64_t) const {
        v[0] = from_acc<T>((typename Acc<T>::type)(to_acc<T>(v[0]) * alpha));
    }
};

template <typename T> struct HadamardF {  // dot.cl:33
    static constexpr unsigned kRead = 3, kWrite = 1;
    __device__ __forceinline__ void operator()(T (&v)[2], uint64_t) const {
        v[0] = from_acc<T>((typename Acc<T>::type)(to_acc<T>(v[0]) * to_acc<T>(v[1])));
    }
};

template <typename T> struct FillF {  // fill.cl:11-16
    static constexpr unsigned kRead = 0, kWrite = 1;
    T value;
    __device__ __forceinline__ void operator()(T (&v)[1], uint64_t) const { v[0] = value; }
};

__device__ __forceinline__ float wk_sin(float x) { return wk_sin_f32(x); }
__device__ __forceinline__ double wk_sin(double x) { return wk_sin_f64(x); }
__device__ __forceinline__ float wk_cos(float x) { return wk_cos_f32(x); }
__device__ __forceinline__ double wk_cos(double x) { return wk_cos_f64(x); }
__device__ __forceinline__ float wk_tan(float x) { return wk_tan_f32(x); }
__device__ __forceinline__ double wk_tan(double x) { return wk_tan_f64(x); }
__device__ __forceinline__ float wk_cosh(float x) { return wk_cosh_f32(x); }
__device__ __forceinline__ double wk_cosh(double x) { return wk_cosh_f64(x); }
__device__ __forceinline__ float wk_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double wk_tanh(double x) { return wk_tanh_f64(x); }
__device__ __forceinline__ float wk_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ double wk_sigmoid(double x) { return wk_sigmoid_f64(x); }

template <typename T, int OP> struct UnaryF {  // trig.cl:3-67, sigmoid.cl:14-15
    static constexpr unsigned kRead = 1, kWrite = 1;
    __device__ __forceinline__ void operator()(T (&v)[1], uint64_t) const {
        const T x = v[0];
        if (OP == WK_OP_SIN) v[0] = wk_sin(x);
        else if (OP == WK_OP_COS) v[0] = wk_cos(x);
        else if (OP == WK_OP_TAN) v[0] = wk_tan(x);
        else if (OP == WK_OP_SINH) v[0] = sinh(x);
        else if (OP == WK_OP_COSH) v[0] = wk_cosh(x);
        else if (OP == WK_OP_TANH) v[0] = wk_tanh(x);
        else v[0] = wk_sigmoid(x);
    }
};

// (libdevice's f64 tan / tanh / cosh and the IEEE division of sigmoid were issue-bound here -- 0.70 / 0.65 / 0.86 / 0.84 of the
// copy bandwidth -- until they got the branch-free constant-bank implementations of common.cuh.)

template <typename BT, int OP> struct UnaryF<Cx<BT>, OP> {  // trig.cl:5-66, complex branches
    static constexpr unsigned kRead = 1, kWrite = 1;
    __device__ __forceinline__ void operator()(Cx<BT> (&v)[1], uint64_t) const {
        const BT a = v[0].re, b = v[0].im;
        Cx<BT> r;
        if (OP == WK_OP_SIN) { r.re = sin(a) * cosh(b); r.im = cos(a) * sinh(b); }
        else if (OP == WK_OP_COS) { r.re = cos(a) * cosh(b); r.im = -sin(a) * sinh(b); }
        else if (OP == WK_OP_TAN) { const BT ta = 2 * a, tb = 2 * b, d = cos(ta) + cosh(tb); r.re = sin(ta) / d; r.im = sinh(tb) / d; }
        else if (OP == WK_OP_SINH) { r.re = sinh(a) * cos(b); r.im = cosh(a) * sin(b); }
        else if (OP == WK_OP_COSH) { r.re = cosh(a) * cos(b); r.im = sinh(a) * sin(b); }
        else { const BT ta = 2 * a, tb = 2 * b, d = cosh(ta) + cos(tb); r.re = sinh(ta) / d; r.im = sin(tb) / d; }
        v[0] = r;
    }
};

template <typename T, int ACT> struct ActDevF {  // sigmoid.cl:31-32 (ACT=1), tanh.cl:16-17 (ACT=2)
    static constexpr unsigned kRead = 1, kWrite = 2;
    __device__ __forceinline__ void operator()(T (&v)[2], uint64_t) const {
        const T y = v[0];
        v[1] = ACT == WK_ACT_SIGMOID ? y * ((T)1 - y) : (T)1 - y * y;
    }
};

// fused: derivative (optionally stored) and sensitivity *= derivative  (linear.zig:608-613)
template <typename T, int ACT, bool STORE> struct ActBackwardF {
    static constexpr unsigned kRead = 1 | 4, kWrite = (STORE ? 2 : 0) | 4;
    __device__ __forceinline__ void operator()(T (&v)[3], uint64_t) const {
        const T y = v[0];
        const T d = ACT == WK_ACT_SIGMOID ? y * ((T)1 - y) : (ACT == WK_ACT_TANH ? (T)1 - y * y : (T)1);
        if (STORE) v[1] = d;
        v[2] = v[2] * d;
    }
};

template <typename T, bool DEV> struct MseF {  // mse.cl:28-33: v = {output, expected, error, dev}
    static constexpr unsigned kRead = 3, kWrite = 4 | (DEV ? 8 : 0);
    __device__ __forceinline__ void operator()(T (&v)[4], uint64_t) const {
        const T err = v[1] - v[0];
        v[2] = err * err;
        if (DEV) v[3] = -(T)2 * err;
    }
};

template <> struct MapUnroll<MseF<double, false>> { static constexpr int value = 2; };  // 1.02 -> 1.07

template <typename T> struct GdmF {  // gdm.cl:28-31: v = {x, g, velocity}
    static constexpr unsigned kRead = 7, kWrite = 5;
    T lr, beta;
    __device__ __forceinline__ void operator()(T (&v)[3], uint64_t) const {
        const T nv = beta * v[2] + lr * v[1];
        v[0] = v[0] - nv;
        v[2] = nv;
    }
};

template <typename T> struct AdagradF {  // adagrad.cl:42-46
    static constexpr unsigned kRead = 7, kWrite = 5;
    T lr;
    __device__ __forceinline__ void operator()(T (&v)[3], uint64_t) const {
        const T g = v[1];
        const T h = v[2] + g * g;
        v[0] = v[0] - lr * g / (sqrt(h) + (T)FLT_EPSILON);
        v[2] = h;
    }
};

template <typename T> struct RmspropF {  // rmsprop.cl:51-55
    static constexpr unsigned kRead = 7, kWrite = 5;
    T lr, gamma;
    __device__ __forceinline__ void operator()(T (&v)[3], uint64_t) const {
        const T g = v[1];
        const T h = gamma * v[2] + ((T)1 - gamma) * g * g;
        v[0] = v[0] - lr * g / (sqrt(h) + (T)FLT_EPSILON);
        v[2] = h;
    }
};

template <> struct MapUnroll<GdmF<double>> { static constexpr int value = 2; };      // 0.96 -> 1.07
template <> struct MtGridCap<AdagradF<double>> { static constexpr int kCtasPerSm = 8; };  // (0.87 capped, 0.81 uncapped at 4 vectors)
template <> struct MtGridCap<RmspropF<double>> { static constexpr int kCtasPerSm = 8; };
template <> struct MapUnroll<AdagradF<double>> { static constexpr int value = 2; };  // 0.89 -> 1.07
template <> struct MapUnroll<RmspropF<double>> { static constexpr int value = 2; };  // 0.89 -> 1.06

template <typename T> struct AdamF {  // textbook Adam (Kingma & Ba alg. 1); v = {x, g, m, v}
    static constexpr unsigned kRead = 15, kWrite = 1 | 4 | 8;
    T lr, b1, b2, eps, c1, c2;  // c1 = 1/(1-b1^t), c2 = 1/(1-b2^t)
    __device__ __forceinline__ void operator()(T (&v)[4], uint64_t) const {
        const T g = v[1];
        const T m = b1 * v[2] + ((T)1 - b1) * g;
        const T s = b2 * v[3] + ((T)1 - b2) * g * g;
        v[0] = v[0] - lr * (m * c1) / (sqrt(s * c2) + eps);
        v[2] = m;
        v[3] = s;
    }
};

template <> struct MapUnroll<AdamF<float>> { static constexpr int value = 2; };  // 0.96 -> 1.05 (f64: 0.99 at 4, 0.94 at 2)

// uniform.cl:32-54 (little-endian branch; `seed2 << 32` is always 0; `*` binds tighter than `^`), spelled out on 32-bit
// halves.  The hash is the whole cost of the kernel (an f32 element is 4 bytes of HBM against ~25 integer instructions, and
// the shift / logic pipe and the multiply pipe each retire 16 lanes per SM sub-partition and clock), so every instruction
// counts and the two pipes are balanced:
//   * combined = (idx_lo << 32) + idx_hi puts the HIGH index word (0 for any tensor below 2^32 elements) into the LOW half
//     of x0, so every term that only depends on x0_lo is an invariant (XxhInv): the x0_lo parts of both rotations, and the
//     x0_lo part of rotl(x0, 24) * C (rotations are linear over disjoint bit fields: rotl(x0,24) = var + const);
//   * x * C mod 2^64 is one widening multiply-add plus two 32-bit multiply-adds into the high word (the `<< 24` of the
//     variable part folds into the constant C_LO << 24);
//   * (x1 >> 35) + 8 fits 32 bits: its product with C is one widening multiply-add (the "+ 8" rides along as the 64-bit
//     addend 8 * C) plus one multiply-add;
//   * right shifts of a 32-bit word are written as mul.hi by a power of two (multiply pipe) where that balances the pipes.
// Same arithmetic mod 2^64, bit for bit (tests/test_gpu_parity.py pins it on the oracle's 64-bit restatement).
struct XxhInv { uint32_t key_hi, kk, k_lo, mc_lo, mc_hi; };
__device__ __forceinline__ uint64_t wk_xxhash64_key(uint64_t seed) { return (0x7C01812CF721AD1CULL ^ 0xDED46DE9839097DBULL) - seed; }
__device__ __forceinline__ XxhInv wk_xxhash64_inv(uint32_t idx_hi, uint32_t key_lo, uint32_t key_hi) {
    const uint32_t x0_lo = idx_hi ^ key_lo;
    const uint64_t mc = (((uint64_t)(x0_lo >> 8) << 32) | (uint32_t)(x0_lo << 24)) * 0x9FB21C651E98DF25ULL;  // const part of rotl(x0,24) * C
    return XxhInv{key_hi, key_hi ^ (x0_lo << 17), x0_lo ^ (x0_lo >> 15), (uint32_t)mc, (uint32_t)(mc >> 32)};
}
__device__ __forceinline__ void wk_xxhash64_elem(uint32_t idx_lo, const XxhInv &k, uint32_t &h_lo, uint32_t &h_hi) {
    constexpr uint32_t C_LO = 0x1E98DF25u, C_HI = 0x9FB21C65u, C_LO24 = C_LO << 24;
    constexpr uint64_t C8 = 8ULL * 0x9FB21C651E98DF25ULL;
    const uint32_t x0_hi = idx_lo ^ k.key_hi;  // x0 = combined ^ key; x0_lo lives in the invariants
    const uint32_t t = idx_lo ^ k.kk;          // x0_hi ^ (x0_lo << 17)
    const uint32_t v = __umulhi(x0_hi, 1u << 24);  // x0_hi >> 8: low word of the variable part of rotl(x0, 24)
    const uint64_t m = (uint64_t)v * C_LO + (((uint64_t)k.mc_hi << 32) | k.mc_lo);
    const uint32_t m_hi = (uint32_t)(m >> 32) + v * C_HI + x0_hi * C_LO24;
    const uint32_t x1_lo = k.k_lo ^ (x0_hi << 17) ^ (uint32_t)m;          // x0 ^ rotl(x0, 49) ^ rotl(x0, 24) * C
    const uint32_t x1_hi = t ^ __umulhi(x0_hi, 1u << 17) ^ m_hi;          // (x0_hi >> 15)
    const uint32_t s = x1_hi >> 3;                                        // x1 >> 35
    const uint64_t p = (uint64_t)s * C_LO + C8;
    const uint32_t p_hi = (uint32_t)(p >> 32) + s * C_HI;
    const uint32_t x2_lo = x1_lo ^ (uint32_t)p, x2_hi = x1_hi ^ p_hi;
    h_lo = x2_lo ^ __funnelshift_r(x2_lo, x2_hi, 28);  // x2 ^ (x2 >> 28)
    h_hi = x2_hi ^ (x2_hi >> 28);
}
__device__ __forceinline__ void wk_xxhash64_halves(uint32_t idx_lo, uint32_t idx_hi, uint32_t key_lo, uint32_t key_hi, uint32_t &h_lo,
                                                   uint32_t &h_hi) {
    wk_xxhash64_elem(idx_lo, wk_xxhash64_inv(idx_hi, key_lo, key_hi), h_lo, h_hi);
}
__device__ __forceinline__ uint64_t wk_xxhash64(uint64_t index, uint64_t seed) {
    const uint64_t key = wk_xxhash64_key(seed);
    uint32_t lo, hi;
    wk_xxhash64_halves((uint32_t)index, (uint32_t)(index >> 32), (uint32_t)key, (uint32_t)(key >> 32), lo, hi);
    return ((uint64_t)hi << 32) | lo;
}

// (T)h for the 64-bit hash h = hi:lo.  Measured dead ends (profiles/uniform_r02.md): replacing the 64-bit I2F by exact 32-bit
// forms (double: two conversions + one FMA; float: hi | sticky through I2FP with a rare 64-bit fallback) costs more integer
// instructions than the conversion unit saves -- the kernel is bound by the multiply (FMA-heavy) and logic pipes, not by XU.
__device__ __forceinline__ double wk_u64_to_double(uint32_t lo, uint32_t hi) { return (double)(((uint64_t)hi << 32) | lo); }
__device__ __forceinline__ float wk_u64_to_float(uint32_t lo, uint32_t hi) { return (float)(((uint64_t)hi << 32) | lo); }
template <typename T, bool RANGE> __device__ __forceinline__ T uniform_value(uint32_t lo, uint32_t hi, T min_value, T range) {
    if (RANGE) {
        // cl_khr_fp64 branch: (double)h / ULONG_MAX; ULONG_MAX converts to 2^64 (a power of two: the division is an exact scaling)
        const double normalized = wk_u64_to_double(lo, hi) * 0x1p-64;
        return (T)((double)min_value + normalized * (double)range);
    } else if (std::is_same<T, float>::value) {
        return (T)(wk_u64_to_float(lo, hi) * 0x1p-64f);  // (float)h / 2^64f
    } else if (std::is_same<T, double>::value) {
        return (T)(wk_u64_to_double(lo, hi) * 0x1p-64);
    }
    return (T)(((uint64_t)hi << 32) | lo);  // & WK_UINT_MAX == truncation
}
template <typename T, bool RANGE> __device__ __forceinline__ T uniform_value(uint64_t h, T min_value, T range) {
    return uniform_value<T, RANGE>((uint32_t)h, (uint32_t)(h >> 32), min_value, range);
}

template <typename T, bool RANGE> struct UniformF {  // uniform.cl:56-185, idx = padded linear index
    static constexpr unsigned kRead = 0, kWrite = 1;
    uint64_t seed;
    T min_value, range;
    __device__ __forceinline__ void operator()(T (&v)[1], uint64_t idx) const {
        v[0] = uniform_value<T, RANGE>(wk_xxhash64(idx, seed), min_value, range);
    }
};
// complex branch, uniform.cl:80-93: the components hash (index << 1) and (index << 1) + 1; bounds are base-type scalars
template <typename BT, bool RANGE> struct UniformF<Cx<BT>, RANGE> {
    static constexpr unsigned kRead = 0, kWrite = 1;
    uint64_t seed;
    BT min_value, range;
    __device__ __forceinline__ void operator()(Cx<BT> (&v)[1], uint64_t idx) const {
        v[0].re = uniform_value<BT, RANGE>(wk_xxhash64(idx << 1, seed), min_value, range);
        v[0].im = uniform_value<BT, RANGE>(wk_xxhash64((idx << 1) + 1, seed), min_value, range);
    }
};

// Dense tensors (padded index == dense index) of a real base type: one thread per 16-byte vector, the index halves and the key
// stay in 32-bit registers, nothing is read.  A complex tensor is the same stream over its base type: its components hash
// (index << 1) and (index << 1) + 1 (uniform.cl:80-93), i.e. the flat index of the component.
template <typename T, bool RANGE, bool SMALL>  // SMALL: n <= 2^32, the high index word is 0 and the invariants are per thread
__global__ void __launch_bounds__(kThreads) uniform_dense_kernel(T *__restrict__ out, uint64_t n, uint32_t key_lo, uint32_t key_hi,
                                                                 T min_value, T range) {
    constexpr int VEC = 16 / (int)sizeof(T);
    union Pack { uint4 u; T e[VEC]; };
    const uint64_t n_vec = n / VEC;
    XxhInv inv = wk_xxhash64_inv(0, key_lo, key_hi);
    for (uint64_t vi = (uint64_t)blockIdx.x * kThreads + threadIdx.x; vi < n_vec; vi += (uint64_t)gridDim.x * kThreads) {
        const uint64_t i0 = vi * VEC;
        const uint32_t i_lo = (uint32_t)i0;
        if (!SMALL) inv = wk_xxhash64_inv((uint32_t)(i0 >> 32), key_lo, key_hi);
        Pack pk;
#pragma unroll
        for (int e = 0; e < VEC; e++) {
            uint32_t lo, hi;
            wk_xxhash64_elem(i_lo | (uint32_t)e, inv, lo, hi);  // (i0 is a multiple of VEC)
            pk.e[e] = uniform_value<T, RANGE>(lo, hi, min_value, range);
        }
        st_na(reinterpret_cast<uint4 *>(out) + vi, pk.u);
    }
    if (blockIdx.x == 0 && n_vec * VEC + threadIdx.x < n) {  // tail: fewer than VEC elements
        const uint64_t i = n_vec * VEC + threadIdx.x;
        uint32_t lo, hi;
        wk_xxhash64_halves((uint32_t)i, (uint32_t)(i >> 32), key_lo, key_hi, lo, hi);
        out[i] = uniform_value<T, RANGE>(lo, hi, min_value, range);
    }
}
template <typename T, bool RANGE>
static int32_t launch_uniform_dense(wk_queue *q, T *out, uint64_t n, uint64_t seed, T min_value, T range) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const uint64_t key = (0x7C01812CF721AD1CULL ^ 0xDED46DE9839097DBULL) - seed;
    uint64_t blocks = (n / VEC + kThreads - 1) / kThreads;
    // (persistent grid on purpose: the per-thread hash invariants are set up once per thread; uncapped measured 0.41 / 0.43)
    const uint64_t cap = (uint64_t)q->sm_count * kCtasPerSm;
    blocks = blocks == 0 ? 1 : (blocks > cap ? cap : blocks);
    if (n <= (1ull << 32))
        uniform_dense_kernel<T, RANGE, true><<<(unsigned)blocks, kThreads, 0, q->stream>>>(out, n, (uint32_t)key, (uint32_t)(key >> 32), min_value, range);
    else
        uniform_dense_kernel<T, RANGE, false><<<(unsigned)blocks, kThreads, 0, q->stream>>>(out, n, (uint32_t)key, (uint32_t)(key >> 32), min_value, range);
    WK_CHECK_LAUNCH();
    return WK_OK;
}

// bias.cl:3-19: out[i] += bias[i % row_pitch], i < n.  Flat streaming over the buffer like map_vec_kernel (128-bit
// accesses, kBiasUnroll vectors of `out` in flight per thread -- only `out` comes from DRAM, so it needs a deeper
// unroll than the two-operand kernels to keep the same bytes in flight); each thread tracks the bias column of its
// vectors incrementally -- one modulo before the loop, an add and a conditional subtract per step -- and reads bias
// through L1 at the point of use.  VECTOR = false is the same loop over single elements (row pitch or n not a
// multiple of the vector width).  rp_units < 2^32 (host-checked).
constexpr int kBiasUnroll = 4;  // (8 under round 1's capped grid; with one chunk per CTA 4 is faster: f64 1.00 -> 1.06)
template <typename T> __device__ __forceinline__ T bias_act_apply(T v, int) { return v; }
template <> __device__ __forceinline__ float bias_act_apply<float>(float v, int act) {
    return act == WK_ACT_SIGMOID ? wk_sigmoid(v) : act == WK_ACT_TANH ? wk_tanh(v) : v;
}
template <> __device__ __forceinline__ double bias_act_apply<double>(double v, int act) {
    return act == WK_ACT_SIGMOID ? wk_sigmoid(v) : act == WK_ACT_TANH ? wk_tanh(v) : v;
}

// ACT: out = act(out + bias) in the same pass (Linear.forward's addBias + Activation.run, linear.zig:499-521, for layers whose
// GEMM epilogue is not hidden behind another tile: see wk_gemm_bias_act)
template <typename T, bool VECTOR, bool ACT = false>
__global__ void __launch_bounds__(kThreads, 4) bias_add_kernel(T *__restrict__ out, const T *__restrict__ bias, uint32_t rp_units,
                                                               uint64_t n_units, int act = 0) {
    constexpr int VEC = VECTOR ? 16 / (int)sizeof(T) : 1;
    union Pack { uint4 u; T e[16 / sizeof(T)]; };
    const uint64_t chunk = (uint64_t)kThreads * kBiasUnroll;
    const uint64_t step = (uint64_t)gridDim.x * chunk;
    const uint32_t step_mod = (uint32_t)(step % rp_units), thr_mod = (uint32_t)(kThreads % rp_units);
    // column of this thread's first vector; the u-th vector sits u * kThreads further on
    uint32_t col0 = (uint32_t)(((uint64_t)blockIdx.x * chunk + threadIdx.x) % rp_units);
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < n_units; base += step) {
        Pack o[kBiasUnroll];
#pragma unroll
        for (int u = 0; u < kBiasUnroll; u++) {
            const uint64_t vi = base + (uint64_t)u * kThreads + threadIdx.x;
            if (vi < n_units) {
                if (VECTOR) o[u].u = ld_rw(reinterpret_cast<const uint4 *>(out) + vi);
                else o[u].e[0] = out[vi];
            }
        }
        uint32_t col = col0;
#pragma unroll
        for (int u = 0; u < kBiasUnroll; u++) {
            const uint64_t vi = base + (uint64_t)u * kThreads + threadIdx.x;
            if (vi < n_units) {
                Pack b;
                if (VECTOR) b.u = __ldg(reinterpret_cast<const uint4 *>(bias) + col);
                else b.e[0] = __ldg(bias + col);
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    o[u].e[e] += b.e[e];
                    if (ACT) o[u].e[e] = bias_act_apply<T>(o[u].e[e], act);
                }
                if (VECTOR) st_na(reinterpret_cast<uint4 *>(out) + vi, o[u].u);
                else out[vi] = o[u].e[0];
            }
            col += thr_mod;
            if (col >= rp_units) col -= rp_units;
        }
        col0 += step_mod;
        if (col0 >= rp_units) col0 -= rp_units;
    }
}

template <typename T>
__global__ void identity_kernel(T *__restrict__ buf, uint64_t size, uint64_t pitch_sum) {  // identity.cl:3-20
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < size) buf[i * pitch_sum] = Conv<T>::one();
}

template <typename T>
__global__ void __launch_bounds__(256) transpose2d_kernel(const T *__restrict__ src, uint64_t sp, T *__restrict__ dst,
                                                          uint64_t dp, uint64_t rows, uint64_t cols) {
    __shared__ T tile[32][33];
    const uint64_t c0 = (uint64_t)blockIdx.x * 32, r0 = (uint64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const uint64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[r * sp + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const uint64_t c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[c * dp + r] = tile[threadIdx.x][j];
    }
}

// 64 x 64 (64 x 32 for 8-byte elements) tile, 128-bit global accesses on both sides, elements cross over through
// a padded shared tile with scalar accesses (2-way bank conflicts at worst, far below the HBM time of the tile).
// Needs rows, cols and both pitches to be multiples of the vector width and 16-byte aligned pointers.
template <typename T>
__global__ void __launch_bounds__(256) transpose2d_vec_kernel(const T *__restrict__ src, uint64_t sp, T *__restrict__ dst,
                                                              uint64_t dp, uint64_t rows, uint64_t cols) {
    // 64 rows x TC columns; 8-byte elements take 32 columns so that a block keeps the 16 KiB footprint (and the 8+
    // resident blocks per SM) of the 4-byte case -- with 32 KiB tiles only 4 blocks fit and their load / store phases
    // no longer overlap enough (ncu: 61 % of DRAM peak)
    constexpr int VEC = 16 / (int)sizeof(T), TS = 64, TC = sizeof(T) == 8 ? 32 : 64, VPR = TC / VEC, ITER = TS * VPR / 256;
    constexpr int VPC = TS / VEC, ITER_OUT = TC * VPC / 256;
    __shared__ T tile[TS][TC + 1];
    union Pack { uint4 u; T e[VEC]; };
    const uint64_t c0 = (uint64_t)blockIdx.x * TC, r0 = (uint64_t)blockIdx.y * TS;
    Pack in[ITER];
#pragma unroll
    for (int k = 0; k < ITER; k++) {
        const int v = k * 256 + threadIdx.x, r = v / VPR, cv = v % VPR;
        const uint64_t gr = r0 + r, gc = c0 + (uint64_t)cv * VEC;
        if (gr < rows && gc < cols) in[k].u = ld_ro(reinterpret_cast<const uint4 *>(src + gr * sp + gc));
    }
#pragma unroll
    for (int k = 0; k < ITER; k++) {
        const int v = k * 256 + threadIdx.x, r = v / VPR, cv = v % VPR;
#pragma unroll
        for (int e = 0; e < VEC; e++) tile[r][cv * VEC + e] = in[k].e[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITER_OUT; k++) {
        const int v = k * 256 + threadIdx.x, c = v / VPC, rv = v % VPC;
        const uint64_t gc = c0 + c, gr = r0 + (uint64_t)rv * VEC;
        if (gc < cols && gr < rows) {
            Pack o;
#pragma unroll
            for (int e = 0; e < VEC; e++) o.e[e] = tile[rv * VEC + e][c];
            st_na(reinterpret_cast<uint4 *>(dst + gc * dp + gr), o.u);
        }
    }
}

// N-D swap of two dimensions through the pitch arrays (transpose.cl:3-42): one thread per PADDED source element,
// padding skipped, the destination index rebuilt dimension by dimension with dim0/dim1's pitches exchanged.
struct TransposeNd {
    uint64_t pa[8], pb[8];
    uint64_t row_pitch, slice_pitch, height, cols, n;
    uint32_t ndim, dim0, dim1;
};
template <typename T>
__global__ void __launch_bounds__(256) transpose_nd_kernel(const T *__restrict__ a, T *__restrict__ b, const TransposeNd p) {
    for (uint64_t i0 = (uint64_t)blockIdx.x * 256 + threadIdx.x; i0 < p.n; i0 += (uint64_t)gridDim.x * 256) {
        if ((i0 % p.row_pitch) >= p.cols || (i0 % p.slice_pitch) >= p.height) continue;
        uint64_t index = i0, b_index = 0;
        for (uint32_t x = 0; x < p.ndim; x++) {
            const uint64_t dim_index = index / p.pa[x];
            index -= dim_index * p.pa[x];
            b_index += dim_index * p.pb[x == p.dim0 ? p.dim1 : (x == p.dim1 ? p.dim0 : x)];
        }
        b[b_index + index] = a[i0];
    }
}

// ------------------------------------------------------------------------------- multi-tensor apply
// One launch over a LIST of tensors (Optimizer.step walks every weight / bias tensor of every layer and enqueues one
// kernel each, rmsprop.zig:168-202).  The list travels in the kernel parameters; work is cut into chunks of
// kThreads*kUnroll vectors, chunk ids are global over the list (prefix sums in the table) and blocks grid-stride over
// them, so a list of tiny tensors costs one launch and a list of huge ones streams exactly like map_vec_kernel.
constexpr int kMtMax = 24;
template <int NP> struct MtTable {
    void *p[NP][kMtMax];
    uint64_t n[kMtMax];
    uint64_t first_chunk[kMtMax + 1];
    int count;
};

template <typename T, int NP, typename F>
__global__ void __launch_bounds__(kThreads) mt_map_kernel(const __grid_constant__ MtTable<NP> tab, F f) {
    constexpr int VEC = 16 / (int)sizeof(T);
    union Pack { uint4 u; T e[VEC]; };
    constexpr uint64_t chunk_vecs = (uint64_t)kThreads * kUnroll;
    const uint64_t total = tab.first_chunk[tab.count];
    int t = 0;
    for (uint64_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        while (tab.first_chunk[t + 1] <= chunk) t++;  // chunk ids only grow within a block
        const uint64_t n = tab.n[t], n_vec = n / VEC;
        const uint64_t base = (chunk - tab.first_chunk[t]) * chunk_vecs;
        Pack reg[NP][kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const uint64_t vi = base + (uint64_t)u * kThreads + threadIdx.x;
            if (vi < n_vec) {
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (F::kRead & (1u << p)) {
                        const uint4 *src = reinterpret_cast<const uint4 *>(tab.p[p][t]) + vi;
                        reg[p][u].u = (F::kWrite & (1u << p)) ? ld_rw(src) : ld_ro(src);
                    }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const uint64_t vi = base + (uint64_t)u * kThreads + threadIdx.x;
            if (vi < n_vec) {
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    T v[NP];
#pragma unroll
                    for (int p = 0; p < NP; p++) v[p] = reg[p][u].e[e];
                    f(v, vi * VEC + e);
#pragma unroll
                    for (int p = 0; p < NP; p++) reg[p][u].e[e] = v[p];
                }
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (F::kWrite & (1u << p)) st_na(reinterpret_cast<uint4 *>(tab.p[p][t]) + vi, reg[p][u].u);
            }
        }
        // the tensor's last chunk also owns its tail of fewer than VEC elements
        if (chunk + 1 == tab.first_chunk[t + 1]) {
            const uint64_t i = n_vec * VEC + threadIdx.x;
            if (i < n) {
                T v[NP];
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (F::kRead & (1u << p)) v[p] = reinterpret_cast<const T *>(tab.p[p][t])[i];
                f(v, i);
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (F::kWrite & (1u << p)) reinterpret_cast<T *>(tab.p[p][t])[i] = v[p];
            }
        }
    }
}

// ptr_of(param, slot) picks the buffer of pointer slot `slot` (functor order); lists longer than kMtMax take several launches
template <typename T, int NP, typename F, typename PtrOf>
static int32_t launch_mt(wk_queue *q, const wk_opt_param_t *params, uint32_t n_params, F f, PtrOf ptr_of) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr uint64_t chunk_elems = (uint64_t)kThreads * kUnroll * VEC;
    for (uint32_t first = 0; first < n_params;) {
        MtTable<NP> tab{};
        uint64_t chunks = 0;
        while (first < n_params && tab.count < kMtMax) {
            const wk_opt_param_t &pr = params[first++];
            if (pr.n == 0) continue;
            for (int p = 0; p < NP; p++) {
                void *ptr = ptr_of(pr, p);
                if (!ptr) {
                    set_error("optimizer_step_multi: null buffer in parameter %u", first - 1);
                    return WK_ERR_INVALID_BUFFER;
                }
                if (!aligned16(ptr)) {
                    set_error("optimizer_step_multi: buffers must be 16-byte aligned (wk_malloc's are)");
                    return WK_ERR_INVALID_VALUE;
                }
                tab.p[p][tab.count] = ptr;
            }
            tab.n[tab.count] = pr.n;
            tab.first_chunk[tab.count] = chunks;
            uint64_t c = (pr.n + chunk_elems - 1) / chunk_elems;
            chunks += c;
            tab.count++;
        }
        if (tab.count == 0) continue;
        tab.first_chunk[tab.count] = chunks;
        for (int i = tab.count + 1; i <= kMtMax; i++) tab.first_chunk[i] = chunks;
        uint64_t blocks = chunks;
        static const int cap_env = env_knob("WK_MAP_CTAS_PER_SM", -1);
        const int ctas_per_sm = cap_env >= 0 ? cap_env : MtGridCap<F>::kCtasPerSm;
        const uint64_t cap = ctas_per_sm > 0 ? (uint64_t)q->sm_count * ctas_per_sm : 0x7fffffffull;
        if (blocks > cap) blocks = cap;
        mt_map_kernel<T, NP, F><<<(unsigned)blocks, kThreads, 0, q->stream>>>(tab, f);
        WK_CHECK_LAUNCH();
    }
    return WK_OK;
}

// axpy.zig:66-91 isSubstracting
template <typename T> static bool is_subtracting(const void *alpha) {
    if (std::is_unsigned<T>::value) return false;
    if (std::is_same<T, float>::value) return fabsf(*(const float *)alpha + 1.0f) < FLT_EPSILON;
    if (std::is_same<T, double>::value) return fabs(*(const double *)alpha + 1.0) < DBL_EPSILON;
    return *(const T *)alpha == (T)-1;
}
// complex: alpha.real == -1 and alpha.imag in {0, -1} (axpy.zig:75-76,84 -- the {-1,-1} case is the reference's quirk Q3)
template <typename BT> static bool is_subtracting_cx(const void *alpha) {
    if (std::is_unsigned<BT>::value) return false;
    const BT *c = (const BT *)alpha;
    if (std::is_same<BT, float>::value)
        return fabsf((float)c[0] + 1.0f) < FLT_EPSILON && (fabsf((float)c[1]) < FLT_EPSILON || fabsf((float)c[1] + 1.0f) < FLT_EPSILON);
    if (std::is_same<BT, double>::value)
        return fabs((double)c[0] + 1.0) < DBL_EPSILON && (fabs((double)c[1]) < DBL_EPSILON || fabs((double)c[1] + 1.0) < DBL_EPSILON);
    return c[0] == (BT)-1 && (c[1] == (BT)0 || c[1] == (BT)-1);
}
template <typename T> static bool is_sub(const void *alpha) {
    if (IsCx<T>::value) return is_subtracting_cx<typename IsCx<T>::base>(alpha);
    return is_subtracting<typename IsCx<T>::base>(alpha);
}

}  // namespace wk

using namespace wk;

WK_API int32_t wk_axpy(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *alpha,
                       const void *x, uint64_t xrp, uint64_t xsp, void *y, uint64_t yrp, uint64_t ysp) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        Ptrs<2> p{{const_cast<void *>(x), y}};
        Pitches<2> pit{{xrp, yrp}, {xsp, ysp}};
        if (!alpha) return launch_map3d<scalar_t, 2>(q, p, pit, depth, rows, cols, AxpyF<scalar_t, 0>{acc_zero<scalar_t>()});
        if (is_sub<scalar_t>(alpha)) return launch_map3d<scalar_t, 2>(q, p, pit, depth, rows, cols, AxpyF<scalar_t, 2>{acc_zero<scalar_t>()});
        return launch_map3d<scalar_t, 2>(q, p, pit, depth, rows, cols, AxpyF<scalar_t, 1>{load_scalar<scalar_t>(alpha)});
    });
}

WK_API int32_t wk_scal(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, const void *alpha, void *x,
                       uint64_t xrp, uint64_t xsp) {
    WK_CHECK_QUEUE(q);
    if (!alpha) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        Ptrs<1> p{{x}};
        Pitches<1> pit{{xrp}, {xsp}};
        return launch_map3d<scalar_t, 1>(q, p, pit, depth, rows, cols, ScalF<scalar_t>{load_scalar<scalar_t>(alpha)});
    });
}

WK_API int32_t wk_hadamard(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, void *x, uint64_t xrp,
                           uint64_t xsp, const void *y, uint64_t yrp, uint64_t ysp) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        Ptrs<2> p{{x, const_cast<void *>(y)}};
        Pitches<2> pit{{xrp, yrp}, {xsp, ysp}};
        return launch_map3d<scalar_t, 2>(q, p, pit, depth, rows, cols, HadamardF<scalar_t>{});
    });
}

WK_API int32_t wk_fill(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, void *buf, uint64_t rp,
                       uint64_t sp, const void *scalar) {
    WK_CHECK_QUEUE(q);
    if (!scalar) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        Ptrs<1> p{{buf}};
        Pitches<1> pit{{rp}, {sp}};
        return launch_map3d<scalar_t, 1>(q, p, pit, depth, rows, cols, FillF<scalar_t>{load_host<scalar_t>(scalar)});
    });
}

namespace wk {
// memory.copy of a dense span (src/tensor/memory/copy.zig:85-98): the streaming map template with a "copy" functor --
// 6.9 TB/s against the 6.4 TB/s of cudaMemcpyAsync device-to-device on this chip (profiles/sweep_stream_r02n.md)
struct CopyF {
    static constexpr unsigned kRead = 2, kWrite = 1;
    __device__ __forceinline__ void operator()(float (&v)[2], uint64_t) const { v[0] = v[1]; }
};
int32_t copy_dense(wk_queue *q, void *dst, const void *src, size_t bytes) {
    Ptrs<2> p{{dst, const_cast<void *>(src)}};
    return launch_map<float, 2>(q, p, bytes / 4, CopyF{});
}
}  // namespace wk

WK_API int32_t wk_uniform(wk_queue *q, int32_t dtype, uint64_t depth, uint64_t rows, uint64_t cols, void *buf, uint64_t rp,
                          uint64_t sp, uint64_t seed, const void *minp, const void *maxp) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        using base_t = typename IsCx<scalar_t>::base;  // min / max / range are scalars of the base type (uniform.zig:64-65)
        Ptrs<1> p{{buf}};
        Pitches<1> pit{{rp}, {sp}};
        // the hash is taken over the PADDED linear index (uniform.cl:73): the dense fast paths are only valid when
        // padded index == dense index (the same density test as launch_map3d's).
        const bool row_dense = rows == 1 || rp == cols;
        const bool dense = buf && depth && rows && cols && row_dense && (depth == 1 || sp == rows * cols) && aligned16(buf);
        constexpr uint64_t per_elem = sizeof(scalar_t) / sizeof(base_t);  // 2 for complex: the flat component index is what hashes
        if (minp || maxp) {
            // uniform.zig:96-106: missing bound = type min / max (floats: -floatMax / floatMax); range = max - min in T
            base_t mn = std::numeric_limits<base_t>::lowest(), mx = std::numeric_limits<base_t>::max();
            if (minp) mn = *(const base_t *)minp;
            if (maxp) mx = *(const base_t *)maxp;
            const base_t range = (base_t)(mx - mn);
            if (dense) return launch_uniform_dense<base_t, true>(q, (base_t *)buf, depth * rows * cols * per_elem, seed, mn, range);
            return launch_map3d<scalar_t, 1>(q, p, pit, depth, rows, cols, UniformF<scalar_t, true>{seed, mn, range});
        }
        if (dense) return launch_uniform_dense<base_t, false>(q, (base_t *)buf, depth * rows * cols * per_elem, seed, 0, 0);
        return launch_map3d<scalar_t, 1>(q, p, pit, depth, rows, cols, UniformF<scalar_t, false>{seed, 0, 0});
    });
}

WK_API int32_t wk_unary(wk_queue *q, int32_t dtype, int32_t op, void *x, uint64_t n) {
    WK_CHECK_QUEUE(q);
    if (op == WK_OP_SIGMOID && dtype >= 10) {  // Sigmoid.run is f32/f64 only (sigmoid.zig:21-24 TypeNotSupported)
        set_error("sigmoid: dtype %d not supported", dtype);
        return WK_ERR_TYPE_NOT_SUPPORTED;
    }
    return WK_DISPATCH_FLOAT_CX(dtype, [&]() -> int32_t {
        Ptrs<1> p{{x}};
        switch (op) {
            case WK_OP_SIN: return launch_map<scalar_t, 1>(q, p, n, UnaryF<scalar_t, WK_OP_SIN>{});
            case WK_OP_COS: return launch_map<scalar_t, 1>(q, p, n, UnaryF<scalar_t, WK_OP_COS>{});
            case WK_OP_TAN: return launch_map<scalar_t, 1>(q, p, n, UnaryF<scalar_t, WK_OP_TAN>{});
            case WK_OP_SINH: return launch_map<scalar_t, 1>(q, p, n, UnaryF<scalar_t, WK_OP_SINH>{});
            case WK_OP_COSH: return launch_map<scalar_t, 1>(q, p, n, UnaryF<scalar_t, WK_OP_COSH>{});
            case WK_OP_TANH: return launch_map<scalar_t, 1>(q, p, n, UnaryF<scalar_t, WK_OP_TANH>{});
            case WK_OP_SIGMOID: return launch_map<scalar_t, 1>(q, p, n, UnaryF<scalar_t, WK_OP_SIGMOID>{});
            default: set_error("unknown unary op %d", op); return WK_ERR_INVALID_VALUE;
        }
    });
}

WK_API int32_t wk_sigmoid_dev(wk_queue *q, int32_t dtype, const void *output, void *derivative, uint64_t n) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        Ptrs<2> p{{const_cast<void *>(output), derivative}};
        return launch_map<scalar_t, 2>(q, p, n, ActDevF<scalar_t, WK_ACT_SIGMOID>{});
    });
}

WK_API int32_t wk_tanh_dev(wk_queue *q, int32_t dtype, const void *output, void *derivative, uint64_t n) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        Ptrs<2> p{{const_cast<void *>(output), derivative}};
        return launch_map<scalar_t, 2>(q, p, n, ActDevF<scalar_t, WK_ACT_TANH>{});
    });
}

WK_API int32_t wk_act_backward(wk_queue *q, int32_t dtype, int32_t act, const void *output, void *derivative, void *sens,
                               uint64_t n) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        // pointer slot 1 must be non-null for the launcher; alias it to `sens` when the derivative is not stored
        Ptrs<3> p{{const_cast<void *>(output), derivative ? derivative : sens, sens}};
        if (act == WK_ACT_SIGMOID) {
            if (derivative) return launch_map<scalar_t, 3>(q, p, n, ActBackwardF<scalar_t, WK_ACT_SIGMOID, true>{});
            return launch_map<scalar_t, 3>(q, p, n, ActBackwardF<scalar_t, WK_ACT_SIGMOID, false>{});
        }
        if (act == WK_ACT_TANH) {
            if (derivative) return launch_map<scalar_t, 3>(q, p, n, ActBackwardF<scalar_t, WK_ACT_TANH, true>{});
            return launch_map<scalar_t, 3>(q, p, n, ActBackwardF<scalar_t, WK_ACT_TANH, false>{});
        }
        set_error("unknown activation %d", act);
        return WK_ERR_INVALID_VALUE;
    });
}

WK_API int32_t wk_bias_add(wk_queue *q, int32_t dtype, void *output, const void *bias, uint64_t row_pitch, uint64_t n) {
    WK_CHECK_QUEUE(q);
    if (!output || !bias) return WK_ERR_INVALID_BUFFER;
    if (row_pitch == 0) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_REAL(dtype, [&]() -> int32_t {
        constexpr uint64_t VEC = 16 / sizeof(scalar_t);
        const bool vec = aligned16(output) && aligned16(bias) && row_pitch % VEC == 0 && n % VEC == 0;
        const uint64_t n_units = vec ? n / VEC : n, rp_units = vec ? row_pitch / VEC : row_pitch;
        if (n_units == 0) return WK_OK;
        if (rp_units >= (1ull << 31)) {
            set_error("bias_add: row pitch too large");
            return WK_ERR_INVALID_VALUE;
        }
        uint64_t blocks = (n_units + (uint64_t)kThreads * kBiasUnroll - 1) / ((uint64_t)kThreads * kBiasUnroll);
        static const int bias_ctas = env_knob("WK_BIAS_CTAS_PER_SM", 0);  // 0 = one chunk per CTA (f32 0.94 -> 1.03; see MapGridCap)
        const uint64_t cap = bias_ctas > 0 ? (uint64_t)q->sm_count * bias_ctas : 0x7fffffffull;
        if (blocks > cap) blocks = cap;
        if (vec)
            bias_add_kernel<scalar_t, true><<<(unsigned)blocks, kThreads, 0, q->stream>>>((scalar_t *)output, (const scalar_t *)bias,
                                                                                         (uint32_t)rp_units, n_units);
        else
            bias_add_kernel<scalar_t, false><<<(unsigned)blocks, kThreads, 0, q->stream>>>((scalar_t *)output, (const scalar_t *)bias,
                                                                                          (uint32_t)rp_units, n_units);
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}

namespace wk {
// out = act(out + bias[col]) over n contiguous elements of rows `row_pitch` long (f32 / f64; act != NONE)
int32_t bias_act(wk_queue *q, int32_t dtype, void *output, const void *bias, uint64_t row_pitch, uint64_t n, int32_t act) {
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        constexpr uint64_t VEC = 16 / sizeof(scalar_t);
        const bool vec = aligned16(output) && aligned16(bias) && row_pitch % VEC == 0 && n % VEC == 0;
        const uint64_t n_units = vec ? n / VEC : n, rp_units = vec ? row_pitch / VEC : row_pitch;
        if (n_units == 0) return WK_OK;
        if (rp_units >= (1ull << 31)) return WK_ERR_INVALID_VALUE;
        uint64_t blocks = (n_units + (uint64_t)kThreads * kBiasUnroll - 1) / ((uint64_t)kThreads * kBiasUnroll);
        static const int bias_ctas = env_knob("WK_BIAS_CTAS_PER_SM", 0);  // 0 = one chunk per CTA (f32 0.94 -> 1.03; see MapGridCap)
        const uint64_t cap = bias_ctas > 0 ? (uint64_t)q->sm_count * bias_ctas : 0x7fffffffull;
        if (blocks > cap) blocks = cap;
        if (vec)
            bias_add_kernel<scalar_t, true, true><<<(unsigned)blocks, kThreads, 0, q->stream>>>((scalar_t *)output, (const scalar_t *)bias,
                                                                                               (uint32_t)rp_units, n_units, act);
        else
            bias_add_kernel<scalar_t, false, true><<<(unsigned)blocks, kThreads, 0, q->stream>>>((scalar_t *)output, (const scalar_t *)bias,
                                                                                                (uint32_t)rp_units, n_units, act);
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}
}  // namespace wk

WK_API int32_t wk_mse(wk_queue *q, int32_t dtype, const void *output, const void *expected, void *err, void *dev, uint64_t n) {
    WK_CHECK_QUEUE(q);
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        Ptrs<4> p{{const_cast<void *>(output), const_cast<void *>(expected), err, dev ? dev : err}};
        if (dev) return launch_map<scalar_t, 4>(q, p, n, MseF<scalar_t, true>{});
        return launch_map<scalar_t, 4>(q, p, n, MseF<scalar_t, false>{});
    });
}

WK_API int32_t wk_gdm(wk_queue *q, int32_t dtype, void *x, const void *g, void *v, const void *lr, const void *beta, uint64_t n) {
    WK_CHECK_QUEUE(q);
    if (!lr || !beta) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        Ptrs<3> p{{x, const_cast<void *>(g), v}};
        return launch_map<scalar_t, 3>(q, p, n, GdmF<scalar_t>{*(const scalar_t *)lr, *(const scalar_t *)beta});
    });
}

WK_API int32_t wk_adagrad(wk_queue *q, int32_t dtype, void *x, const void *g, void *h, const void *lr, uint64_t n) {
    WK_CHECK_QUEUE(q);
    if (!lr) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        Ptrs<3> p{{x, const_cast<void *>(g), h}};
        return launch_map<scalar_t, 3>(q, p, n, AdagradF<scalar_t>{*(const scalar_t *)lr});
    });
}

WK_API int32_t wk_rmsprop(wk_queue *q, int32_t dtype, void *x, const void *g, void *h, const void *lr, const void *gamma,
                          uint64_t n) {
    WK_CHECK_QUEUE(q);
    if (!lr || !gamma) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        Ptrs<3> p{{x, const_cast<void *>(g), h}};
        return launch_map<scalar_t, 3>(q, p, n, RmspropF<scalar_t>{*(const scalar_t *)lr, *(const scalar_t *)gamma});
    });
}

WK_API int32_t wk_adam(wk_queue *q, int32_t dtype, void *x, const void *g, void *m, void *v, const void *lr, const void *b1,
                       const void *b2, const void *eps, uint64_t t, uint64_t n) {
    WK_CHECK_QUEUE(q);
    if (!lr || !b1 || !b2 || !eps || t == 0) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        Ptrs<4> p{{x, const_cast<void *>(g), m, v}};
        const double B1 = (double)*(const scalar_t *)b1, B2 = (double)*(const scalar_t *)b2;
        AdamF<scalar_t> f{*(const scalar_t *)lr, *(const scalar_t *)b1, *(const scalar_t *)b2, *(const scalar_t *)eps,
                          (scalar_t)(1.0 / (1.0 - pow(B1, (double)t))), (scalar_t)(1.0 / (1.0 - pow(B2, (double)t)))};
        return launch_map<scalar_t, 4>(q, p, n, f);
    });
}

WK_API int32_t wk_identity(wk_queue *q, int32_t dtype, void *buf, uint64_t n_total, uint64_t size, uint64_t pitch_sum) {
    WK_CHECK_QUEUE(q);
    if (!buf) return WK_ERR_INVALID_BUFFER;
    return WK_DISPATCH_ALL(dtype, [&]() -> int32_t {
        WK_CUDA(cudaMemsetAsync(buf, 0, n_total * sizeof(scalar_t), q->stream));  // fill.zeroes, identity.zig:28
        identity_kernel<scalar_t><<<(unsigned)((size + 255) / 256), 256, 0, q->stream>>>((scalar_t *)buf, size, pitch_sum);
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}

WK_API int32_t wk_transpose_nd(wk_queue *q, int32_t dtype, uint32_t ndim, const void *src, const uint64_t *src_pitches, void *dst,
                               const uint64_t *dst_pitches, uint64_t row_pitch, uint64_t slice_pitch, uint64_t height, uint64_t cols,
                               uint64_t n_elements, uint32_t dim0, uint32_t dim1) {
    WK_CHECK_QUEUE(q);
    if (!src || !dst || !src_pitches || !dst_pitches) return WK_ERR_INVALID_BUFFER;
    if (ndim == 0 || ndim > 8 || dim0 >= ndim || dim1 >= ndim || row_pitch == 0 || slice_pitch == 0) return WK_ERR_INVALID_VALUE;
    if (n_elements == 0) return WK_OK;
    return WK_DISPATCH_SIZE(dtype, [&]() -> int32_t {
        TransposeNd p{};
        for (uint32_t i = 0; i < ndim; i++) {
            if (src_pitches[i] == 0) return WK_ERR_INVALID_VALUE;
            p.pa[i] = src_pitches[i];
            p.pb[i] = dst_pitches[i];
        }
        p.row_pitch = row_pitch; p.slice_pitch = slice_pitch; p.height = height; p.cols = cols; p.n = n_elements;
        p.ndim = ndim; p.dim0 = dim0; p.dim1 = dim1;
        uint64_t blocks = (n_elements + 255) / 256;
        const uint64_t cap = (uint64_t)q->sm_count * 16;
        if (blocks > cap) blocks = cap;
        transpose_nd_kernel<scalar_t><<<(unsigned)blocks, 256, 0, q->stream>>>((const scalar_t *)src, (scalar_t *)dst, p);
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}

WK_API int32_t wk_transpose2d(wk_queue *q, int32_t dtype, uint64_t rows, uint64_t cols, const void *src, uint64_t sp,
                              void *dst, uint64_t dp) {
    WK_CHECK_QUEUE(q);
    if (!src || !dst) return WK_ERR_INVALID_BUFFER;
    return WK_DISPATCH_SIZE(dtype, [&]() -> int32_t {
        if constexpr (sizeof(scalar_t) <= 8) {
            constexpr uint64_t VEC = 16 / sizeof(scalar_t);
            if (aligned16(src) && aligned16(dst) && rows % VEC == 0 && cols % VEC == 0 && sp % VEC == 0 && dp % VEC == 0 &&
                (rows + 63) / 64 <= 65535) {
                constexpr uint64_t TC = sizeof(scalar_t) == 8 ? 32 : 64;
                dim3 vgrid((unsigned)((cols + TC - 1) / TC), (unsigned)((rows + 63) / 64));
                transpose2d_vec_kernel<scalar_t><<<vgrid, 256, 0, q->stream>>>((const scalar_t *)src, sp, (scalar_t *)dst, dp, rows,
                                                                              cols);
                WK_CHECK_LAUNCH();
                return WK_OK;
            }
        }
        dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
        if (grid.y > 65535) {
            set_error("transpose2d: too many rows");
            return WK_ERR_INVALID_VALUE;
        }
        transpose2d_kernel<scalar_t><<<grid, dim3(32, 8), 0, q->stream>>>((const scalar_t *)src, sp, (scalar_t *)dst, dp,
                                                                        rows, cols);
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}

WK_API int32_t wk_optimizer_step_multi(wk_queue *q, int32_t dtype, int32_t kind, const wk_opt_param_t *params, uint32_t n_params,
                                       const void *lr, const void *h0, const void *h1, const void *h2, uint64_t t) {
    WK_CHECK_QUEUE(q);
    if (n_params == 0) return WK_OK;
    if (!params) return WK_ERR_INVALID_BUFFER;
    if (!lr) return WK_ERR_INVALID_VALUE;
    return WK_DISPATCH_FLOAT(dtype, [&]() -> int32_t {
        using S = scalar_t;
        auto xgs = [](const wk_opt_param_t &p, int slot) -> void * {  // {x, g, state0, state1}
            return slot == 0 ? p.x : slot == 1 ? const_cast<void *>(p.grad) : slot == 2 ? p.state0 : p.state1;
        };
        switch (kind) {
            case WK_OPT_GD: {  // blas.axpy(g, lr, w): pointer order {x = g, y = w}, the three variants of axpy.cl
                auto gw = [](const wk_opt_param_t &p, int slot) -> void * { return slot == 0 ? const_cast<void *>(p.grad) : p.x; };
                if (is_sub<S>(lr)) return launch_mt<S, 2>(q, params, n_params, AxpyF<S, 2>{acc_zero<S>()}, gw);
                return launch_mt<S, 2>(q, params, n_params, AxpyF<S, 1>{load_scalar<S>(lr)}, gw);
            }
            case WK_OPT_GDM:
                if (!h0) return WK_ERR_INVALID_VALUE;
                return launch_mt<S, 3>(q, params, n_params, GdmF<S>{*(const S *)lr, *(const S *)h0}, xgs);
            case WK_OPT_ADAGRAD:
                return launch_mt<S, 3>(q, params, n_params, AdagradF<S>{*(const S *)lr}, xgs);
            case WK_OPT_RMSPROP:
                if (!h0) return WK_ERR_INVALID_VALUE;
                return launch_mt<S, 3>(q, params, n_params, RmspropF<S>{*(const S *)lr, *(const S *)h0}, xgs);
            case WK_OPT_ADAM: {
                if (!h0 || !h1 || !h2 || t == 0) return WK_ERR_INVALID_VALUE;
                const double B1 = (double)*(const S *)h0, B2 = (double)*(const S *)h1;
                AdamF<S> f{*(const S *)lr, *(const S *)h0, *(const S *)h1, *(const S *)h2,
                           (S)(1.0 / (1.0 - pow(B1, (double)t))), (S)(1.0 / (1.0 - pow(B2, (double)t)))};
                return launch_mt<S, 4>(q, params, n_params, f, xgs);
            }
            default:
                set_error("optimizer_step_multi: unknown kind %d", kind);
                return WK_ERR_INVALID_VALUE;
        }
    });
}
