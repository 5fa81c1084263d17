/*
 * oracle/wekua_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See wekua_oracle.h.
 *
 * Host-side launch logic of the reference restated in C; the kernels are in kernels.inc.
 */
#include "wekua_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ kernel instantiations */
#define CAT_(a, b, c) a##_##b##_v##c
#define CAT(a, b, c) CAT_(a, b, c)
#define NAME(x) CAT(x, SUF, VW)

#define IS_COMPLEX 0
/* integers: lanes computed in uint64_t (wrap-around), truncated on store */
#ifndef WKO_FLOAT_ONLY
#define IS_FLOAT 0
#define ACC uint64_t
#define T int8_t
#define SUF i8
#include "vw_all.inc"
#undef T
#undef SUF
#define T uint8_t
#define SUF u8
#include "vw_all.inc"
#undef T
#undef SUF
#define T int16_t
#define SUF i16
#include "vw_all.inc"
#undef T
#undef SUF
#define T uint16_t
#define SUF u16
#include "vw_all.inc"
#undef T
#undef SUF
#define T int32_t
#define SUF i32
#include "vw_all.inc"
#undef T
#undef SUF
#define T uint32_t
#define SUF u32
#include "vw_all.inc"
#undef T
#undef SUF
#define T int64_t
#define SUF i64
#include "vw_all.inc"
#undef T
#undef SUF
#define T uint64_t
#define SUF u64
#include "vw_all.inc"
#undef T
#undef SUF
#undef ACC
#undef IS_FLOAT
#endif /* WKO_FLOAT_ONLY */

#define IS_FLOAT 1
#define T float
#define ACC float
#define SUF f32
#define FN_SIN sinf
#define FN_COS cosf
#define FN_TAN tanf
#define FN_SINH sinhf
#define FN_COSH coshf
#define FN_TANH tanhf
#define FN_EXP expf
#define FN_SQRT sqrtf
#include "vw_all.inc"
#undef T
#undef ACC
#undef SUF
#undef FN_SIN
#undef FN_COS
#undef FN_TAN
#undef FN_SINH
#undef FN_COSH
#undef FN_TANH
#undef FN_EXP
#undef FN_SQRT
#define T double
#define ACC double
#define SUF f64
#define FN_SIN sin
#define FN_COS cos
#define FN_TAN tan
#define FN_SINH sinh
#define FN_COSH cosh
#define FN_TANH tanh
#define FN_EXP exp
#define FN_SQRT sqrt
#include "vw_all.inc"
#undef T
#undef ACC
#undef SUF
#undef IS_FLOAT

#undef FN_SIN
#undef FN_COS
#undef FN_TAN
#undef FN_SINH
#undef FN_COSH
#undef FN_TANH
#undef FN_EXP
#undef FN_SQRT
#undef IS_COMPLEX
/* complex types (ids 10-19, src/core/types.zig:74-83): struct { BT re, im; }, vector width always 1 */
#define IS_COMPLEX 1
#define VW 1
#ifndef WKO_FLOAT_ONLY
#define IS_FLOAT 0
#define ACC uint64_t
#define CX_INT(BTYPE, SUFFIX)                        \
    typedef struct { BTYPE re, im; } cx_##SUFFIX;
CX_INT(int8_t, i8) CX_INT(uint8_t, u8) CX_INT(int16_t, i16) CX_INT(uint16_t, u16)
CX_INT(int32_t, i32) CX_INT(uint32_t, u32) CX_INT(int64_t, i64) CX_INT(uint64_t, u64)
#define T cx_i8
#define BT int8_t
#define SUF ci8
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#define T cx_u8
#define BT uint8_t
#define SUF cu8
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#define T cx_i16
#define BT int16_t
#define SUF ci16
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#define T cx_u16
#define BT uint16_t
#define SUF cu16
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#define T cx_i32
#define BT int32_t
#define SUF ci32
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#define T cx_u32
#define BT uint32_t
#define SUF cu32
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#define T cx_i64
#define BT int64_t
#define SUF ci64
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#define T cx_u64
#define BT uint64_t
#define SUF cu64
#include "kernels.inc"
#undef T
#undef BT
#undef SUF
#undef ACC
#undef IS_FLOAT
#endif /* WKO_FLOAT_ONLY */
#define IS_FLOAT 1
typedef struct { float re, im; } cx_f32;
typedef struct { double re, im; } cx_f64;
#define T cx_f32
#define BT float
#define ACC float
#define SUF cf32
#define FN_SIN sinf
#define FN_COS cosf
#define FN_SINH sinhf
#define FN_COSH coshf
#include "kernels.inc"
#undef T
#undef BT
#undef ACC
#undef SUF
#undef FN_SIN
#undef FN_COS
#undef FN_SINH
#undef FN_COSH
#define T cx_f64
#define BT double
#define ACC double
#define SUF cf64
#define FN_SIN sin
#define FN_COS cos
#define FN_SINH sinh
#define FN_COSH cosh
#include "kernels.inc"
#undef T
#undef BT
#undef ACC
#undef SUF
#undef FN_SIN
#undef FN_COS
#undef FN_SINH
#undef FN_COSH
#undef IS_FLOAT
#undef IS_COMPLEX
#undef VW

#define LOAD_INT(p) load_int_scalar(dtype, p)
#define LOAD_F32(p) (*(const float *)(p))
#define LOAD_F64(p) (*(const double *)(p))
#ifdef WKO_FLOAT_ONLY /* the -O3 cpu-baseline build: f32/f64 only */
#define INT_TYPES(X)
#else
#define INT_TYPES(X)                                                                                  \
    X(0, i8, int8_t, uint64_t, LOAD_INT) X(1, u8, uint8_t, uint64_t, LOAD_INT)                        \
    X(2, i16, int16_t, uint64_t, LOAD_INT) X(3, u16, uint16_t, uint64_t, LOAD_INT)                    \
    X(4, i32, int32_t, uint64_t, LOAD_INT) X(5, u32, uint32_t, uint64_t, LOAD_INT)                    \
    X(6, i64, int64_t, uint64_t, LOAD_INT) X(7, u64, uint64_t, uint64_t, LOAD_INT)
#endif
#define ALL_TYPES(X) INT_TYPES(X) X(8, f32, float, float, LOAD_F32) X(9, f64, double, double, LOAD_F64)
/* complex: X(id, suffix, storage struct, base type, lane type) */
#ifdef WKO_FLOAT_ONLY
#define CX_INT_TYPES(X)
#else
#define CX_INT_TYPES(X)                                                                              \
    X(10, ci8, cx_i8, int8_t, uint64_t) X(11, cu8, cx_u8, uint8_t, uint64_t)                         \
    X(12, ci16, cx_i16, int16_t, uint64_t) X(13, cu16, cx_u16, uint16_t, uint64_t)                   \
    X(14, ci32, cx_i32, int32_t, uint64_t) X(15, cu32, cx_u32, uint32_t, uint64_t)                   \
    X(16, ci64, cx_i64, int64_t, uint64_t) X(17, cu64, cx_u64, uint64_t, uint64_t)
#endif
#define CX_TYPES(X) CX_INT_TYPES(X) X(18, cf32, cx_f32, float, float) X(19, cf64, cx_f64, double, double)

static const uint64_t DTYPE_SIZE[20] = {1, 1, 2, 2, 4, 4, 8, 8, 4, 8, 2, 2, 4, 4, 8, 8, 16, 16, 8, 16};
#define IS_CX(d) ((d) >= 10)
/* load a complex scalar {re, im} of storage type TT into lanes of type ACCT */
#define CX_LOAD(vecT, TT, ACCT, p) ((vecT){(ACCT)((const TT *)(p))->re, (ACCT)((const TT *)(p))->im})

/* dispatch over dtype (d) and vector width (vw in {1,2,4,8,16}) */
#define DISPATCH_VW(fn, suf, vw, ...)               \
    switch (vw) {                                   \
        case 1: fn##_##suf##_v1(__VA_ARGS__); break;   \
        case 2: fn##_##suf##_v2(__VA_ARGS__); break;   \
        case 4: fn##_##suf##_v4(__VA_ARGS__); break;   \
        case 8: fn##_##suf##_v8(__VA_ARGS__); break;   \
        case 16: fn##_##suf##_v16(__VA_ARGS__); break; \
        default: return -1;                         \
    }

/* scalar -> ACC conversion per dtype */
static inline uint64_t load_int_scalar(int32_t d, const void *p) {
    switch (d) {
        case 0: return (uint64_t)*(const int8_t *)p;
        case 1: return (uint64_t)*(const uint8_t *)p;
        case 2: return (uint64_t)*(const int16_t *)p;
        case 3: return (uint64_t)*(const uint16_t *)p;
        case 4: return (uint64_t)*(const int32_t *)p;
        case 5: return (uint64_t)*(const uint32_t *)p;
        case 6: return (uint64_t)*(const int64_t *)p;
        default: return *(const uint64_t *)p;
    }
}

/* ------------------------------------------------------------------ devices */
void wko_device_cpu(wko_device *d, uint64_t vw_f32) {
    /* PoCL native widths after the min(...,16) clamp at command_queue.zig:97, SURVEY A.2 */
    const uint64_t w16[10] = {16, 16, 16, 16, 16, 16, 8, 8, 16, 8};
    const uint64_t w8[10] = {16, 16, 16, 16, 8, 8, 4, 4, 8, 4};
    memcpy(d->vector_widths, vw_f32 >= 16 ? w16 : w8, sizeof(w16));
    d->local_mem_type = WKO_MEM_GLOBAL;
    d->local_mem_size = 0;
    d->max_work_group_size = 4096;
}

void wko_device_gpu(wko_device *d) {
    for (int i = 0; i < 10; i++) d->vector_widths[i] = 1;
    d->local_mem_type = WKO_MEM_LOCAL;
    d->local_mem_size = 48 * 1024;
    d->max_work_group_size = 1024;
}

void wko_device_b200(wko_device *d) {
    for (int i = 0; i < 10; i++) d->vector_widths[i] = 1;
    d->local_mem_type = WKO_MEM_LOCAL;
    d->local_mem_size = 227 * 1024;
    d->max_work_group_size = 1024;
}

int32_t wko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* timing harness only: launchers such as torch.distributed.run export OMP_NUM_THREADS=1 into their children */
void wko_set_num_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ utils.zig:6-32 */
void wko_calculate_work_items(const uint64_t *global, uint64_t *local, uint64_t n, uint64_t max_wg) {
    const uint64_t max_per_cu = (uint64_t)pow((double)max_wg, 1.0 / (double)n);
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t g = global[i];
        if (g < max_per_cu) { local[i] = g; continue; }
        uint64_t li = max_per_cu;
        while (g % li != 0) li -= 1;
        local[i] = li;
    }
}

/* ------------------------------------------------------------------ work_configuration.zig:110-193 */
static int32_t init_gemm_algorithm(const wko_device *dev, int32_t dtype, uint64_t gwi_h, uint64_t gwi_w) {
    uint64_t max_block_length = 0;
    for (uint64_t bl = 2; bl < 128; bl *= 2)
        if (gwi_h % bl == 0 && gwi_w % bl == 0) max_block_length = bl;
    /* NB: the reference keeps the LAST dividing block length (:134-138), not the largest contiguous run */
    const uint64_t vw = IS_CX(dtype) ? 1 : dev->vector_widths[dtype]; /* work_configuration.zig:145-151 */
    int32_t algorithm = 0, idx = 0;
    for (uint64_t bl = 2; bl < 128; bl *= 2, idx++) {
        const uint64_t block_size = vw * bl * DTYPE_SIZE[dtype];
        int fit;
        if (dev->local_mem_type == WKO_MEM_LOCAL) fit = (block_size * bl * 2 <= dev->local_mem_size);
        else fit = (block_size * bl) <= 16 * 1024;
        if (bl <= max_block_length && fit) {
            if (dev->local_mem_type == WKO_MEM_LOCAL) {
                if (((bl * bl) / 4) < dev->max_work_group_size) algorithm = idx;
            } else {
                algorithm = idx;
            }
        }
    }
    return algorithm;
}

/* ------------------------------------------------------------------ tensor/main.zig:113-251 */
int32_t wko_layout_init(wko_layout *l, const wko_device *dev, int32_t dtype, const uint64_t *shape, uint64_t ndim,
                        int32_t vectors_enabled_cfg) {
    if (ndim == 0 || ndim > WKO_MAX_DIMS || dtype < 0 || dtype > 19) return -1;
    memset(l, 0, sizeof(*l));
    l->ndim = ndim;
    for (uint64_t i = 0; i < ndim; i++) {
        if (shape[i] == 0) return -1;
        l->shape[i] = shape[i];
        l->vl_shape[i] = shape[i];
    }
    int vectors_enabled = !IS_CX(dtype) && vectors_enabled_cfg != 0; /* main.zig:142 `!is_complex and config.vectors_enabled` */
    uint64_t vw = 1;
    if (vectors_enabled) {
        if (dev->vector_widths[dtype] > vw) vw = dev->vector_widths[dtype];
        vectors_enabled = vw > 1;
    }
    l->vectors_enabled = vectors_enabled;
    l->vector_width = vw;

    const uint64_t last = ndim - 1;
    const uint64_t pen = last > 0 ? last - 1 : 0; /* saturating `-|` */
    uint64_t depth = 1;
    for (uint64_t i = 0; i < pen; i++) depth *= shape[i];
    const uint64_t pen_size = ndim >= 2 ? shape[pen] : 1;
    const uint64_t last_size = shape[last];
    const uint64_t padded_pen = pen_size + (pen_size % 2);
    l->depth = depth;
    l->rows = pen_size;
    l->rows_padded = padded_pen;
    l->cols = last_size;
    l->number_of_elements_without_padding = depth * last_size * pen_size;

    uint64_t row_pitch = last_size;
    if (vectors_enabled && vw > 1) {
        const uint64_t rem = row_pitch % vw;
        if (rem > 0) row_pitch += vw - rem;
    }
    uint64_t rpv = row_pitch / vw;
    l->vl_shape[last] = rpv;
    const uint64_t rpv_rem = rpv % 2;
    rpv += rpv_rem;
    row_pitch += vw * rpv_rem;
    l->row_pitch = row_pitch;
    l->row_pitch_for_vectors = rpv;
    l->slice_pitch = row_pitch * padded_pen;
    l->number_of_elements = l->slice_pitch * depth;
    l->slice_pitch_for_vectors = l->slice_pitch / vw;
    l->number_of_vectors = l->number_of_elements / vw;

    const uint64_t ante = pen > 0 ? pen - 1 : 0;
    uint64_t pitch = l->number_of_elements;
    for (uint64_t i = 0; i < ante; i++) { pitch /= shape[i]; l->pitches[i] = pitch; }
    if (ndim >= 3) l->pitches[ante] = l->slice_pitch;
    if (ndim >= 2) l->pitches[pen] = row_pitch;
    l->pitches[last] = 1;

    l->gemm_algorithm = init_gemm_algorithm(dev, dtype, padded_pen, row_pitch);
    return 0;
}

/* ------------------------------------------------------------------ gemm.zig:42-58 */
int32_t wko_get_algorithm(int32_t default_algorithm, uint64_t k_size) {
    for (int32_t a = 5; a >= 0; a--) {
        const uint64_t bs = 2ull << a;
        if ((k_size % bs) == 0 && default_algorithm >= a) return a;
    }
    return -1; /* @panic("Unsupported block size") */
}

/* ------------------------------------------------------------------ gemm.zig:117-186 */
int32_t wko_packed_init(wko_packed_geom *g, const wko_device *dev, int32_t dtype, uint64_t n_size, uint64_t m_size,
                        uint64_t k_size, int32_t default_algorithm, int32_t vectors_enabled) {
    if (IS_CX(dtype)) vectors_enabled = 0; /* gemm.zig:129,150,181 `!is_complex and vectors_enabled` */
    const uint64_t vw = IS_CX(dtype) ? 1 : dev->vector_widths[dtype];
    uint64_t padded_k = k_size;
    if (vectors_enabled) {
        const uint64_t rem = padded_k % (vw * 2);
        if (rem > 0) padded_k += (vw * 2) - rem;
        padded_k /= vw;
    } else {
        padded_k += padded_k % 2;
    }
    const int32_t algo = wko_get_algorithm(default_algorithm, padded_k);
    if (algo < 0) return -1;
    const uint64_t bs = 2ull << algo;
    const uint64_t padded_n = n_size + (n_size % bs); /* sic (Q9) */
    const uint64_t padded_m = m_size + (m_size % bs);
    const uint64_t row_size = padded_k / bs;
    uint64_t col_size = bs * bs;
    if (vectors_enabled) col_size *= vw;
    const uint64_t sa[3] = {padded_n / bs, row_size, col_size};
    const uint64_t sb[3] = {padded_m / bs, row_size, col_size};
    if (wko_layout_init(&g->a, dev, dtype, sa, 3, vectors_enabled)) return -1;
    if (wko_layout_init(&g->b, dev, dtype, sb, 3, vectors_enabled)) return -1;
    g->algorithm = algo;
    g->vectors_enabled = vectors_enabled; /* `!is_complex and vectors_enabled`; vw==1 degenerates to scalar kernels */
    g->m_size = m_size;
    g->n_size = n_size;
    g->k_size = k_size;
    return 0;
}

#define PACK_CASE(suf, TT)                                                                                        \
    {                                                                                                             \
        DISPATCH_VW(pack, suf, vw, (const TT *)a, (TT *)pa, la->row_pitch, g->a.slice_pitch, g->a.row_pitch,      \
                    la->shape[0], la->shape[1], a_transpose, bs, g->a.shape[0], g->a.shape[1], g->a.shape[2]);    \
        DISPATCH_VW(pack, suf, vw, (const TT *)b, (TT *)pb, lb->row_pitch, g->b.slice_pitch, g->b.row_pitch,      \
                    lb->shape[0], lb->shape[1], b_transpose, bs, g->b.shape[0], g->b.shape[1], g->b.shape[2]);    \
    }                                                                                                             \
    break;

int32_t wko_pack(const wko_packed_geom *g, int32_t dtype, const void *a, const wko_layout *la, int32_t op_a,
                 const void *b, const wko_layout *lb, int32_t op_b, void *pa, void *pb) {
    /* validateTensors gemm.zig:250-270 */
    int valid = op_a ? (la->shape[1] == g->n_size && la->shape[0] == g->k_size)
                     : (la->shape[0] == g->n_size && la->shape[1] == g->k_size);
    valid &= op_b ? (lb->shape[1] == g->k_size && lb->shape[0] == g->m_size)
                  : (lb->shape[0] == g->k_size && lb->shape[1] == g->m_size);
    if (!valid) return -1;
    const int a_transpose = (op_a == 1);
    const int b_transpose = (op_b == 0); /* inverted for B, gemm.zig:286 */
    const uint64_t bs = 2ull << g->algorithm;
    const uint64_t vw = g->vectors_enabled ? g->a.vector_width : 1;
    /* the pack kernel is compiled with vectors_enabled => WK_VECTOR_WIDTH = device width */
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) case id: PACK_CASE(suf, TT)
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT)                                                                                 \
    case id:                                                                                                      \
        pack_##suf##_v1((const TT *)a, (TT *)pa, la->row_pitch, g->a.slice_pitch, g->a.row_pitch, la->shape[0],   \
                        la->shape[1], a_transpose, bs, g->a.shape[0], g->a.shape[1], g->a.shape[2]);              \
        pack_##suf##_v1((const TT *)b, (TT *)pb, lb->row_pitch, g->b.slice_pitch, g->b.row_pitch, lb->shape[0],   \
                        lb->shape[1], b_transpose, bs, g->b.shape[0], g->b.shape[1], g->b.shape[2]);              \
        break;
        CX_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------ blas.gemm */
static int validate_gemm(const wko_layout *la, const wko_layout *lb, const wko_layout *lc, int op_a, int op_b) {
    if (lc->ndim != 2 || la->ndim != 2 || lb->ndim != 2) return -1;
    const uint64_t a_m = la->shape[0], a_k = la->shape[1], b_k = lb->shape[0], b_n = lb->shape[1];
    const uint64_t c_m = lc->shape[0], c_n = lc->shape[1];
    int match;
    if (op_a) match = op_b ? (a_m == b_n && b_k == c_n && a_k == c_m) : (a_m == b_k && b_n == c_n && a_k == c_m);
    else match = op_b ? (a_k == b_n && b_k == c_n && a_m == c_m) : (a_k == b_k && b_n == c_n && a_m == c_m);
    return match ? 0 : -1;
}

#define GEMM_UNPACKED_CASE(suf, TT, ACCT, LOADS)                                                                  \
    {                                                                                                             \
        const ACCT al = has_alpha ? (alpha ? LOADS(alpha) : (ACCT)1) : (ACCT)0;                                   \
        const ACCT be = has_beta ? LOADS(beta) : (ACCT)0;                                                         \
        if (dev->local_mem_type == WKO_MEM_LOCAL) {                                                               \
            DISPATCH_VW(gemm_nxn_gpu, suf, vw, (const TT *)a, (const TT *)b, (TT *)c, a_rp, b_rp, lc->row_pitch,  \
                        k_size, op_a, op_b, has_alpha, has_beta, al, be, gh / 2, gw / 2, bs);                     \
        } else if (algo == 0) {                                                                                   \
            DISPATCH_VW(gemm_2x2, suf, vw, (const TT *)a, (const TT *)b, (TT *)c, a_rp, b_rp, lc->row_pitch,      \
                        k_size, op_a, op_b, has_alpha, has_beta, al, be, gh / 2, gw / 2);                         \
        } else {                                                                                                  \
            DISPATCH_VW(gemm_nxn, suf, vw, (const TT *)a, (const TT *)b, (TT *)c, a_rp, b_rp, lc->row_pitch,      \
                        k_size, op_a, op_b, has_alpha, has_beta, al, be, gh / bs, gw / bs, bs);                   \
        }                                                                                                         \
    }                                                                                                             \
    break;

#define GEMM_PACKED_CASE(suf, TT, ACCT, LOADS)                                                                    \
    {                                                                                                             \
        const ACCT al = has_alpha ? (alpha ? LOADS(alpha) : (ACCT)1) : (ACCT)0;                                   \
        const ACCT be = has_beta ? LOADS(beta) : (ACCT)0;                                                         \
        if (dev->local_mem_type == WKO_MEM_LOCAL) {                                                               \
            DISPATCH_VW(gemm_nxn_pack_gpu, suf, vw, (const TT *)pa, (const TT *)pb, (TT *)c, A_sp, A_rp, B_sp,    \
                        B_rp, lc->row_pitch, cols, has_alpha, has_beta, al, be, gh / 2, gw / 2, bs);              \
        } else if (algo == 0) {                                                                                   \
            DISPATCH_VW(gemm_2x2_pack, suf, vw, (const TT *)pa, (const TT *)pb, (TT *)c, A_sp, A_rp, B_sp, B_rp,  \
                        lc->row_pitch, cols, has_alpha, has_beta, al, be, gh / 2, gw / 2);                        \
        } else {                                                                                                  \
            DISPATCH_VW(gemm_nxn_pack, suf, vw, (const TT *)pa, (const TT *)pb, (TT *)c, A_sp, A_rp, B_sp, B_rp,  \
                        lc->row_pitch, cols, has_alpha, has_beta, al, be, gh / bs, gw / bs, bs);                  \
        }                                                                                                         \
    }                                                                                                             \
    break;

/* complex: alpha defaults to {1, 0} when only beta is given (gemm.zig:592-596) */
#define CX_SCALARS(suf, TT, ACCT)                                                                                 \
    const vec_##suf##_v1 al = has_alpha ? (alpha ? CX_LOAD(vec_##suf##_v1, TT, ACCT, alpha)                       \
                                                 : (vec_##suf##_v1){(ACCT)1, (ACCT)0})                             \
                                        : (vec_##suf##_v1){(ACCT)0, (ACCT)0};                                      \
    const vec_##suf##_v1 be = has_beta ? CX_LOAD(vec_##suf##_v1, TT, ACCT, beta) : (vec_##suf##_v1){(ACCT)0, (ACCT)0};

#define GEMM_UNPACKED_CX_CASE(suf, TT, ACCT)                                                                      \
    {                                                                                                             \
        CX_SCALARS(suf, TT, ACCT)                                                                                 \
        if (dev->local_mem_type == WKO_MEM_LOCAL)                                                                 \
            gemm_nxn_gpu_##suf##_v1((const TT *)a, (const TT *)b, (TT *)c, a_rp, b_rp, lc->row_pitch, k_size,     \
                                    op_a, op_b, has_alpha, has_beta, al, be, gh / 2, gw / 2, bs);                 \
        else if (algo == 0)                                                                                       \
            gemm_2x2_##suf##_v1((const TT *)a, (const TT *)b, (TT *)c, a_rp, b_rp, lc->row_pitch, k_size, op_a,   \
                                op_b, has_alpha, has_beta, al, be, gh / 2, gw / 2);                               \
        else                                                                                                      \
            gemm_nxn_##suf##_v1((const TT *)a, (const TT *)b, (TT *)c, a_rp, b_rp, lc->row_pitch, k_size, op_a,   \
                                op_b, has_alpha, has_beta, al, be, gh / bs, gw / bs, bs);                         \
    }                                                                                                             \
    break;

#define GEMM_PACKED_CX_CASE(suf, TT, ACCT)                                                                        \
    {                                                                                                             \
        CX_SCALARS(suf, TT, ACCT)                                                                                 \
        if (dev->local_mem_type == WKO_MEM_LOCAL)                                                                 \
            gemm_nxn_pack_gpu_##suf##_v1((const TT *)pa, (const TT *)pb, (TT *)c, A_sp, A_rp, B_sp, B_rp,         \
                                         lc->row_pitch, cols, has_alpha, has_beta, al, be, gh / 2, gw / 2, bs);   \
        else if (algo == 0)                                                                                       \
            gemm_2x2_pack_##suf##_v1((const TT *)pa, (const TT *)pb, (TT *)c, A_sp, A_rp, B_sp, B_rp,             \
                                     lc->row_pitch, cols, has_alpha, has_beta, al, be, gh / 2, gw / 2);           \
        else                                                                                                      \
            gemm_nxn_pack_##suf##_v1((const TT *)pa, (const TT *)pb, (TT *)c, A_sp, A_rp, B_sp, B_rp,             \
                                     lc->row_pitch, cols, has_alpha, has_beta, al, be, gh / bs, gw / bs, bs);     \
    }                                                                                                             \
    break;

static int32_t gemm_without_packing(const wko_device *dev, int32_t dtype, const void *alpha, const void *a,
                                    const wko_layout *la, int32_t op_a, const void *b, const wko_layout *lb,
                                    int32_t op_b, const void *beta, void *c, const wko_layout *lc) {
    /* gemm.zig:487-617 */
    const int has_alpha = (alpha != NULL || beta != NULL), has_beta = (beta != NULL);
    int vectors_enabled = la->vectors_enabled && lb->vectors_enabled;
    if (IS_CX(dtype) || dev->vector_widths[dtype] == 1) vectors_enabled = 0; /* gemm.zig:504 */
    else vectors_enabled &= (op_a == 0 && op_b == 1);
    uint64_t k_size;
    if (vectors_enabled) k_size = la->row_pitch_for_vectors;
    else { k_size = la->shape[1 - op_a]; k_size += k_size % 2; }
    const int32_t algo = wko_get_algorithm(lc->gemm_algorithm, k_size);
    if (algo < 0) return -1;
    const uint64_t a_rp = vectors_enabled ? la->row_pitch_for_vectors : la->row_pitch;
    const uint64_t b_rp = vectors_enabled ? lb->row_pitch_for_vectors : lb->row_pitch;
    const uint64_t vw = vectors_enabled ? dev->vector_widths[dtype] : 1;
    const uint64_t bs = 2ull << algo;
    const uint64_t gh = lc->rows_padded, gw = lc->row_pitch; /* work_configuration.zig:121-122 */
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) case id: GEMM_UNPACKED_CASE(suf, TT, ACCT, LOADS)
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT) case id: GEMM_UNPACKED_CX_CASE(suf, TT, ACCT)
        CX_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}

static int32_t gemm_with_packing(const wko_device *dev, int32_t dtype, const void *alpha, const void *a,
                                 const wko_layout *la, int32_t op_a, const void *b, const wko_layout *lb, int32_t op_b,
                                 const void *beta, void *c, const wko_layout *lc, int32_t pack_vectors) {
    /* PackedTensors.init(pipeline, c, K, vectors) gemm.zig:77-97 then gemmWithPacking :695-832 */
    const uint64_t K = la->shape[1 - op_a];
    wko_packed_geom g;
    if (wko_packed_init(&g, dev, dtype, lc->shape[0], lc->shape[1], K, lc->gemm_algorithm, pack_vectors)) return -1;
    const uint64_t es = DTYPE_SIZE[dtype];
    void *pa = calloc(g.a.number_of_elements, es), *pb = calloc(g.b.number_of_elements, es);
    if (!pa || !pb) { free(pa); free(pb); return -2; }
    int32_t rc = wko_pack(&g, dtype, a, la, op_a, b, lb, op_b, pa, pb);
    if (rc) { free(pa); free(pb); return rc; }
    const int has_alpha = (alpha != NULL || beta != NULL), has_beta = (beta != NULL);
    const int vectors_enabled = g.vectors_enabled && g.a.vector_width > 1;
    const int32_t algo = g.algorithm;
    const uint64_t bs = 2ull << algo;
    const uint64_t vw = vectors_enabled ? g.a.vector_width : 1;
    const uint64_t A_sp = vectors_enabled ? g.a.slice_pitch_for_vectors : g.a.slice_pitch;
    const uint64_t A_rp = vectors_enabled ? g.a.row_pitch_for_vectors : g.a.row_pitch;
    const uint64_t B_sp = vectors_enabled ? g.b.slice_pitch_for_vectors : g.b.slice_pitch;
    const uint64_t B_rp = vectors_enabled ? g.b.row_pitch_for_vectors : g.b.row_pitch;
    const uint64_t cols = g.a.shape[1];
    const uint64_t gh = lc->rows_padded, gw = lc->row_pitch;
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) case id: GEMM_PACKED_CASE(suf, TT, ACCT, LOADS)
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT) case id: GEMM_PACKED_CX_CASE(suf, TT, ACCT)
        CX_TYPES(X)
#undef X
        default: free(pa); free(pb); return -1;
    }
    free(pa);
    free(pb);
    return 0;
}

int32_t wko_gemm(const wko_device *dev, int32_t dtype, const void *alpha, const void *a, const wko_layout *la,
                 int32_t op_a, const void *b, const wko_layout *lb, int32_t op_b, const void *beta, void *c,
                 const wko_layout *lc, int32_t use_packing, int32_t pack_vectors) {
    if (validate_gemm(la, lb, lc, op_a, op_b)) return -1;
    if (use_packing) return gemm_with_packing(dev, dtype, alpha, a, la, op_a, b, lb, op_b, beta, c, lc, pack_vectors);
    return gemm_without_packing(dev, dtype, alpha, a, la, op_a, b, lb, op_b, beta, c, lc);
}

/* ------------------------------------------------------------------ blas.axpy */
static int same_shape(const wko_layout *a, const wko_layout *b) {
    if (a->ndim != b->ndim) return 0;
    for (uint64_t i = 0; i < a->ndim; i++)
        if (a->shape[i] != b->shape[i]) return 0;
    return 1;
}

/* axpy.zig:66-91 isSubstracting */
static int is_subtracting(int32_t dtype, const void *alpha) {
    switch (dtype) {
        case 0: return *(const int8_t *)alpha == -1;
        case 2: return *(const int16_t *)alpha == -1;
        case 4: return *(const int32_t *)alpha == -1;
        case 6: return *(const int64_t *)alpha == -1;
        case 8: return fabsf(*(const float *)alpha + 1.0f) < FLT_EPSILON;
        case 9: return fabs(*(const double *)alpha + 1.0) < DBL_EPSILON;
        /* complex: alpha.real == -1 and alpha.imag in {0, -1} (axpy.zig:75-76,84 -- SURVEY Q3) */
        case 10: { const int8_t *c = (const int8_t *)alpha; return c[0] == -1 && (c[1] == 0 || c[1] == -1); }
        case 12: { const int16_t *c = (const int16_t *)alpha; return c[0] == -1 && (c[1] == 0 || c[1] == -1); }
        case 14: { const int32_t *c = (const int32_t *)alpha; return c[0] == -1 && (c[1] == 0 || c[1] == -1); }
        case 16: { const int64_t *c = (const int64_t *)alpha; return c[0] == -1 && (c[1] == 0 || c[1] == -1); }
        case 18: { const float *c = (const float *)alpha;
                   return fabsf(c[0] + 1.0f) < FLT_EPSILON && (fabsf(c[1]) < FLT_EPSILON || fabsf(c[1] + 1.0f) < FLT_EPSILON); }
        case 19: { const double *c = (const double *)alpha;
                   return fabs(c[0] + 1.0) < DBL_EPSILON && (fabs(c[1]) < DBL_EPSILON || fabs(c[1] + 1.0) < DBL_EPSILON); }
        default: return 0; /* unsigned */
    }
}

#define AXPY_CASE(suf, TT, ACCT, LOADS)                                                                          \
    {                                                                                                            \
        const ACCT al = mode == 1 ? LOADS(alpha) : (ACCT)0;                                                      \
        if (vectors_enabled) {                                                                                   \
            DISPATCH_VW(axpy_1d, suf, lx->vector_width, (const TT *)x, (TT *)y, lx->number_of_vectors, mode, al); \
        } else {                                                                                                 \
            axpy_3d_##suf##_v1((const TT *)x, (TT *)y, lx->depth, lx->rows, lx->vl_shape[lx->ndim - 1],          \
                               lx->slice_pitch_for_vectors, lx->row_pitch_for_vectors,                           \
                               ly->slice_pitch_for_vectors, ly->row_pitch_for_vectors, mode, al);                \
        }                                                                                                        \
    }                                                                                                            \
    break;

int32_t wko_axpy(int32_t dtype, const void *x, const wko_layout *lx, const void *alpha, void *y, const wko_layout *ly) {
    if (!same_shape(lx, ly)) return -1;
    int mode = 0;
    if (alpha) mode = is_subtracting(dtype, alpha) ? 2 : 1;
    const int vectors_enabled = lx->vectors_enabled && ly->vectors_enabled;
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) case id: AXPY_CASE(suf, TT, ACCT, LOADS)
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT)                                                                                 \
    case id: {                                                                                                    \
        const vec_##suf##_v1 al = mode == 1 ? CX_LOAD(vec_##suf##_v1, TT, ACCT, alpha) : (vec_##suf##_v1){(ACCT)0, (ACCT)0}; \
        axpy_3d_##suf##_v1((const TT *)x, (TT *)y, lx->depth, lx->rows, lx->vl_shape[lx->ndim - 1],               \
                           lx->slice_pitch_for_vectors, lx->row_pitch_for_vectors, ly->slice_pitch_for_vectors,   \
                           ly->row_pitch_for_vectors, mode, al);                                                  \
    } break;
        CX_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}

#define HAD_CASE(suf, TT)                                                                                   \
    {                                                                                                       \
        if (vectors_enabled) {                                                                              \
            DISPATCH_VW(hadamard_1d, suf, lx->vector_width, (TT *)x, (const TT *)y, lx->number_of_vectors); \
        } else {                                                                                            \
            hadamard_3d_##suf##_v1((TT *)x, (const TT *)y, lx->depth, lx->rows, lx->vl_shape[lx->ndim - 1], \
                                   lx->slice_pitch_for_vectors, lx->row_pitch_for_vectors,                  \
                                   ly->slice_pitch_for_vectors, ly->row_pitch_for_vectors);                 \
        }                                                                                                   \
    }                                                                                                       \
    break;

int32_t wko_hadamard(int32_t dtype, void *x, const wko_layout *lx, const void *y, const wko_layout *ly) {
    if (!same_shape(lx, ly)) return -1;
    const int vectors_enabled = lx->vectors_enabled && ly->vectors_enabled;
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) case id: HAD_CASE(suf, TT)
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT)                                                                                 \
    case id:                                                                                                      \
        hadamard_3d_##suf##_v1((TT *)x, (const TT *)y, lx->depth, lx->rows, lx->vl_shape[lx->ndim - 1],           \
                               lx->slice_pitch_for_vectors, lx->row_pitch_for_vectors,                            \
                               ly->slice_pitch_for_vectors, ly->row_pitch_for_vectors);                           \
        break;
        CX_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}

#define SUM_CASE(suf, TT)                                                                                     \
    {                                                                                                         \
        const TT *src = (const TT *)x;                                                                        \
        if (last_dim > 1) {                                                                                   \
            DISPATCH_VW(row_sum, suf, lx->vector_width, (const TT *)x, (TT *)tmp, lx->depth, lx->rows,        \
                        lx->row_pitch_for_vectors, lx->slice_pitch_for_vectors, lx->rows);                    \
            src = (const TT *)tmp;                                                                            \
        }                                                                                                     \
        seq_sum_##suf##_v1(src, row_length, (TT *)out);                                                       \
    }                                                                                                         \
    break;

int32_t wko_sum(const wko_device *dev, int32_t dtype, const void *x, const wko_layout *lx, void *out) {
    /* basic.zig:131-203: device row sums into a fresh [1,row_length] tensor, then a host loop over the
     * first row_length mapped elements */
    uint64_t row_length = 1;
    for (uint64_t i = 0; i + 1 < lx->ndim; i++) row_length *= lx->shape[i];
    const uint64_t last_dim = lx->shape[lx->ndim - 1];
    void *tmp = NULL;
    if (last_dim > 1) {
        wko_layout lt;
        const uint64_t ts[2] = {1, row_length};
        if (wko_layout_init(&lt, dev, dtype, ts, 2, 1)) return -1;
        tmp = calloc(lt.number_of_elements, DTYPE_SIZE[dtype]);
        if (!tmp) return -2;
    }
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) case id: SUM_CASE(suf, TT)
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT)                                                                                 \
    case id: {                                                                                                    \
        const TT *src = (const TT *)x;                                                                            \
        if (last_dim > 1) {                                                                                       \
            row_sum_##suf##_v1((const TT *)x, (TT *)tmp, lx->depth, lx->rows, lx->row_pitch_for_vectors,          \
                               lx->slice_pitch_for_vectors, lx->rows);                                            \
            src = (const TT *)tmp;                                                                                \
        }                                                                                                         \
        seq_sum_##suf##_v1(src, row_length, (TT *)out);                                                           \
    } break;
        CX_TYPES(X)
#undef X
        default: free(tmp); return -1;
    }
    free(tmp);
    return 0;
}

int32_t wko_mean(const wko_device *dev, int32_t dtype, const void *x, const wko_layout *lx, void *out) {
    int32_t rc = wko_sum(dev, dtype, x, lx, out);
    if (rc) return rc;
    const uint64_t n = lx->number_of_elements_without_padding;
    switch (dtype) { /* basic.zig:229-236 : @divTrunc for ints, / for floats */
        case 0: *(int8_t *)out = (int8_t)(*(int8_t *)out / (int8_t)n); break;
        case 1: *(uint8_t *)out = (uint8_t)(*(uint8_t *)out / (uint8_t)n); break;
        case 2: *(int16_t *)out = (int16_t)(*(int16_t *)out / (int16_t)n); break;
        case 3: *(uint16_t *)out = (uint16_t)(*(uint16_t *)out / (uint16_t)n); break;
        case 4: *(int32_t *)out = *(int32_t *)out / (int32_t)n; break;
        case 5: *(uint32_t *)out = *(uint32_t *)out / (uint32_t)n; break;
        case 6: *(int64_t *)out = *(int64_t *)out / (int64_t)n; break;
        case 7: *(uint64_t *)out = *(uint64_t *)out / n; break;
        case 8: *(float *)out = *(float *)out / (float)n; break;
        case 9: *(double *)out = *(double *)out / (double)n; break;
        /* complex: each component on its own (basic.zig:216-227) */
#define CXM(id, BTT) case id: { BTT *c = (BTT *)out; c[0] = (BTT)(c[0] / (BTT)n); c[1] = (BTT)(c[1] / (BTT)n); } break;
        CXM(10, int8_t) CXM(11, uint8_t) CXM(12, int16_t) CXM(13, uint16_t) CXM(14, int32_t) CXM(15, uint32_t)
        CXM(16, int64_t) CXM(17, uint64_t) CXM(18, float) CXM(19, double)
#undef CXM
        default: return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------ float-only 1-D kernels */
#define F_DISPATCH(call32, call64) \
    if (dtype == 8) { call32; }    \
    else if (dtype == 9) { call64; } \
    else return -3; /* KernelsSet.Errors.TypeNotSupported (sigmoid.zig:21-24) */ \
    return 0;

int32_t wko_unary(int32_t dtype, void *x, uint64_t n, int32_t op) {
    if (dtype == 18 || dtype == 19) { /* trig.cl complex branches; sigmoid's complex branch is not restated */
        if (op < 0 || op > 5) return -3;
        if (dtype == 18) unary_cf32_v1((cx_f32 *)x, n, op);
        else unary_cf64_v1((cx_f64 *)x, n, op);
        return 0;
    }
    F_DISPATCH(unary_f32_v1((float *)x, n, op), unary_f64_v1((double *)x, n, op))
}
int32_t wko_sigmoid_dev(int32_t dtype, const void *o, void *d, uint64_t n) {
    F_DISPATCH(sigmoid_dev_f32_v1((const float *)o, (float *)d, n), sigmoid_dev_f64_v1((const double *)o, (double *)d, n))
}
int32_t wko_tanh_dev(int32_t dtype, const void *o, void *d, uint64_t n) {
    F_DISPATCH(tanh_dev_f32_v1((const float *)o, (float *)d, n), tanh_dev_f64_v1((const double *)o, (double *)d, n))
}
int32_t wko_bias(int32_t dtype, void *o, const void *b, uint64_t rp, uint64_t n) {
    F_DISPATCH(bias_f32_v1((float *)o, (const float *)b, rp, n), bias_f64_v1((double *)o, (const double *)b, rp, n))
}
int32_t wko_bias_step(int32_t dtype, const void *dev, void *bg, uint64_t rp, uint64_t rows, uint64_t n) {
    F_DISPATCH(bias_step_f32_v1((const float *)dev, (float *)bg, rp, rows, n),
               bias_step_f64_v1((const double *)dev, (double *)bg, rp, rows, n))
}
int32_t wko_mse(int32_t dtype, const void *o, const void *e, void *err, void *dev, uint64_t n) {
    F_DISPATCH(mse_f32_v1((const float *)o, (const float *)e, (float *)err, (float *)dev, n),
               mse_f64_v1((const double *)o, (const double *)e, (double *)err, (double *)dev, n))
}
int32_t wko_gdm(int32_t dtype, void *x, const void *g, void *v, const void *lr, const void *beta, uint64_t n) {
    F_DISPATCH(gdm_f32_v1((float *)x, (const float *)g, (float *)v, LOAD_F32(lr), LOAD_F32(beta), n),
               gdm_f64_v1((double *)x, (const double *)g, (double *)v, LOAD_F64(lr), LOAD_F64(beta), n))
}
int32_t wko_adagrad(int32_t dtype, void *x, const void *g, void *h, const void *lr, uint64_t n) {
    F_DISPATCH(adagrad_f32_v1((float *)x, (const float *)g, (float *)h, LOAD_F32(lr), n),
               adagrad_f64_v1((double *)x, (const double *)g, (double *)h, LOAD_F64(lr), n))
}
int32_t wko_rmsprop(int32_t dtype, void *x, const void *g, void *h, const void *lr, const void *gamma, uint64_t n) {
    F_DISPATCH(rmsprop_f32_v1((float *)x, (const float *)g, (float *)h, LOAD_F32(lr), LOAD_F32(gamma), n),
               rmsprop_f64_v1((double *)x, (const double *)g, (double *)h, LOAD_F64(lr), LOAD_F64(gamma), n))
}

/* ------------------------------------------------------------------ tensor utilities */
int32_t wko_fill(int32_t dtype, void *buf, const wko_layout *l, const void *scalar) {
    /* fill.zig:46-50 : 3-D over global_work_items_without_vectors = [depth, rows, cols] */
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) \
    case id: fill_##suf##_v1((TT *)buf, l->depth, l->rows, l->cols, l->row_pitch, l->slice_pitch, *(const TT *)scalar); break;
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT) \
    case id: fill_##suf##_v1((TT *)buf, l->depth, l->rows, l->cols, l->row_pitch, l->slice_pitch, *(const TT *)scalar); break;
        CX_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}

int32_t wko_identity(int32_t dtype, void *buf, const wko_layout *l) {
    const uint64_t size = l->shape[0];
    for (uint64_t i = 1; i < l->ndim; i++)
        if (l->shape[i] != size) return -1;
    memset(buf, 0, l->number_of_elements * DTYPE_SIZE[dtype]); /* fill.zeroes, identity.zig:28 */
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS) \
    case id: identity_##suf##_v1((TT *)buf, l->pitches, l->ndim, size); break;
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT) \
    case id: identity_##suf##_v1((TT *)buf, l->pitches, l->ndim, size); break;
        CX_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}

int32_t wko_transpose(int32_t dtype, const void *a, const wko_layout *la, void *b, const wko_layout *lb, uint64_t dim0,
                      uint64_t dim1) {
    /* transpose.zig:15-113 with (result_tensor=b, tensor=a) */
    if (la->ndim != lb->ndim) return -5;
    if (dim0 >= la->ndim || dim1 >= la->ndim) return -1;
    if (la->number_of_elements_without_padding != lb->number_of_elements_without_padding) return -5;
    if (lb->shape[dim0] != la->shape[dim1] || lb->shape[dim1] != la->shape[dim0]) return -1;
    if (dim0 == dim1) return -6; /* memory.copy path, not restated */
    const uint64_t d0 = dim0 > dim1 ? dim1 : dim0, d1 = dim0 > dim1 ? dim0 : dim1;
    const uint64_t width = la->cols, height = la->rows * la->row_pitch;
    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS)                                                                                              \
    case id:                                                                                                        \
        transpose_##suf##_v1((const TT *)a, la->pitches, (TT *)b, lb->pitches, la->row_pitch, la->slice_pitch, height, \
                             width, d0, d1, la->ndim, la->number_of_elements);                                       \
        break;
        ALL_TYPES(X)
#undef X
#define X(id, suf, TT, BTT, ACCT)                                                                                   \
    case id:                                                                                                        \
        transpose_##suf##_v1((const TT *)a, la->pitches, (TT *)b, lb->pitches, la->row_pitch, la->slice_pitch, height, \
                             width, d0, d1, la->ndim, la->number_of_elements);                                       \
        break;
        CX_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}

/* uniform.cl:32-54 (little-endian branch).  `seed2 << 32` discards every kept bit, so mixed == seed;
 * `^` binds looser than `*`: x1 = x0 ^ rotl(x0,49) ^ (rotl(x0,24) * PRIME). */
static inline uint64_t rotl64(uint64_t x, unsigned k) { return (x << k) | (x >> (64 - k)); }
uint64_t wko_xxhash64(uint64_t index, uint64_t global_seed) {
    const uint64_t F1 = 0x7C01812CF721AD1CULL, F2 = 0xDED46DE9839097DBULL, PRIME = 0x9FB21C651E98DF25ULL;
    const uint64_t seed2 = global_seed & 0xFFFFFFFF00000000ULL;
    const uint64_t mixed = global_seed ^ (seed2 << 32);
    const uint64_t key = (F1 ^ F2) - mixed;
    const uint64_t lo = index & 0xFFFFFFFFULL, hi = index >> 32;
    const uint64_t combined = (lo << 32) + hi;
    const uint64_t x0 = combined ^ key;
    const uint64_t x1 = x0 ^ rotl64(x0, 49) ^ (rotl64(x0, 24) * PRIME);
    const uint64_t x2 = x1 ^ (((x1 >> 35) + 8) * PRIME);
    return x2 ^ (x2 >> 28);
}

int32_t wko_uniform(int32_t dtype, void *buf, const wko_layout *l, uint64_t seed, const void *minp, const void *maxp) {
    const int range_defined = (minp != NULL || maxp != NULL);
    const uint64_t ncomp = IS_CX(dtype) ? 2 : 1;
    if (dtype < 0 || dtype > 19) return -1;
    dtype %= 10; /* from here on: the base type; `index` addresses base-type scalars */
    for (uint64_t i = 0; i < l->depth; i++)
        for (uint64_t j = 0; j < l->rows; j++)
            for (uint64_t k = 0; k < l->cols; k++) {
                const uint64_t eindex = i * l->slice_pitch + j * l->row_pitch + k;
                /* complex: the two components hash (index << 1) and (index << 1) + 1 (uniform.cl:80-93) and are
                 * stored next to each other; min/max/range are scalars of the base type (uniform.zig:64-65) */
                for (uint64_t comp = 0; comp < ncomp; comp++) {
                const uint64_t index = ncomp == 2 ? (eindex << 1) + comp : eindex;
                const uint64_t h = wko_xxhash64(index, seed);
                if (range_defined) {
                    /* uniform.cl:73-96 (cl_khr_fp64 branch): min + (double)h/ULONG_MAX * range, cast to T.
                     * range = max - min computed in T on the host (uniform.zig:106). */
                    const double normalized = ((double)h) / (double)UINT64_MAX;
                    switch (dtype) {
#define X(id, suf, TT, ACCT, LOADS)                                                                              \
    case id: {                                                                                      \
        const TT mn = minp ? *(const TT *)minp : MINV_##suf;                                        \
        const TT mx = maxp ? *(const TT *)maxp : MAXV_##suf;                                        \
        const TT range = (TT)(mx - mn);                                                             \
        ((TT *)buf)[index] = (TT)(mn + normalized * range);                                         \
    } break;
#define MINV_i8 INT8_MIN
#define MAXV_i8 INT8_MAX
#define MINV_u8 0
#define MAXV_u8 UINT8_MAX
#define MINV_i16 INT16_MIN
#define MAXV_i16 INT16_MAX
#define MINV_u16 0
#define MAXV_u16 UINT16_MAX
#define MINV_i32 INT32_MIN
#define MAXV_i32 INT32_MAX
#define MINV_u32 0
#define MAXV_u32 UINT32_MAX
#define MINV_i64 INT64_MIN
#define MAXV_i64 INT64_MAX
#define MINV_u64 0
#define MAXV_u64 UINT64_MAX
#define MINV_f32 (-FLT_MAX)
#define MAXV_f32 FLT_MAX
#define MINV_f64 (-DBL_MAX)
#define MAXV_f64 DBL_MAX
                        ALL_TYPES(X)
#undef X
                        default: return -1;
                    }
                } else {
                    switch (dtype) { /* uniform.cl:97-160 */
                        case 8: ((float *)buf)[index] = (float)h / ((float)UINT64_MAX); break;
                        case 9: ((double *)buf)[index] = (double)h / ((double)UINT64_MAX); break;
                        case 6: case 7: ((uint64_t *)buf)[index] = h; break;
                        case 4: case 5: ((uint32_t *)buf)[index] = (uint32_t)(h & 0xFFFFFFFFu); break;
                        case 2: case 3: ((uint16_t *)buf)[index] = (uint16_t)(h & 0xFFFFu); break;
                        case 0: case 1: ((uint8_t *)buf)[index] = (uint8_t)(h & 0xFFu); break;
                        default: return -1;
                    }
                }
                }
            }
    return 0;
}

/* memory/read_from_buffer.zig:13-63 / write_to_buffer.zig:13-63 : rect copy of the logical region */
int32_t wko_read_from_buffer(int32_t dtype, void *tb, const wko_layout *l, const void *host) {
    const uint64_t es = DTYPE_SIZE[dtype];
    for (uint64_t d = 0; d < l->depth; d++)
        for (uint64_t r = 0; r < l->rows; r++)
            memcpy((char *)tb + (d * l->slice_pitch + r * l->row_pitch) * es,
                   (const char *)host + ((d * l->rows + r) * l->cols) * es, l->cols * es);
    return 0;
}

int32_t wko_write_to_buffer(int32_t dtype, const void *tb, const wko_layout *l, void *host) {
    const uint64_t es = DTYPE_SIZE[dtype];
    for (uint64_t d = 0; d < l->depth; d++)
        for (uint64_t r = 0; r < l->rows; r++)
            memcpy((char *)host + ((d * l->rows + r) * l->cols) * es,
                   (const char *)tb + (d * l->slice_pitch + r * l->row_pitch) * es, l->cols * es);
    return 0;
}
