// complex.cu -- operand preparation of the complex GEMM (type ids 10-19, src/core/types.zig:74-83).
//
// The reference multiplies complex matrices element by element with a 3-multiplication product inside its SIMT
// tile loops (COMPLEX_MUL, src/core/wekua_cl_lib.cl:648-653; gemm_2x2.cl:158-163).  On B200 a complex contraction
// is ONE real contraction of twice the inner and twice the output width, which the real tensor-core kernels
// (3xTF32 tcgen05 for Complex(f32), DMMA for Complex(f64)) and the wrap-around SIMT kernel (complex integers: the
// identity below holds in any commutative ring, so results stay exact mod 2^bits) already run at their rooflines:
//
//   C'[m, 2n+c] = sum_{k,d} A'[m, 2k+d] * B''[2k+d, 2n+c]        A' = A's interleaved storage read as reals
//   B''[2k, 2n] = Br[k,n]   B''[2k, 2n+1] = Bi[k,n]   B''[2k+1, 2n] = -Bi[k,n]   B''[2k+1, 2n+1] = Br[k,n]
//
// so C' is C's interleaved storage, A (op_a = N) and C are used IN PLACE, and only B is rewritten: every stored
// row of B becomes two real rows -- the row itself and its "times i" partner -- in one streaming pass
// (cx_expand_b; a complex alpha is folded into the same pass).  op_a = T needs A's components on separate rows
// (cx_split_a).  Both passes are HBM-bound copies: read 2*R*C*s, write 4*R*C*s (expand) / 2*R*C*s (split) bytes.
#include "common.cuh"

namespace wk {

// rows x cols complex elements (pitch ldb) -> 2*rows real rows of 2*cols reals (pitch ldo reals).
//   op_b = N (B stored [K,N]):  out[2r]   = ( br,  bi) pairs      out[2r+1] = (-bi, br)
//   op_b = T (B stored [N,K]):  out[2r]   = ( br, -bi) pairs      out[2r+1] = ( bi, br)
// (the transposed form is B''^T restricted to the same two rows: rows index n, columns index (k,d)).
template <typename BT, bool SCALE>
__global__ void __launch_bounds__(256) cx_expand_b_kernel(const Cx<BT> *__restrict__ B, uint64_t ldb, Cx<BT> *__restrict__ out,
                                                          uint64_t ldo_cx, uint64_t rows, uint64_t cols, int op_b,
                                                          CxAcc<typename Acc<BT>::type> alpha) {
    using A = typename Acc<BT>::type;
    for (uint64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const Cx<BT> *src = B + r * ldb;
        Cx<BT> *o0 = out + (2 * r) * ldo_cx, *o1 = o0 + ldo_cx;
        for (uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x; c < cols; c += (uint64_t)gridDim.x * 256) {
            CxAcc<A> b = to_acc<Cx<BT>>(src[c]);
            if (SCALE) b = b * alpha;  // alpha * (A B) == A (B alpha): acc * alpha of gemm_2x2.cl:191-236 moved onto B
            const A nbi = (A)((A)0 - b.im);
            if (op_b == 0) {
                o0[c] = from_acc<Cx<BT>>(CxAcc<A>{b.re, b.im});
                o1[c] = from_acc<Cx<BT>>(CxAcc<A>{nbi, b.re});
            } else {
                o0[c] = from_acc<Cx<BT>>(CxAcc<A>{b.re, nbi});
                o1[c] = from_acc<Cx<BT>>(CxAcc<A>{b.im, b.re});
            }
        }
    }
}

// op_a = T: A stored [K, M] complex -> real [2K, M] (row 2k = real parts of row k, row 2k+1 = imaginary parts)
template <typename BT>
__global__ void __launch_bounds__(256) cx_split_a_kernel(const Cx<BT> *__restrict__ Ain, uint64_t lda, BT *__restrict__ out,
                                                         uint64_t ldo, uint64_t rows, uint64_t cols) {
    for (uint64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const Cx<BT> *src = Ain + r * lda;
        BT *o0 = out + (2 * r) * ldo, *o1 = o0 + ldo;
        for (uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x; c < cols; c += (uint64_t)gridDim.x * 256) {
            const Cx<BT> v = src[c];
            o0[c] = v.re;
            o1[c] = v.im;
        }
    }
}

static dim3 rows_grid(const wk_queue *q, uint64_t rows, uint64_t cols) {
    uint64_t gx = (cols + 255) / 256;
    if (gx > 64) gx = 64;
    uint64_t gy = ((uint64_t)q->sm_count * 8 + gx - 1) / gx;
    if (gy > rows) gy = rows;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)gx, (unsigned)gy);
}

int32_t cx_expand_b(wk_queue *q, int32_t base_dtype, int32_t op_b, uint64_t rows, uint64_t cols, const void *B, uint64_t ldb,
                    void *out, uint64_t ldo, const void *alpha) {
    return WK_DISPATCH_REAL(base_dtype, [&]() -> int32_t {
        using CT = Cx<scalar_t>;
        using A = typename Acc<scalar_t>::type;
        const dim3 grid = rows_grid(q, rows, cols);
        if (alpha)
            cx_expand_b_kernel<scalar_t, true><<<grid, 256, 0, q->stream>>>((const CT *)B, ldb, (CT *)out, ldo / 2, rows, cols, op_b,
                                                                           load_scalar<CT>(alpha));
        else
            cx_expand_b_kernel<scalar_t, false><<<grid, 256, 0, q->stream>>>((const CT *)B, ldb, (CT *)out, ldo / 2, rows, cols, op_b,
                                                                            CxAcc<A>{(A)0, (A)0});
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}

int32_t cx_split_a(wk_queue *q, int32_t base_dtype, uint64_t rows, uint64_t cols, const void *A, uint64_t lda, void *out,
                   uint64_t ldo) {
    return WK_DISPATCH_REAL(base_dtype, [&]() -> int32_t {
        cx_split_a_kernel<scalar_t><<<rows_grid(q, rows, cols), 256, 0, q->stream>>>((const Cx<scalar_t> *)A, lda, (scalar_t *)out, ldo,
                                                                                    rows, cols);
        WK_CHECK_LAUNCH();
        return WK_OK;
    });
}

}  // namespace wk
