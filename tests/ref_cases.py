"""Case tables transcribed from the reference's own unit tests, shared by the oracle KATs (CPU) and the
CUDA parity tests (GPU) so both read like the reference's tests.

gemm: src/blas/gemm.zig:945-1307 driven by src/blas/test_helpers.zig:88-247
      tuple = (kind, m, x, op_a, op_b, use_packing, alpha, beta); kind "AI" = A*I (x = k), "IB" = I*B (x = n)
"""
N, T = 0, 1

_AI_SHAPES = [(6, 10), (5, 7), (12, 20), (24, 8), (8, 24), (16, 48), (48, 16), (32, 96), (64, 128)]
_IB_SHAPES = [(10, 6), (7, 5), (20, 12), (8, 24), (48, 16), (96, 32), (128, 64)]

# "gemm cpu/gpu - all algorithms, non-complex" (gemm.zig:945-1006, :1133-1186) -- identical lists
GEMM_UNPACKED = (
    [("AI", m, k, N, N, False, None, None) for m, k in _AI_SHAPES]
    + [("IB", m, n, N, N, False, None, None) for m, n in _IB_SHAPES]
    + [("AI", m, k, N, N, False, 2, None) for m, k in [(6, 10), (12, 20), (24, 8), (16, 48)]]
    + [("AI", m, k, N, N, False, 2, 3) for m, k in [(6, 10), (12, 20), (24, 8), (16, 48)]]
    + [("AI", m, k, T, N, False, None, None) for m, k in [(6, 10), (12, 20), (24, 8), (16, 48)]]
    + [("AI", m, k, N, T, False, None, None) for m, k in [(6, 10), (12, 20)]]
    + [("IB", m, n, N, T, False, None, None) for m, n in [(10, 6), (20, 12), (8, 24)]]
)

# "gemm cpu/gpu - all algorithms with packing, non-complex" (gemm.zig:1054-1094, :1230-1270)
GEMM_PACKED = (
    [("AI", m, k, N, N, True, None, None) for m, k in [(6, 10), (5, 7), (12, 20), (24, 8), (8, 24), (16, 48), (32, 96), (64, 128)]]
    + [("IB", m, n, N, N, True, None, None) for m, n in [(10, 6), (7, 5), (20, 12), (48, 16)]]
    + [("AI", m, k, N, N, True, 2, 3) for m, k in [(6, 10), (12, 20), (24, 8)]]
    + [("AI", m, k, T, N, True, None, None) for m, k in [(6, 10), (12, 20)]]
    + [("IB", 10, 6, N, T, True, None, None)]
)

# "gemm cpu - all algorithms, complex" (gemm.zig:1008-1052): data {i+1, 0}, alpha {2,0}, beta {3,0}; complex floats only
GEMM_COMPLEX_CPU = (
    [("AI", m, k, N, N, False, None, None) for m, k in [(6, 10), (5, 7), (12, 20), (24, 8), (16, 48), (32, 96), (64, 128)]]
    + [("IB", m, n, N, N, False, None, None) for m, n in [(10, 6), (7, 5), (20, 12), (48, 16)]]
    + [("AI", m, k, N, N, False, 2, None) for m, k in [(6, 10), (12, 20)]]
    + [("AI", m, k, N, N, False, 2, 3) for m, k in [(6, 10), (12, 20)]]
    + [("AI", m, k, T, N, False, None, None) for m, k in [(6, 10), (12, 20)]]
    + [("IB", 10, 6, N, T, False, None, None)]
)
# "gemm gpu - all algorithms, complex" (gemm.zig:1188-1228): the same list without the alpha-only cases
GEMM_COMPLEX_GPU = [c for c in GEMM_COMPLEX_CPU if not (c[6] is not None and c[7] is None)]
# "gemm cpu|gpu - all algorithms with packing, complex" (gemm.zig:1096-1131, :1272-1307)
GEMM_COMPLEX_PACKED = (
    [("AI", m, k, N, N, True, None, None) for m, k in [(6, 10), (5, 7), (12, 20), (24, 8), (16, 48)]]
    + [("IB", m, n, N, N, True, None, None) for m, n in [(10, 6), (20, 12)]]
    + [("AI", m, k, N, N, True, 2, 3) for m, k in [(6, 10), (12, 20)]]
    + [("AI", 6, 10, T, N, True, None, None), ("IB", 10, 6, N, T, True, None, None)]
)

# "gemm - invalid shapes" (gemm.zig:900-941): A[4,5] x B[6,7] -> C[4,7] must fail with InvalidValue
GEMM_INVALID = [((4, 5), (6, 7), (4, 7), N, N)]


def gemm_case_shapes(kind, m, x, op_a, op_b):
    """(a_shape, b_shape, c_shape, which operand carries data) per test_helpers.zig:103-105, :185-187"""
    if kind == "AI":
        k = x
        a_shape = (k, m) if op_a == T else (m, k)
        return a_shape, (k, k), (m, k)
    n = x
    b_shape = (n, m) if op_b == T else (m, n)
    return (m, m), b_shape, (m, n)


def gemm_case_expected(kind, m, x, op_a, op_b, alpha, beta, np, dtype):
    """closed form of test_helpers.zig:46-74,149-164,231-246: alpha*data + beta*1"""
    if kind == "AI":
        k = x
        if op_a == T:
            data = (np.arange(k * m).reshape(k, m) + 1).T
        else:
            data = np.arange(m * k).reshape(m, k) + 1
    else:
        n = x
        if op_b == T:
            data = (np.arange(n * m).reshape(n, m) + 1).T
        else:
            data = np.arange(m * n).reshape(m, n) + 1
    exp = data.astype(dtype)
    if alpha is not None:
        exp = (dtype(alpha) * exp).astype(dtype)
    if beta is not None:
        exp = (exp + dtype(beta)).astype(dtype)
    return exp
