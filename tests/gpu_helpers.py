"""Shared helpers of the -m gpu parity tests: one Context/Pipeline per session, tensor pairs (CUDA tensor +
oracle tensor holding the same data), tolerance rules."""
import numpy as np

_state = {}


def wk():
    import wekua_b200

    return wekua_b200


def ctx_pipe():
    """Context.initFromDeviceType + Pipeline.init on command_queues[0], like every reference test"""
    if "ctx" not in _state:
        w = wk()
        _state["ctx"] = w.Context.init_from_device_type("all")
        _state["pipe"] = w.Pipeline.init(_state["ctx"].command_queues[0])
    return _state["ctx"], _state["pipe"]


def oracle_dev(oracle):
    """the device record our CUDA CommandQueue reports (vector width 1, .local memory)"""
    if "odev" not in _state:
        _state["odev"] = oracle.device("b200")
    return _state["odev"]


def make_pair(oracle, dtype, shape, data=None):
    """(cuda Tensor, oracle OTensor) with identical logical contents (zero padding)"""
    w = wk()
    ctx, pipe = ctx_pipe()
    t = w.Tensor.alloc(ctx, pipe, shape, dtype)
    o = oracle.OTensor(oracle_dev(oracle), dtype, shape)
    assert (t.row_pitch, t.slice_pitch, t.number_of_elements) == (o.layout.row_pitch, o.layout.slice_pitch, o.layout.number_of_elements)
    if data is not None:
        data = np.ascontiguousarray(data, dtype=dtype)
        w.tensor.memory.read_from_buffer(pipe, t, data)
        o.read_from(data)
    return t, o


def to_np(t):
    _, pipe = ctx_pipe()
    return wk().tensor.memory.to_numpy(pipe, t)


def padded(t):
    _, pipe = ctx_pipe()
    return wk().tensor.memory.padded_to_numpy(pipe, t)


def rand_data(rng, dtype, shape, lo=-1.0, hi=1.0):
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return rng.uniform(lo, hi, size=shape).astype(dtype)
    info = np.iinfo(dt)
    return rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)


def wrap_to_dtype(obj, dtype):
    """python-int array -> dtype with two's-complement wrap-around (exact arithmetic mod 2^bits)"""
    dt = np.dtype(dtype)
    bits = dt.itemsize * 8
    r = np.vectorize(lambda v: int(v) % (1 << bits), otypes=[object])(obj)
    if dt.kind == "i":
        r = np.vectorize(lambda v: v - (1 << bits) if v >= (1 << (bits - 1)) else v, otypes=[object])(r)
    return r.astype(dt)


def gemm_ideal(a, op_a, b, op_b, alpha, beta, c0):
    """fp64 (floats) / exact-integer-mod-2^bits (ints) ideal of C = alpha*op(A)op(B) + beta*C"""
    dt = c0.dtype
    A = a.T if op_a else a
    B = b.T if op_b else b
    if dt.kind == "f":
        r = A.astype(np.float64) @ B.astype(np.float64)
        if alpha is not None or beta is not None:
            r = (1.0 if alpha is None else float(dt.type(alpha))) * r
        if beta is not None:
            r = r + float(dt.type(beta)) * c0.astype(np.float64)
        return r
    A = A.astype(object)
    B = B.astype(object)
    r = A.dot(B)
    if alpha is not None or beta is not None:
        r = (1 if alpha is None else int(dt.type(alpha))) * r
    if beta is not None:
        r = r + int(dt.type(beta)) * c0.astype(object)
    return wrap_to_dtype(r, dt)


def gemm_float_bound(a, op_a, b, op_b, alpha, beta, c0, tol):
    """|c - c_ideal| <= (tol*K + 16) * eps * (|alpha| |A||B| + |beta||C|)   

    tol*K*eps*sum|a||b| is the classic worst-case bound of a length-K dot product in precision eps; the +16 eps
    covers the 3xTF32 split of an f32 product (each a*b carries <= 2^-20 relative error because the lo*lo term is
    dropped) and the alpha/beta epilogue."""
    dt = c0.dtype
    A = np.abs((a.T if op_a else a).astype(np.float64))
    B = np.abs((b.T if op_b else b).astype(np.float64))
    K = A.shape[1]
    bound = A @ B
    if alpha is not None:
        bound = abs(float(alpha)) * bound
    if beta is not None:
        bound = bound + abs(float(beta)) * np.abs(c0.astype(np.float64))
    return (tol * K + 16) * float(np.finfo(dt).eps) * bound + float(np.finfo(dt).tiny)
