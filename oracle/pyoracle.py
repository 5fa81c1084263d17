"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes/numpy front-end of the CPU restatement in wekua_oracle.c (see wekua_oracle.h for what each
entry point follows in the reference).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; wekua_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MAX_DIMS = 8
REAL_DTYPES = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]


def cx(base) -> np.dtype:
    """Complex(T) of src/core/types.zig:6-11 as a numpy structured dtype {re, im} (numpy has no complex integers)"""
    return np.dtype([("re", np.dtype(base)), ("im", np.dtype(base))])


CX_DTYPES = [cx(d) for d in REAL_DTYPES]
NP_DTYPES = REAL_DTYPES + CX_DTYPES  # index = getTypeIndex, types.zig:60-87
DTYPE_NAMES = ["i8", "u8", "i16", "u16", "i32", "u32", "i64", "u64", "f32", "f64"]
DTYPE_NAMES += ["c" + n for n in DTYPE_NAMES]
NO_TRANSPOSE, TRANSPOSE = 0, 1
UNARY_OPS = {"sin": 0, "cos": 1, "tan": 2, "sinh": 3, "cosh": 4, "tanh": 5, "sigmoid": 6}


def dtype_id(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.complex64:
        return 18
    if dt == np.complex128:
        return 19
    for i, d in enumerate(NP_DTYPES):
        if np.dtype(d) == dt:
            return i
    raise TypeError(f"unsupported dtype {dt}")


def is_complex(dt) -> bool:
    return dtype_id(dt) >= 10


def as_dtype(values, dt) -> np.ndarray:
    """host values -> contiguous array of element type dt; for Complex(T): real input becomes {v, 0}, python /
    numpy complex input becomes {re, im}, (..., 2)-shaped pairs are NOT guessed (build them with cx_pairs)"""
    dt = NP_DTYPES[dtype_id(dt)]
    dt = np.dtype(dt)
    a = np.asarray(values)
    if dt.names is None:
        return np.ascontiguousarray(a, dtype=dt)
    if a.dtype == dt:
        return np.ascontiguousarray(a)
    out = np.zeros(a.shape, dtype=dt)
    if a.dtype.kind == "c":
        out["re"], out["im"] = a.real, a.imag
    else:
        out["re"] = a
    return out


def cx_pairs(re, im, base) -> np.ndarray:
    out = np.zeros(np.shape(re), dtype=cx(base))
    out["re"], out["im"] = re, im
    return out


class Device(C.Structure):
    _fields_ = [
        ("vector_widths", C.c_uint64 * 10),
        ("local_mem_type", C.c_int32),
        ("local_mem_size", C.c_uint64),
        ("max_work_group_size", C.c_uint64),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("ndim", C.c_uint64),
        ("shape", C.c_uint64 * MAX_DIMS),
        ("vl_shape", C.c_uint64 * MAX_DIMS),
        ("pitches", C.c_uint64 * MAX_DIMS),
        ("number_of_elements", C.c_uint64),
        ("number_of_elements_without_padding", C.c_uint64),
        ("row_pitch", C.c_uint64),
        ("row_pitch_for_vectors", C.c_uint64),
        ("slice_pitch", C.c_uint64),
        ("slice_pitch_for_vectors", C.c_uint64),
        ("number_of_vectors", C.c_uint64),
        ("depth", C.c_uint64),
        ("rows", C.c_uint64),
        ("rows_padded", C.c_uint64),
        ("cols", C.c_uint64),
        ("vector_width", C.c_uint64),
        ("vectors_enabled", C.c_int32),
        ("gemm_algorithm", C.c_int32),
    ]


class PackedGeom(C.Structure):
    _fields_ = [
        ("a", Layout),
        ("b", Layout),
        ("algorithm", C.c_int32),
        ("vectors_enabled", C.c_int32),
        ("m_size", C.c_uint64),
        ("n_size", C.c_uint64),
        ("k_size", C.c_uint64),
    ]


def build(force: bool = False) -> None:
    """Compile the restatement with the committed Makefile (gcc only)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE, "-j3"], check=True, capture_output=True)


def _cpu_has(flag: str) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


_libs: dict[str, C.CDLL] = {}


def _bind(lib: C.CDLL) -> C.CDLL:
    lib.wko_xxhash64.restype = C.c_uint64
    lib.wko_xxhash64.argtypes = [C.c_uint64, C.c_uint64]
    lib.wko_num_threads.restype = C.c_int32
    if hasattr(lib, "wko_set_num_threads"):
        lib.wko_set_num_threads.restype = None
        lib.wko_set_num_threads.argtypes = [C.c_int32]
    u64, i32, vp = C.c_uint64, C.c_int32, C.c_void_p
    LP, DP = C.POINTER(Layout), C.POINTER(Device)
    sig = {
        "wko_layout_init": [LP, DP, i32, C.POINTER(u64), u64, i32],
        "wko_get_algorithm": [i32, u64],
        "wko_packed_init": [C.POINTER(PackedGeom), DP, i32, u64, u64, u64, i32, i32],
        "wko_pack": [C.POINTER(PackedGeom), i32, vp, LP, i32, vp, LP, i32, vp, vp],
        "wko_gemm": [DP, i32, vp, vp, LP, i32, vp, LP, i32, vp, vp, LP, i32, i32],
        "wko_axpy": [i32, vp, LP, vp, vp, LP],
        "wko_hadamard": [i32, vp, LP, vp, LP],
        "wko_sum": [DP, i32, vp, LP, vp],
        "wko_mean": [DP, i32, vp, LP, vp],
        "wko_unary": [i32, vp, u64, i32],
        "wko_sigmoid_dev": [i32, vp, vp, u64],
        "wko_tanh_dev": [i32, vp, vp, u64],
        "wko_bias": [i32, vp, vp, u64, u64],
        "wko_bias_step": [i32, vp, vp, u64, u64, u64],
        "wko_mse": [i32, vp, vp, vp, vp, u64],
        "wko_gdm": [i32, vp, vp, vp, vp, vp, u64],
        "wko_adagrad": [i32, vp, vp, vp, vp, u64],
        "wko_rmsprop": [i32, vp, vp, vp, vp, vp, u64],
        "wko_fill": [i32, vp, LP, vp],
        "wko_identity": [i32, vp, LP],
        "wko_transpose": [i32, vp, LP, vp, LP, u64, u64],
        "wko_uniform": [i32, vp, LP, u64, vp, vp],
        "wko_read_from_buffer": [i32, vp, LP, vp],
        "wko_write_to_buffer": [i32, vp, LP, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int32
        fn.argtypes = args
    lib.wko_calculate_work_items.restype = None
    lib.wko_calculate_work_items.argtypes = [C.POINTER(u64), C.POINTER(u64), u64, u64]
    for name in ("wko_device_gpu", "wko_device_b200"):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [DP]
    lib.wko_device_cpu.restype = None
    lib.wko_device_cpu.argtypes = [DP, u64]
    return lib


def lib(fast: bool = False) -> C.CDLL:
    """The checker build (all dtypes, -ffp-contract=off) or, with fast=True, the -O3 float-only
    timing build (AVX-512 variant when the CPU supports it)."""
    key = "fast" if fast else "check"
    if key in _libs:
        return _libs[key]
    if fast:
        name = "libwekua_oracle_fast512.so" if _cpu_has("avx512f") else "libwekua_oracle_fast.so"
    else:
        name = "libwekua_oracle.so"
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    _libs[key] = _bind(C.CDLL(path))
    return _libs[key]


def device(kind: str = "b200", vw_f32: int = 16) -> Device:
    d = Device()
    if kind == "cpu":
        lib().wko_device_cpu(C.byref(d), vw_f32)
    elif kind == "gpu":
        lib().wko_device_gpu(C.byref(d))
    elif kind == "b200":
        lib().wko_device_b200(C.byref(d))
    else:
        raise ValueError(kind)
    return d


def calculate_work_items(global_items, max_wg: int):
    n = len(global_items)
    g = (C.c_uint64 * n)(*global_items)
    l = (C.c_uint64 * n)()
    lib().wko_calculate_work_items(g, l, n, max_wg)
    return list(l)


def _ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _scalar(dt, v):
    if v is None:
        return None
    dt = np.dtype(dt)
    if dt.names is not None:
        if isinstance(v, tuple):
            return cx_pairs([v[0]], [v[1]], dt["re"])
        return as_dtype(np.array([v]), dt)
    return np.array([v]).astype(dt)


class OTensor:
    """A reference-layout tensor on the host: padded numpy buffer + the reference's layout record
    (src/tensor/main.zig:62-279)."""

    def __init__(self, dev: Device, dtype, shape, vectors_enabled: bool = True, fast: bool = False):
        self.dev = dev
        self.dtype = dtype_id(dtype)
        self.np_dtype = np.dtype(NP_DTYPES[self.dtype])
        self.base_dtype = np.dtype(REAL_DTYPES[self.dtype % 10])
        self.layout = Layout()
        self._lib = lib(fast)
        shp = (C.c_uint64 * len(shape))(*shape)
        rc = self._lib.wko_layout_init(C.byref(self.layout), C.byref(dev), self.dtype, shp, len(shape), int(vectors_enabled))
        if rc:
            raise ValueError("InvalidValue")
        self.shape = tuple(shape)
        self.buf = np.zeros(self.layout.number_of_elements, dtype=self.np_dtype)  # Tensor.alloc zero-fills

    # memory.readFromBuffer / writeToBuffer
    def read_from(self, host):
        host = as_dtype(host, self.np_dtype).reshape(-1)
        if host.size != self.layout.number_of_elements_without_padding:
            raise ValueError("InvalidBuffer")
        self._lib.wko_read_from_buffer(self.dtype, _ptr(self.buf), C.byref(self.layout), _ptr(host))
        return self

    def to_host(self):
        out = np.empty(self.layout.number_of_elements_without_padding, dtype=self.np_dtype)
        self._lib.wko_write_to_buffer(self.dtype, _ptr(self.buf), C.byref(self.layout), _ptr(out))
        return out.reshape(self.shape)

    def fill(self, v):
        s = _scalar(self.np_dtype, v)
        self._lib.wko_fill(self.dtype, _ptr(self.buf), C.byref(self.layout), _ptr(s))
        return self

    def identity(self):
        if self._lib.wko_identity(self.dtype, _ptr(self.buf), C.byref(self.layout)):
            raise ValueError("InvalidValue")
        return self

    def uniform(self, seed, lo=None, hi=None):
        a, b = _scalar(self.base_dtype, lo), _scalar(self.base_dtype, hi)  # bounds are scalars of the base type
        self._lib.wko_uniform(self.dtype, _ptr(self.buf), C.byref(self.layout), seed, _ptr(a), _ptr(b))
        return self


def gemm(alpha, a: OTensor, op_a: int, b: OTensor, op_b: int, beta, c: OTensor, packed: bool = False,
         pack_vectors: bool = True, fast: bool = False) -> None:
    al, be = _scalar(c.np_dtype, alpha), _scalar(c.np_dtype, beta)
    rc = lib(fast).wko_gemm(C.byref(c.dev), c.dtype, _ptr(al), _ptr(a.buf), C.byref(a.layout), op_a, _ptr(b.buf),
                            C.byref(b.layout), op_b, _ptr(be), _ptr(c.buf), C.byref(c.layout), int(packed),
                            int(pack_vectors))
    if rc == -1:
        raise ValueError("InvalidValue")
    if rc:
        raise RuntimeError(f"wko_gemm rc={rc}")


def packed_geom(dev: Device, dtype, n, m, k, default_algorithm, vectors_enabled=True) -> PackedGeom:
    g = PackedGeom()
    rc = lib().wko_packed_init(C.byref(g), C.byref(dev), dtype_id(dtype), n, m, k, default_algorithm, int(vectors_enabled))
    if rc:
        raise ValueError("InvalidValue")
    return g


def pack(g: PackedGeom, a: OTensor, op_a: int, b: OTensor, op_b: int):
    pa = np.zeros(g.a.number_of_elements, dtype=a.np_dtype)
    pb = np.zeros(g.b.number_of_elements, dtype=a.np_dtype)
    rc = lib().wko_pack(C.byref(g), a.dtype, _ptr(a.buf), C.byref(a.layout), op_a, _ptr(b.buf), C.byref(b.layout), op_b,
                        _ptr(pa), _ptr(pb))
    if rc:
        raise ValueError("InvalidValue")
    return pa, pb


def axpy(x: OTensor, alpha, y: OTensor, fast: bool = False) -> None:
    al = _scalar(x.np_dtype, alpha)
    if lib(fast).wko_axpy(x.dtype, _ptr(x.buf), C.byref(x.layout), _ptr(al), _ptr(y.buf), C.byref(y.layout)):
        raise ValueError("UnqualTensorsShape")


def hadamard(x: OTensor, y: OTensor) -> None:
    if lib().wko_hadamard(x.dtype, _ptr(x.buf), C.byref(x.layout), _ptr(y.buf), C.byref(y.layout)):
        raise ValueError("UnqualTensorsShape")


def tsum(x: OTensor):
    out = np.zeros(1, dtype=x.np_dtype)
    lib().wko_sum(C.byref(x.dev), x.dtype, _ptr(x.buf), C.byref(x.layout), _ptr(out))
    return out[0]


def mean(x: OTensor):
    out = np.zeros(1, dtype=x.np_dtype)
    lib().wko_mean(C.byref(x.dev), x.dtype, _ptr(x.buf), C.byref(x.layout), _ptr(out))
    return out[0]


def unary(x: OTensor, op: str) -> None:
    if lib().wko_unary(x.dtype, _ptr(x.buf), x.layout.number_of_elements, UNARY_OPS[op]):
        raise TypeError("TypeNotSupported")


def sigmoid_dev(out: OTensor, dev: OTensor) -> None:
    lib().wko_sigmoid_dev(out.dtype, _ptr(out.buf), _ptr(dev.buf), out.layout.number_of_elements)


def tanh_dev(out: OTensor, dev: OTensor) -> None:
    lib().wko_tanh_dev(out.dtype, _ptr(out.buf), _ptr(dev.buf), out.layout.number_of_elements)


def bias(out: OTensor, b: OTensor) -> None:
    # linear.zig:451-470 with vectors disabled: row_pitch, number_of_elements
    lib().wko_bias(out.dtype, _ptr(out.buf), _ptr(b.buf), out.layout.row_pitch, out.layout.number_of_elements)


def bias_step(sens: OTensor, bias_grad: OTensor) -> None:
    # linear.zig:547-570
    lib().wko_bias_step(sens.dtype, _ptr(sens.buf), _ptr(bias_grad.buf), sens.layout.row_pitch_for_vectors,
                        sens.layout.shape[0], bias_grad.layout.row_pitch_for_vectors)


def mse(out: OTensor, expected: OTensor, err: OTensor, dev: OTensor | None) -> None:
    lib().wko_mse(out.dtype, _ptr(out.buf), _ptr(expected.buf), _ptr(err.buf), _ptr(dev.buf) if dev is not None else None,
                  out.layout.number_of_elements)


def gdm(x: OTensor, g: OTensor, v: OTensor, lr, beta) -> None:
    lib().wko_gdm(x.dtype, _ptr(x.buf), _ptr(g.buf), _ptr(v.buf), _ptr(_scalar(x.np_dtype, lr)),
                  _ptr(_scalar(x.np_dtype, beta)), x.layout.number_of_elements)


def adagrad(x: OTensor, g: OTensor, h: OTensor, lr) -> None:
    lib().wko_adagrad(x.dtype, _ptr(x.buf), _ptr(g.buf), _ptr(h.buf), _ptr(_scalar(x.np_dtype, lr)),
                      x.layout.number_of_elements)


def rmsprop(x: OTensor, g: OTensor, h: OTensor, lr, gamma) -> None:
    lib().wko_rmsprop(x.dtype, _ptr(x.buf), _ptr(g.buf), _ptr(h.buf), _ptr(_scalar(x.np_dtype, lr)),
                      _ptr(_scalar(x.np_dtype, gamma)), x.layout.number_of_elements)


def transpose(result: OTensor, src: OTensor, dim0: int, dim1: int) -> None:
    rc = lib().wko_transpose(src.dtype, _ptr(src.buf), C.byref(src.layout), _ptr(result.buf), C.byref(result.layout),
                             dim0, dim1)
    if rc:
        raise ValueError(f"transpose rc={rc}")


def xxhash64(index: int, seed: int) -> int:
    return lib().wko_xxhash64(index, seed)
