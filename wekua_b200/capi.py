"""ctypes binding of libwekua_b200.so (include/wekua_b200.h).  No torch types cross this boundary.

The library is built in-tree by `__graft_entry__.build()` / `make -C wekua_b200/csrc`.  If it is missing the
import FAILS LOUDLY -- there is no CPU or eager fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwekua_b200.so")

# status codes (include/wekua_b200.h) -> the reference's error names (src/tensor/main.zig:25-33)
OK = 0
_ERR_NAMES = {
    1: "InvalidValue", 2: "InvalidCoordinates", 3: "InvalidBuffer", 4: "UnqualTensorsAttribute",
    5: "UnqualTensorsShape", 6: "UnqualTensorsDimension", 7: "UnqualTensorsContext", 8: "OutOfMemory",
    9: "TypeNotSupported", 10: "DevicesArrayEmpty", 11: "CudaError",
}


class WekuaError(Exception):
    """Base of the TensorErrors set."""
    name = "WekuaError"


def _mk(name):
    return type(name, (WekuaError,), {"name": name})


InvalidValue = _mk("InvalidValue")
InvalidCoordinates = _mk("InvalidCoordinates")
InvalidBuffer = _mk("InvalidBuffer")
UnqualTensorsAttribute = _mk("UnqualTensorsAttribute")
UnqualTensorsShape = _mk("UnqualTensorsShape")
UnqualTensorsDimension = _mk("UnqualTensorsDimension")
UnqualTensorsContext = _mk("UnqualTensorsContext")
OutOfMemory = _mk("OutOfMemory")
TypeNotSupported = _mk("TypeNotSupported")
DevicesArrayEmpty = _mk("DevicesArrayEmpty")
CudaError = _mk("CudaError")
_ERR_CLASSES = {
    1: InvalidValue, 2: InvalidCoordinates, 3: InvalidBuffer, 4: UnqualTensorsAttribute, 5: UnqualTensorsShape,
    6: UnqualTensorsDimension, 7: UnqualTensorsContext, 8: OutOfMemory, 9: TypeNotSupported, 10: DevicesArrayEmpty,
    11: CudaError,
}


class QueueInfo(C.Structure):
    _fields_ = [
        ("device_name", C.c_char * 256),
        ("device_ordinal", C.c_int32),
        ("wekua_id", C.c_int32),
        ("compute_units", C.c_uint32),
        ("max_work_group_size", C.c_uint64),
        ("local_mem_size", C.c_uint64),
        ("local_mem_type", C.c_int32),
        ("cache_line_size", C.c_uint32),
        ("vector_widths", C.c_uint16 * 10),
        ("global_mem_size", C.c_uint64),
        ("cc_major", C.c_int32),
        ("cc_minor", C.c_int32),
    ]


class OptParam(C.Structure):
    """wk_opt_param_t: one (parameter, gradient, state) record of wk_optimizer_step_multi"""
    _fields_ = [("x", C.c_void_p), ("grad", C.c_void_p), ("state0", C.c_void_p), ("state1", C.c_void_p), ("n", C.c_uint64)]


_u64, _i32, _vp, _sz = C.c_uint64, C.c_int32, C.c_void_p, C.c_size_t
_pp = C.POINTER(C.c_void_p)

# every symbol include/wekua_b200.h declares (tests/test_capi_symbols.py checks the header against this table)
SIGNATURES = {
    "wk_device_count": [C.POINTER(_i32)],
    "wk_context_create": [C.POINTER(_i32), _i32, _pp],
    "wk_context_create_all": [_pp],
    "wk_context_destroy": [_vp],
    "wk_context_num_queues": [_vp, C.POINTER(_i32)],
    "wk_context_queue": [_vp, _i32, _pp],
    "wk_queue_wrap_stream": [_i32, _vp, _pp],
    "wk_queue_release": [_vp],
    "wk_queue_info": [_vp, C.POINTER(QueueInfo)],
    "wk_queue_finish": [_vp],
    "wk_queue_stream": [_vp, _pp],
    "wk_event_record": [_vp, _pp],
    "wk_queue_wait_event": [_vp, _vp],
    "wk_event_wait": [_vp],
    "wk_event_elapsed_ms": [_vp, _vp, C.POINTER(C.c_float)],
    "wk_event_release": [_vp],
    "wk_graph_begin_capture": [_vp],
    "wk_graph_end_capture": [_vp, _pp],
    "wk_graph_launch": [_vp, _vp],
    "wk_graph_num_kernels": [_vp, C.POINTER(C.c_uint64)],
    "wk_graph_release": [_vp],
    "wk_malloc": [_vp, _sz, _pp],
    "wk_free": [_vp, _vp],
    "wk_host_alloc": [_sz, _pp],
    "wk_host_free": [_vp],
    "wk_memset_zero": [_vp, _vp, _sz],
    "wk_h2d_rect": [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _sz],
    "wk_d2h_rect": [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _sz],
    "wk_d2d": [_vp, _vp, _vp, _sz],
    "wk_d2d_rect": [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _sz, _sz, _sz],
    "wk_put_value": [_vp, _vp, _sz, _vp, _sz],
    "wk_get_value": [_vp, _vp, _sz, _vp, _sz],
    "wk_gemm": [_vp, _i32, _i32, _i32, _u64, _u64, _u64, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _u64],
    "wk_gemm_bias_act": [_vp, _i32, _i32, _i32, _u64, _u64, _u64, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _i32],
    "wk_gemm_set_path": [_i32],
    "wk_axpy": [_vp, _i32, _u64, _u64, _u64, _vp, _vp, _u64, _u64, _vp, _u64, _u64],
    "wk_scal": [_vp, _i32, _u64, _u64, _u64, _vp, _vp, _u64, _u64],
    "wk_dot_reduce": [_vp, _i32, _u64, _u64, _u64, _vp, _u64, _u64, _vp, _u64, _u64, _vp],
    "wk_hadamard": [_vp, _i32, _u64, _u64, _u64, _vp, _u64, _u64, _vp, _u64, _u64],
    "wk_sum": [_vp, _i32, _u64, _u64, _u64, _u64, _vp, _vp],
    "wk_sum_async": [_vp, _i32, _u64, _u64, _u64, _u64, _vp, _vp],
    "wk_dot_reduce_async": [_vp, _i32, _u64, _u64, _u64, _vp, _u64, _u64, _vp, _u64, _u64, _vp],
    "wk_unary": [_vp, _i32, _i32, _vp, _u64],
    "wk_sigmoid_dev": [_vp, _i32, _vp, _vp, _u64],
    "wk_tanh_dev": [_vp, _i32, _vp, _vp, _u64],
    "wk_bias_add": [_vp, _i32, _vp, _vp, _u64, _u64],
    "wk_bias_step": [_vp, _i32, _vp, _vp, _u64, _u64, _u64],
    "wk_mse": [_vp, _i32, _vp, _vp, _vp, _vp, _u64],
    "wk_act_backward": [_vp, _i32, _i32, _vp, _vp, _vp, _u64],
    "wk_linear_backward_set_mode": [_i32],
    "wk_linear_backward": [_vp, _i32, _i32, _u64, _u64, _u64, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _vp, _u64],
    "wk_gdm": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _u64],
    "wk_adagrad": [_vp, _i32, _vp, _vp, _vp, _vp, _u64],
    "wk_rmsprop": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _u64],
    "wk_adam": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u64],
    "wk_optimizer_step_multi": [_vp, _i32, _i32, C.POINTER(OptParam), C.c_uint32, _vp, _vp, _vp, _vp, _u64],
    "wk_fill": [_vp, _i32, _u64, _u64, _u64, _vp, _u64, _u64, _vp],
    "wk_identity": [_vp, _i32, _vp, _u64, _u64, _u64],
    "wk_uniform": [_vp, _i32, _u64, _u64, _u64, _vp, _u64, _u64, _u64, _vp, _vp],
    "wk_transpose2d": [_vp, _i32, _u64, _u64, _vp, _u64, _vp, _u64],
    "wk_transpose_nd": [_vp, _i32, C.c_uint32, _vp, C.POINTER(_u64), _vp, C.POINTER(_u64), _u64, _u64, _u64, _u64, _u64,
                        C.c_uint32, C.c_uint32],
    "wk_gemm_rowshard_allgather": [_vp, _i32, _i32, _i32, _u64, _u64, _u64, _vp, _vp, _u64, _vp, _u64, _vp, _u64, _pp,
                                   _i32, _i32, _u64],
    "wk_ipc_get_handle": [_vp, _vp, _vp],
    "wk_ipc_open_handle": [_vp, _vp, _pp],
    "wk_ipc_close_handle": [_vp, _vp],
    "wk_enable_peer_access": [_vp, _i32],
}
_NON_STATUS = {"wk_last_error": (C.c_char_p, []), "wk_version": (C.c_char_p, []), "wk_launch_count": (C.c_uint64, [])}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  wekua_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = C.c_int32
            fn.argtypes = args
        for name, (res, args) in _NON_STATUS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    return lib().wk_last_error().decode(errors="replace")


def check(status: int) -> None:
    if status != OK:
        cls = _ERR_CLASSES.get(status, WekuaError)
        raise cls(f"{_ERR_NAMES.get(status, status)}: {last_error()}")


def launch_count() -> int:
    return int(lib().wk_launch_count())
