"""Host-side mirror of src/core: Context (context.zig:13-190), CommandQueue (command_queue.zig:10-229) and
Pipeline (pipeline.zig:6-66).  An OpenCL context + in-order queue per device becomes a CUDA device list + one
stream per device behind the C ABI; the `prevEvents -> enqueue -> append` chain of Pipeline is the stream order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

# src/core/types.zig:36-57 real types, then :74-83 Complex(T) of each of them (ids 10-19)
REAL_TYPES = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]


def Complex(base) -> np.dtype:
    """core.types.Complex(T) (types.zig:6-11): struct { real: T, imag: T } as a numpy structured dtype -- numpy has
    complex64/128 only, the reference also has complex integers"""
    return np.dtype([("re", np.dtype(base)), ("im", np.dtype(base))])


SUPPORTED_TYPES = REAL_TYPES + [Complex(t) for t in REAL_TYPES]


def is_complex(dtype) -> bool:
    return get_type_index(dtype) >= 10


def base_type(dtype) -> np.dtype:
    return np.dtype(REAL_TYPES[get_type_index(dtype) % 10])


def storage_dtype(dtype) -> np.dtype:
    """the dtype host buffers of a tensor use: complex64/128 are accepted as names for Complex(f32/f64)"""
    return np.dtype(SUPPORTED_TYPES[get_type_index(dtype)])


def as_elements(values, dtype) -> np.ndarray:
    """host values -> contiguous array of the tensor's element type.  Complex(T): real input becomes {v, 0} (the
    reference tests' makeDataValue), python / numpy complex input {re, im}, (re, im) tuples for scalars"""
    dt = storage_dtype(dtype)
    if dt.names is None:
        return np.ascontiguousarray(values, dtype=dt)
    if isinstance(values, tuple) and len(values) == 2 and all(np.ndim(v) == 0 for v in values):
        out = np.zeros((), dtype=dt)
        out["re"], out["im"] = values
        return out
    a = np.asarray(values)
    if a.dtype == dt:
        return np.ascontiguousarray(a)
    out = np.zeros(a.shape, dtype=dt)
    if a.dtype.kind == "c":
        out["re"], out["im"] = a.real, a.imag
    else:
        out["re"] = a
    return out


def get_type_index(dtype) -> int:
    """core.types.getTypeIndex (types.zig:60-87)"""
    dt = np.dtype(dtype)
    if dt == np.complex64:
        return 18
    if dt == np.complex128:
        return 19
    for i, t in enumerate(SUPPORTED_TYPES):
        if np.dtype(t) == dt:
            return i
    raise capi.TypeNotSupported(f"Type not supported: {dt}")


get_type_id = get_type_index  # getTypeId (types.zig:89-104)


class CommandQueue:
    """command_queue.zig:10-28: the device capabilities callers read + the native queue handle."""

    def __init__(self, context, handle, owned_by_context=True):
        self.context = context
        self._h = handle
        self._owned_by_context = owned_by_context
        info = capi.QueueInfo()
        capi.check(capi.lib().wk_queue_info(handle, C.byref(info)))
        self.device_name = info.device_name.decode()
        self.device_ordinal = info.device_ordinal
        self.wekua_id = info.wekua_id
        self.compute_units = info.compute_units
        self.max_work_group_size = info.max_work_group_size
        self.local_mem_size = info.local_mem_size
        self.local_mem_type = "local" if info.local_mem_type == 1 else "global"
        self.cache_line_size = info.cache_line_size
        self.vector_widths = list(info.vector_widths)
        self.global_mem_size = info.global_mem_size
        self.compute_capability = (info.cc_major, info.cc_minor)

    def is_type_supported(self, dtype) -> bool:
        """CommandQueue.isTypeSupported (command_queue.zig:227-229)"""
        try:
            get_type_index(dtype)
            return True
        except capi.TypeNotSupported:
            return False

    def finish(self):
        capi.check(capi.lib().wk_queue_finish(self._h))

    @classmethod
    def from_stream(cls, device_ordinal: int, cuda_stream: int) -> "CommandQueue":
        """Adopt an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        h = C.c_void_p()
        capi.check(capi.lib().wk_queue_wrap_stream(device_ordinal, C.c_void_p(cuda_stream), C.byref(h)))
        ctx = Context.__new__(Context)
        ctx._h = None
        q = cls(ctx, h, owned_by_context=False)
        ctx.command_queues = [q]
        return q

    def release(self):
        if not self._owned_by_context and self._h is not None:
            capi.lib().wk_queue_release(self._h)
            self._h = None


class Context:
    """context.zig:13-190.  `init(devices)`, `init_from_device_type()` (= every visible B200)."""

    def __init__(self, device_ordinals):
        ids = list(device_ordinals)
        if len(ids) == 0:
            raise capi.DevicesArrayEmpty("DevicesArrayEmpty")
        arr = (C.c_int32 * len(ids))(*ids)
        h = C.c_void_p()
        capi.check(capi.lib().wk_context_create(arr, len(ids), C.byref(h)))
        self._h = h
        self._init_queues()

    def _init_queues(self):
        n = C.c_int32()
        capi.check(capi.lib().wk_context_num_queues(self._h, C.byref(n)))
        self.command_queues = []
        for i in range(n.value):
            qh = C.c_void_p()
            capi.check(capi.lib().wk_context_queue(self._h, i, C.byref(qh)))
            self.command_queues.append(CommandQueue(self, qh))

    @classmethod
    def init(cls, device_ordinals) -> "Context":
        return cls(device_ordinals)

    @classmethod
    def init_from_device_type(cls, device_type="all") -> "Context":
        """Context.initFromDeviceType (context.zig:27-37); every CUDA device is a 'gpu'."""
        if device_type not in ("all", "gpu", "default"):
            raise capi.DevicesArrayEmpty(f"no device of type {device_type}")
        self = cls.__new__(cls)
        h = C.c_void_p()
        capi.check(capi.lib().wk_context_create_all(C.byref(h)))
        self._h = h
        self._init_queues()
        return self

    @classmethod
    def init_from_best_device(cls, device_type="all") -> "Context":
        """Context.initFromBestDevice (context.zig:39-104): score = compute units x max work-group size"""
        n = C.c_int32()
        capi.check(capi.lib().wk_device_count(C.byref(n)))
        if n.value == 0 or device_type not in ("all", "gpu", "default"):
            raise capi.DevicesArrayEmpty("DevicesArrayEmpty")
        best, best_score = 0, -1
        for d in range(n.value):
            c = cls([d])
            q = c.command_queues[0]
            score = q.compute_units * q.max_work_group_size
            c.deinit()
            if score > best_score:
                best, best_score = d, score
        return cls([best])

    def deinit(self):
        if getattr(self, "_h", None) is not None:
            capi.lib().wk_context_destroy(self._h)
            self._h = None
            self.command_queues = []


class Pipeline:
    """pipeline.zig:6-66.  Ops enqueue on `command_queue` in order; `wait_and_cleanup` is the only sync point."""

    def __init__(self, command_queue: CommandQueue):
        self.command_queue = command_queue
        self._events = []

    @classmethod
    def init(cls, command_queue: CommandQueue) -> "Pipeline":
        return cls(command_queue)

    @property
    def q(self):
        return self.command_queue._h

    def prealloc(self, n: int) -> None:  # pipeline.zig:25-33 (capacity hint only)
        pass

    def prev_events(self):
        return self._events[-1:] if self._events else None

    def append(self) -> None:
        """records an event after the ops enqueued so far (pipeline.append)"""
        ev = C.c_void_p()
        capi.check(capi.lib().wk_event_record(self.q, C.byref(ev)))
        self._events.append(ev)

    def record_event(self):
        """the event the last enqueue returned (cl_event out-parameter of clEnqueue*); owned by this pipeline"""
        self.append()
        return self._events[-1]

    def wait_for(self, event) -> None:
        """make everything enqueued on this pipeline from now on wait for `event` of ANOTHER pipeline -- the
        wait-list argument of clEnqueue* (pipeline.prevEvents), for copy / compute overlap across queues"""
        capi.check(capi.lib().wk_queue_wait_event(self.q, event))

    def begin_capture(self) -> None:
        """record, instead of executing, everything enqueued on this pipeline until end_capture()"""
        capi.check(capi.lib().wk_graph_begin_capture(self.q))

    def end_capture(self) -> "Graph":
        g = C.c_void_p()
        capi.check(capi.lib().wk_graph_end_capture(self.q, C.byref(g)))
        return Graph(g)

    def wait_and_cleanup(self) -> None:
        capi.check(capi.lib().wk_queue_finish(self.q))
        self.clear()

    def clear(self) -> None:
        for ev in self._events:
            capi.lib().wk_event_release(ev)
        self._events = []

    def deinit(self) -> None:
        self.wait_and_cleanup()


class Graph:
    """a captured op sequence (CUDA graph): `launch(pipeline)` replays it as one launch"""

    def __init__(self, handle):
        self._h = handle
        n = C.c_uint64()
        capi.check(capi.lib().wk_graph_num_kernels(self._h, C.byref(n)))
        self.num_kernels = int(n.value)

    def launch(self, pipeline: Pipeline) -> None:
        capi.check(capi.lib().wk_graph_launch(self._h, pipeline.q))

    def release(self) -> None:
        if self._h is not None:
            capi.lib().wk_graph_release(self._h)
            self._h = None
