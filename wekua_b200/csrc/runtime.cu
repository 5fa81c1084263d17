// runtime.cu -- device context, queues (CUDA streams), events and memory of libwekua_b200.so.
// Replaces src/core/context.zig, command_queue.zig, pipeline.zig and src/tensor/memory/*.zig of the
// reference (OpenCL context / in-order command queue / cl_event chain / clEnqueue*BufferRect).
#include <stdarg.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace wk {

std::atomic<uint64_t> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();  // clear the sticky-less error state
    if (e == cudaErrorMemoryAllocation) return WK_ERR_OUT_OF_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return WK_ERR_NO_DEVICE;
    return WK_ERR_CUDA;
}

// Replace *buf by an allocation of want_bytes.  Outside of graphs: drain the stream, free the old one.  While capturing,
// or once this queue has captured a graph (its nodes keep the old address): retire the old buffer instead.
int32_t grow_buffer(wk_queue *q, void **buf, size_t *cur_bytes, size_t want_bytes) {
    if (*buf) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(q->stream, &st);
        if (st != cudaStreamCaptureStatusNone) q->ever_captured = true;
        if (q->ever_captured) {
            q->retired.push_back(*buf);  // never freed under a graph that may still replay; released with the queue
        } else {
            WK_CUDA(cudaStreamSynchronize(q->stream));
            WK_CUDA(cudaFree(*buf));
        }
        *buf = nullptr;
        *cur_bytes = 0;
    }
    WK_CUDA(cudaMalloc(buf, want_bytes));
    *cur_bytes = want_bytes;
    return WK_OK;
}

int32_t ensure_scratch(wk_queue *q, size_t bytes) {
    if (q->scratch_bytes >= bytes) return WK_OK;
    return grow_buffer(q, &q->scratch, &q->scratch_bytes, bytes < (1u << 20) ? (1u << 20) : bytes);
}

int32_t ensure_workspace(wk_queue *q, size_t bytes) {
    if (q->ws_bytes >= bytes) return WK_OK;
    return grow_buffer(q, &q->ws, &q->ws_bytes, bytes);
}

int32_t ensure_splitk(wk_queue *q, size_t ws_bytes, size_t n_tickets) {
    if (q->splitk_ws_bytes < ws_bytes) {
        int32_t rc = grow_buffer(q, &q->splitk_ws, &q->splitk_ws_bytes, ws_bytes);
        if (rc != WK_OK) return rc;
    }
    // The tickets are a fixed-capacity array created with the queue and zeroed with a BLOCKING memset there: zeroing them
    // here on the stream would only be *recorded* when the first split-K GEMM of a queue happens during graph capture, and
    // an eager launch before the graph's first replay would read uninitialised counters.
    if (q->splitk_n_tickets < n_tickets) {
        set_error("split-K needs %zu tickets, the queue has %zu", n_tickets, q->splitk_n_tickets);
        return WK_ERR_INVALID_VALUE;
    }
    return WK_OK;
}

static int32_t queue_init(wk_queue *q, int device, int wekua_id, cudaStream_t adopt) {
    q->device = device;
    q->wekua_id = wekua_id;
    WK_CUDA(cudaSetDevice(device));
    WK_CUDA(cudaGetDeviceProperties(&q->prop, device));
    q->sm_count = q->prop.multiProcessorCount;
    if (adopt != nullptr || wekua_id < 0) {
        q->stream = adopt;
        q->owns_stream = false;
        if (wekua_id < 0) q->wekua_id = 0;
    } else {
        WK_CUDA(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
        q->owns_stream = true;
    }
    WK_CUDA(cudaMallocHost(&q->pinned, 256));
    memset(q->pinned, 0, 256);
    WK_CUDA(cudaMalloc((void **)&q->reduce_ticket, 64));
    WK_CUDA(cudaMemset(q->reduce_ticket, 0, 64));
    // split-K tickets: 2 per 128-row block of a split tile (f32) / 1 per tile (f64); split launches never have more work
    // items than resident CTAs (<= 2 x 148), so 4096 is a hard upper bound with room
    q->splitk_n_tickets = 4096;
    WK_CUDA(cudaMalloc((void **)&q->splitk_tickets, q->splitk_n_tickets * sizeof(unsigned)));
    WK_CUDA(cudaMemset(q->splitk_tickets, 0, q->splitk_n_tickets * sizeof(unsigned)));
    return WK_OK;
}

static void queue_fini(wk_queue *q) {
    cudaSetDevice(q->device);
    if (q->scratch) cudaFree(q->scratch);
    if (q->ws) cudaFree(q->ws);
    q->ws = nullptr;
    for (void *r : q->retired) cudaFree(r);
    q->retired.clear();
    if (q->splitk_ws) cudaFree(q->splitk_ws);
    if (q->align_ws) cudaFree(q->align_ws);
    q->align_ws = nullptr;
    q->align_ws_bytes = 0;
    if (q->int_ws) cudaFree(q->int_ws);
    q->int_ws = nullptr;
    q->int_ws_bytes = 0;
    if (q->presplit_ws) cudaFree(q->presplit_ws);
    q->presplit_ws = nullptr;
    q->presplit_bytes = 0;
    if (q->splitk_tickets) cudaFree(q->splitk_tickets);
    q->splitk_ws = nullptr;
    q->splitk_tickets = nullptr;
    q->splitk_ws_bytes = q->splitk_n_tickets = 0;
    q->ws_bytes = 0;
    if (q->pinned) cudaFreeHost(q->pinned);
    if (q->reduce_ticket) cudaFree(q->reduce_ticket);
    q->reduce_ticket = nullptr;
    if (q->owns_stream && q->stream) cudaStreamDestroy(q->stream);
    q->scratch = nullptr;
    q->pinned = nullptr;
    q->stream = nullptr;
}

}  // namespace wk

using namespace wk;

WK_API const char *wk_last_error(void) { return g_err; }
WK_API const char *wk_version(void) { return "wekua_b200 0.1 (sm_100a)"; }
WK_API uint64_t wk_launch_count(void) { return g_launches.load(); }

WK_API int32_t wk_device_count(int32_t *count) {
    if (!count) return WK_ERR_INVALID_VALUE;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    *count = n;
    return WK_OK;
}

WK_API int32_t wk_context_create(const int32_t *device_ordinals, int32_t n, wk_context **out) {
    if (!out) return WK_ERR_INVALID_VALUE;
    *out = nullptr;
    if (n <= 0 || !device_ordinals) {  // Context.init: DevicesArrayEmpty
        set_error("empty device list");
        return WK_ERR_NO_DEVICE;
    }
    int have = 0;
    WK_CUDA(cudaGetDeviceCount(&have));
    for (int i = 0; i < n; i++)
        if (device_ordinals[i] < 0 || device_ordinals[i] >= have) {
            set_error("device ordinal %d out of range (%d devices)", device_ordinals[i], have);
            return WK_ERR_INVALID_VALUE;
        }
    wk_context *ctx = new (std::nothrow) wk_context();
    if (!ctx) return WK_ERR_OUT_OF_MEMORY;
    ctx->queues = new (std::nothrow) wk_queue[n];
    if (!ctx->queues) {
        delete ctx;
        return WK_ERR_OUT_OF_MEMORY;
    }
    ctx->n = n;
    for (int i = 0; i < n; i++) {
        int32_t rc = queue_init(&ctx->queues[i], device_ordinals[i], i, nullptr);
        if (rc != WK_OK) {
            for (int j = 0; j <= i; j++) queue_fini(&ctx->queues[j]);
            delete[] ctx->queues;
            delete ctx;
            return rc;
        }
    }
    // peers of one context can address each other's memory (row-sharded GEMM epilogue)
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j || ctx->queues[i].device == ctx->queues[j].device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctx->queues[i].device, ctx->queues[j].device);
            if (can) {
                cudaSetDevice(ctx->queues[i].device);
                cudaError_t e = cudaDeviceEnablePeerAccess(ctx->queues[j].device, 0);
                if (e != cudaSuccess) cudaGetLastError();
            }
        }
    *out = ctx;
    return WK_OK;
}

WK_API int32_t wk_context_create_all(wk_context **out) {
    int n = 0;
    int32_t rc = wk_device_count(&n);
    if (rc != WK_OK) return rc;
    if (n == 0) {
        set_error("no CUDA device");
        return WK_ERR_NO_DEVICE;
    }
    int32_t ids[64];
    if (n > 64) n = 64;
    for (int i = 0; i < n; i++) ids[i] = i;
    return wk_context_create(ids, n, out);
}

WK_API int32_t wk_context_destroy(wk_context *ctx) {
    if (!ctx) return WK_OK;
    for (int i = 0; i < ctx->n; i++) {
        cudaSetDevice(ctx->queues[i].device);
        cudaStreamSynchronize(ctx->queues[i].stream);
        queue_fini(&ctx->queues[i]);
    }
    delete[] ctx->queues;
    delete ctx;
    return WK_OK;
}

WK_API int32_t wk_context_num_queues(const wk_context *ctx, int32_t *n) {
    if (!ctx || !n) return WK_ERR_INVALID_VALUE;
    *n = ctx->n;
    return WK_OK;
}

WK_API int32_t wk_context_queue(wk_context *ctx, int32_t index, wk_queue **out) {
    if (!ctx || !out || index < 0 || index >= ctx->n) return WK_ERR_INVALID_VALUE;
    *out = &ctx->queues[index];
    return WK_OK;
}

WK_API int32_t wk_queue_wrap_stream(int32_t device_ordinal, void *cuda_stream, wk_queue **out) {
    if (!out) return WK_ERR_INVALID_VALUE;
    wk_queue *q = new (std::nothrow) wk_queue();
    if (!q) return WK_ERR_OUT_OF_MEMORY;
    int32_t rc = queue_init(q, device_ordinal, -1, (cudaStream_t)cuda_stream);
    if (rc != WK_OK) {
        delete q;
        return rc;
    }
    *out = q;
    return WK_OK;
}

WK_API int32_t wk_queue_release(wk_queue *q) {
    if (!q) return WK_OK;
    queue_fini(q);
    delete q;
    return WK_OK;
}

WK_API int32_t wk_queue_info(const wk_queue *q, wk_queue_info_t *info) {
    if (!q || !info) return WK_ERR_INVALID_VALUE;
    memset(info, 0, sizeof(*info));
    strncpy(info->device_name, q->prop.name, sizeof(info->device_name) - 1);
    info->device_ordinal = q->device;
    info->wekua_id = q->wekua_id;
    info->compute_units = (uint32_t)q->prop.multiProcessorCount;
    info->max_work_group_size = (uint64_t)q->prop.maxThreadsPerBlock;
    info->local_mem_size = (uint64_t)q->prop.sharedMemPerBlockOptin;
    info->local_mem_type = 1;
    info->cache_line_size = 128;
    for (int i = 0; i < 10; i++) info->vector_widths[i] = 1;
    info->global_mem_size = (uint64_t)q->prop.totalGlobalMem;
    info->cc_major = q->prop.major;
    info->cc_minor = q->prop.minor;
    return WK_OK;
}

WK_API int32_t wk_queue_finish(wk_queue *q) {
    WK_CHECK_QUEUE(q);
    WK_CUDA(cudaStreamSynchronize(q->stream));
    return WK_OK;
}

WK_API int32_t wk_queue_stream(const wk_queue *q, void **cuda_stream) {
    if (!q || !cuda_stream) return WK_ERR_INVALID_VALUE;
    *cuda_stream = (void *)q->stream;
    return WK_OK;
}

// ------------------------------------------------------------------------------------------ events
WK_API int32_t wk_event_record(wk_queue *q, wk_event **out) {
    WK_CHECK_QUEUE(q);
    if (!out) return WK_ERR_INVALID_VALUE;
    wk_event *e = new (std::nothrow) wk_event();
    if (!e) return WK_ERR_OUT_OF_MEMORY;
    e->device = q->device;
    cudaError_t ce = cudaEventCreate(&e->ev);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->ev, q->stream);
    if (ce != cudaSuccess) {
        if (e->ev) cudaEventDestroy(e->ev);
        delete e;
        return cuda_fail(ce, "event record", __FILE__, __LINE__);
    }
    *out = e;
    return WK_OK;
}

WK_API int32_t wk_queue_wait_event(wk_queue *q, wk_event *ev) {
    WK_CHECK_QUEUE(q);
    if (!ev) return WK_ERR_INVALID_VALUE;
    WK_CUDA(cudaStreamWaitEvent(q->stream, ev->ev, 0));
    return WK_OK;
}

WK_API int32_t wk_event_wait(wk_event *ev) {
    if (!ev) return WK_ERR_INVALID_VALUE;
    WK_CUDA(cudaEventSynchronize(ev->ev));
    return WK_OK;
}

WK_API int32_t wk_event_elapsed_ms(wk_event *start, wk_event *end, float *ms) {
    if (!start || !end || !ms) return WK_ERR_INVALID_VALUE;
    WK_CUDA(cudaEventElapsedTime(ms, start->ev, end->ev));
    return WK_OK;
}

WK_API int32_t wk_event_release(wk_event *ev) {
    if (!ev) return WK_OK;
    cudaSetDevice(ev->device);
    cudaEventDestroy(ev->ev);
    delete ev;
    return WK_OK;
}

// ------------------------------------------------------------------------------------------ graphs
// A training step of a small network is ~25 launches of microsecond kernels: launch-bound.  The ops
// enqueued between begin and end are captured instead of executed and replay as ONE graph launch.  Blocking entries
// (wk_sum, wk_get_value, wk_dot_reduce, wk_free) must not be called while capturing.
struct wk_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int device = 0;
    uint64_t kernel_nodes = 0;
};

WK_API int32_t wk_graph_begin_capture(wk_queue *q) {
    WK_CHECK_QUEUE(q);
    WK_CUDA(cudaStreamBeginCapture(q->stream, cudaStreamCaptureModeRelaxed));
    q->ever_captured = true;
    return WK_OK;
}

WK_API int32_t wk_graph_end_capture(wk_queue *q, wk_graph **out) {
    WK_CHECK_QUEUE(q);
    if (!out) return WK_ERR_INVALID_VALUE;
    *out = nullptr;
    wk_graph *g = new (std::nothrow) wk_graph();
    if (!g) return WK_ERR_OUT_OF_MEMORY;
    g->device = q->device;
    cudaError_t e = cudaStreamEndCapture(q->stream, &g->graph);
    if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
        return cuda_fail(e, "graph capture / instantiate", __FILE__, __LINE__);
    }
    size_t n = 0;
    if (cudaGraphGetNodes(g->graph, nullptr, &n) == cudaSuccess && n > 0) {
        cudaGraphNode_t *nodes = new (std::nothrow) cudaGraphNode_t[n];
        if (nodes && cudaGraphGetNodes(g->graph, nodes, &n) == cudaSuccess)
            for (size_t i = 0; i < n; i++) {
                cudaGraphNodeType t;
                if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) g->kernel_nodes++;
            }
        delete[] nodes;
    }
    *out = g;
    return WK_OK;
}

WK_API int32_t wk_graph_launch(wk_graph *g, wk_queue *q) {
    WK_CHECK_QUEUE(q);
    if (!g || !g->exec) return WK_ERR_INVALID_VALUE;
    WK_CUDA(cudaGraphLaunch(g->exec, q->stream));
    count_launch(g->kernel_nodes);
    return WK_OK;
}

WK_API int32_t wk_graph_num_kernels(const wk_graph *g, uint64_t *n) {
    if (!g || !n) return WK_ERR_INVALID_VALUE;
    *n = g->kernel_nodes;
    return WK_OK;
}

WK_API int32_t wk_graph_release(wk_graph *g) {
    if (!g) return WK_OK;
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return WK_OK;
}

// ------------------------------------------------------------------------------------------ memory
WK_API int32_t wk_malloc(wk_queue *q, size_t bytes, void **dptr) {
    WK_CHECK_QUEUE(q);
    if (!dptr || bytes == 0) return WK_ERR_INVALID_VALUE;
    WK_CUDA(cudaMalloc(dptr, bytes));
    return WK_OK;
}

WK_API int32_t wk_free(wk_queue *q, void *dptr) {
    WK_CHECK_QUEUE(q);
    if (!dptr) return WK_OK;
    WK_CUDA(cudaStreamSynchronize(q->stream));  // Tensor.release syncs the pipeline first (main.zig:254)
    WK_CUDA(cudaFree(dptr));
    return WK_OK;
}

WK_API int32_t wk_host_alloc(size_t bytes, void **hptr) {
    if (!hptr || bytes == 0) return WK_ERR_INVALID_VALUE;
    WK_CUDA(cudaMallocHost(hptr, bytes));
    return WK_OK;
}

WK_API int32_t wk_host_free(void *hptr) {
    if (!hptr) return WK_OK;
    WK_CUDA(cudaFreeHost(hptr));
    return WK_OK;
}

WK_API int32_t wk_memset_zero(wk_queue *q, void *dptr, size_t bytes) {
    WK_CHECK_QUEUE(q);
    if (!dptr) return WK_ERR_INVALID_BUFFER;
    WK_CUDA(cudaMemsetAsync(dptr, 0, bytes, q->stream));
    return WK_OK;
}

static int32_t rect_copy(wk_queue *q, void *dst, size_t dpitch, size_t dslice, const void *src, size_t spitch,
                         size_t sslice, size_t width, size_t height, size_t depth, cudaMemcpyKind kind) {
    if (!dst || !src) return WK_ERR_INVALID_BUFFER;
    if (width == 0 || height == 0 || depth == 0) return WK_ERR_INVALID_VALUE;
    if (dpitch == width && spitch == width && dslice == width * height && sslice == width * height) {
        WK_CUDA(cudaMemcpyAsync(dst, src, width * height * depth, kind, q->stream));
        return WK_OK;
    }
    for (size_t d = 0; d < depth; d++) {
        WK_CUDA(cudaMemcpy2DAsync((char *)dst + d * dslice, dpitch, (const char *)src + d * sslice, spitch, width, height,
                                  kind, q->stream));
    }
    return WK_OK;
}

WK_API int32_t wk_h2d_rect(wk_queue *q, void *dst, size_t dst_row_pitch, size_t dst_slice_pitch, const void *src_host,
                           size_t width_bytes, size_t height, size_t depth) {
    WK_CHECK_QUEUE(q);
    return rect_copy(q, dst, dst_row_pitch, dst_slice_pitch, src_host, width_bytes, width_bytes * height, width_bytes,
                     height, depth, cudaMemcpyHostToDevice);
}

WK_API int32_t wk_d2h_rect(wk_queue *q, void *dst_host, const void *src, size_t src_row_pitch, size_t src_slice_pitch,
                           size_t width_bytes, size_t height, size_t depth) {
    WK_CHECK_QUEUE(q);
    return rect_copy(q, dst_host, width_bytes, width_bytes * height, src, src_row_pitch, src_slice_pitch, width_bytes,
                     height, depth, cudaMemcpyDeviceToHost);
}

WK_API int32_t wk_d2d(wk_queue *q, void *dst, const void *src, size_t bytes) {
    WK_CHECK_QUEUE(q);
    if (!dst || !src) return WK_ERR_INVALID_BUFFER;
    // Large aligned spans inside THIS device's memory stream through the map kernel (faster than the driver's copy, and an
    // ordinary kernel node in graphs).  Anything that touches a peer (IPC-mapped buffers: the B shares of rowshard) stays on
    // the copy engines, which is the point of those pushes.
    if (bytes >= (8u << 20) && bytes % 4 == 0 && aligned16(dst) && aligned16(src)) {
        cudaPointerAttributes ad{}, as{};
        if (cudaPointerGetAttributes(&ad, dst) == cudaSuccess && cudaPointerGetAttributes(&as, src) == cudaSuccess &&
            ad.type == cudaMemoryTypeDevice && as.type == cudaMemoryTypeDevice && ad.device == q->device && as.device == q->device)
            return wk::copy_dense(q, dst, src, bytes);  // (counts itself as a launch)
        cudaGetLastError();
    }
    WK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, q->stream));
    return WK_OK;
}

WK_API int32_t wk_d2d_rect(wk_queue *q, void *dst, size_t dst_row_pitch, size_t dst_slice_pitch, const void *src,
                           size_t src_row_pitch, size_t src_slice_pitch, size_t width_bytes, size_t height, size_t depth) {
    WK_CHECK_QUEUE(q);
    return rect_copy(q, dst, dst_row_pitch, dst_slice_pitch, src, src_row_pitch, src_slice_pitch, width_bytes, height,
                     depth, cudaMemcpyDeviceToDevice);
}

WK_API int32_t wk_put_value(wk_queue *q, void *dptr, size_t byte_offset, const void *host_value, size_t size) {
    WK_CHECK_QUEUE(q);
    if (!dptr || !host_value || size == 0 || size > 16) return WK_ERR_INVALID_VALUE;
    // the host value may die right after the call: stage it in the stream (cudaMemcpyAsync from pageable memory
    // copies synchronously into a driver staging buffer before returning)
    WK_CUDA(cudaMemcpyAsync((char *)dptr + byte_offset, host_value, size, cudaMemcpyHostToDevice, q->stream));
    return WK_OK;
}

WK_API int32_t wk_get_value(wk_queue *q, const void *dptr, size_t byte_offset, void *host_value, size_t size) {
    WK_CHECK_QUEUE(q);
    if (!dptr || !host_value || size == 0 || size > 16) return WK_ERR_INVALID_VALUE;
    WK_CUDA(cudaMemcpyAsync(q->pinned, (const char *)dptr + byte_offset, size, cudaMemcpyDeviceToHost, q->stream));
    WK_CUDA(cudaStreamSynchronize(q->stream));
    memcpy(host_value, q->pinned, size);
    return WK_OK;
}

// ---------------------------------------------------------------------------------------- CUDA IPC
WK_API int32_t wk_ipc_get_handle(wk_queue *q, void *dptr, void *handle64) {
    WK_CHECK_QUEUE(q);
    if (!dptr || !handle64) return WK_ERR_INVALID_VALUE;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    WK_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, dptr));
    return WK_OK;
}

WK_API int32_t wk_ipc_open_handle(wk_queue *q, const void *handle64, void **dptr) {
    WK_CHECK_QUEUE(q);
    if (!dptr || !handle64) return WK_ERR_INVALID_VALUE;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    WK_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return WK_OK;
}

WK_API int32_t wk_ipc_close_handle(wk_queue *q, void *dptr) {
    WK_CHECK_QUEUE(q);
    if (!dptr) return WK_OK;
    WK_CUDA(cudaIpcCloseMemHandle(dptr));
    return WK_OK;
}

WK_API int32_t wk_enable_peer_access(wk_queue *q, int32_t peer) {
    WK_CHECK_QUEUE(q);
    if (peer == q->device) return WK_OK;
    int can = 0;
    WK_CUDA(cudaDeviceCanAccessPeer(&can, q->device, peer));
    if (!can) {
        set_error("device %d cannot access peer %d", q->device, peer);
        return WK_ERR_INVALID_VALUE;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return WK_OK;
    }
    WK_CUDA(e);
    return WK_OK;
}
