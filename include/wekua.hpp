// wekua.hpp -- header-only C++17 host mirror of wekua's Zig API over the C ABI of libwekua_b200.so.
//
// The reference host language is Zig (src/wekua.zig:1-16: core, tensor, blas, math, nn); Zig is not in the build
// image, so this is the compiled-language host side: same module / type / function names, argument order and error
// behaviour as the Zig modules, every body one call into include/wekua_b200.h.  examples/xor_neural_network.cpp and
// benchmark/{gemm,axpy}.cpp are the reference's programs rewritten line by line against it.  Errors are thrown as
// wekua::Error carrying the TensorErrors member name (src/tensor/main.zig:25-33).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "wekua_b200.h"

namespace wekua {

// ------------------------------------------------------------------------------------------------ errors
struct Error : std::runtime_error {
    int32_t status;
    Error(int32_t s, const std::string &what) : std::runtime_error(what), status(s) {}
};
inline const char *errorName(int32_t s) {
    switch (s) {
        case WK_ERR_INVALID_VALUE: return "InvalidValue";
        case WK_ERR_INVALID_COORDINATES: return "InvalidCoordinates";
        case WK_ERR_INVALID_BUFFER: return "InvalidBuffer";
        case WK_ERR_UNEQUAL_ATTRIBUTE: return "UnqualTensorsAttribute";
        case WK_ERR_UNEQUAL_SHAPE: return "UnqualTensorsShape";
        case WK_ERR_UNEQUAL_DIMENSION: return "UnqualTensorsDimension";
        case WK_ERR_UNEQUAL_CONTEXT: return "UnqualTensorsContext";
        case WK_ERR_OUT_OF_MEMORY: return "OutOfMemory";
        case WK_ERR_TYPE_NOT_SUPPORTED: return "TypeNotSupported";
        case WK_ERR_NO_DEVICE: return "DevicesArrayEmpty";
        default: return "CudaFailure";
    }
}
inline void check(int32_t s) {
    if (s != WK_OK) throw Error(s, std::string(errorName(s)) + ": " + wk_last_error());
}
[[noreturn]] inline void fail(int32_t s) { throw Error(s, errorName(s)); }

// ------------------------------------------------------------------------------------------------ core
namespace core {
namespace types {  // src/core/types.zig:60-87 (real dtypes)
template <typename T> constexpr int32_t getTypeIndex() {
    if (std::is_same<T, int8_t>::value) return 0;
    if (std::is_same<T, uint8_t>::value) return 1;
    if (std::is_same<T, int16_t>::value) return 2;
    if (std::is_same<T, uint16_t>::value) return 3;
    if (std::is_same<T, int32_t>::value) return 4;
    if (std::is_same<T, uint32_t>::value) return 5;
    if (std::is_same<T, int64_t>::value) return 6;
    if (std::is_same<T, uint64_t>::value) return 7;
    if (std::is_same<T, float>::value) return 8;
    if (std::is_same<T, double>::value) return 9;
    return -1;
}
}  // namespace types

struct Context;
struct CommandQueue {  // src/core/command_queue.zig:10-28
    wk_queue *handle = nullptr;
    Context *context = nullptr;
    wk_queue_info_t info{};
    int32_t wekua_id() const { return info.wekua_id; }
    const uint16_t *vector_widths() const { return info.vector_widths; }
    void finish() const { check(wk_queue_finish(handle)); }
};

struct Context {  // src/core/context.zig:13-190
    wk_context *handle = nullptr;
    std::vector<CommandQueue> command_queues;
    // tensors that outlive their Context must not touch it in their destructor: they hold this token and check it
    std::shared_ptr<bool> alive = std::make_shared<bool>(true);

    static std::unique_ptr<Context> init(const std::vector<int32_t> &device_ordinals) {  // Context.init :13
        std::unique_ptr<Context> c(new Context());
        check(wk_context_create(device_ordinals.data(), (int32_t)device_ordinals.size(), &c->handle));
        c->adopt();
        return c;
    }
    static std::unique_ptr<Context> initFromDeviceType() {  // Context.initFromDeviceType :27 (cl.device.Type.all)
        std::unique_ptr<Context> c(new Context());
        check(wk_context_create_all(&c->handle));
        c->adopt();
        return c;
    }
    ~Context() {
        *alive = false;
        if (handle) wk_context_destroy(handle);
    }

   private:
    void adopt() {
        int32_t n = 0;
        check(wk_context_num_queues(handle, &n));
        command_queues.resize(n);
        for (int32_t i = 0; i < n; i++) {
            check(wk_context_queue(handle, i, &command_queues[i].handle));
            check(wk_queue_info(command_queues[i].handle, &command_queues[i].info));
            command_queues[i].context = this;
        }
    }
};

struct Graph {  // a captured op sequence: launch() replays it as ONE launch (launch-bound layer steps)
    wk_graph *handle = nullptr;
    explicit Graph(wk_graph *g) : handle(g) {}
    Graph(const Graph &) = delete;
    Graph &operator=(const Graph &) = delete;
    ~Graph() { wk_graph_release(handle); }
    uint64_t numKernels() const {
        uint64_t n = 0;
        check(wk_graph_num_kernels(handle, &n));
        return n;
    }
};

struct Pipeline {  // src/core/pipeline.zig:6-66: the in-order CUDA stream IS the prevEvents -> append chain
    CommandQueue *command_queue;
    explicit Pipeline(CommandQueue *q) : command_queue(q) {}
    static std::unique_ptr<Pipeline> init(CommandQueue *q) { return std::unique_ptr<Pipeline>(new Pipeline(q)); }
    wk_queue *q() const { return command_queue->handle; }
    void prealloc(size_t) {}
    void waitAndCleanup() { command_queue->finish(); }
    void clear() {}
    void beginCapture() { check(wk_graph_begin_capture(q())); }
    std::unique_ptr<Graph> endCapture() {
        wk_graph *g = nullptr;
        check(wk_graph_end_capture(q(), &g));
        return std::unique_ptr<Graph>(new Graph(g));
    }
    void launch(const Graph &g) { check(wk_graph_launch(g.handle, q())); }
};
}  // namespace core
using core::CommandQueue;
using core::Context;
using core::Pipeline;

// ------------------------------------------------------------------------------------------------ tensor
struct CreateConfig {  // src/tensor/main.zig:35-39
    bool vectors_enabled = true;
};

template <typename T> struct Tensor {  // src/tensor/main.zig:62-279
    Context *context = nullptr;
    std::shared_ptr<bool> context_alive;
    void *buffer = nullptr;
    std::vector<uint64_t> shape, pitches;
    uint64_t depth = 1, rows = 1, rows_padded = 2, cols = 1;
    uint64_t row_pitch = 0, slice_pitch = 0, number_of_elements = 0, number_of_elements_without_padding = 0, size = 0;
    bool vectors_enabled = false;
    static constexpr int32_t type_index = core::types::getTypeIndex<T>();

    // Tensor.empty, main.zig:113-251 (layout law :142-222 with the queue's vector width, 1 on this backend)
    static std::unique_ptr<Tensor> empty(Context *ctx, Pipeline *p, const std::vector<uint64_t> &shape, CreateConfig cfg = {}) {
        if (shape.empty()) fail(WK_ERR_INVALID_VALUE);
        for (uint64_t s : shape)
            if (s == 0) fail(WK_ERR_INVALID_VALUE);
        std::unique_ptr<Tensor> t(new Tensor());
        t->context = ctx;
        t->context_alive = ctx->alive;
        t->shape = shape;
        uint64_t vw = 1;
        if (cfg.vectors_enabled)
            for (auto &q : ctx->command_queues) vw = std::max<uint64_t>(vw, q.info.vector_widths[type_index]);
        t->vectors_enabled = vw > 1;
        const size_t nd = shape.size(), last = nd - 1, pen = last >= 1 ? last - 1 : 0;
        for (size_t i = 0; i < pen; i++) t->depth *= shape[i];
        t->rows = nd >= 2 ? shape[pen] : 1;
        t->cols = shape[last];
        t->rows_padded = t->rows + t->rows % 2;
        t->number_of_elements_without_padding = t->depth * t->rows * t->cols;
        uint64_t rp = t->cols;
        if (vw > 1 && rp % vw) rp += vw - rp % vw;
        uint64_t rpv = rp / vw;
        if (rpv % 2) rp += vw;
        t->row_pitch = rp;
        t->slice_pitch = rp * t->rows_padded;
        t->number_of_elements = t->slice_pitch * t->depth;
        t->pitches.assign(nd, 0);
        uint64_t pitch = t->number_of_elements;
        const size_t ante = pen >= 1 ? pen - 1 : 0;
        for (size_t i = 0; i < ante; i++) {
            pitch /= shape[i];
            t->pitches[i] = pitch;
        }
        if (nd >= 3) t->pitches[ante] = t->slice_pitch;
        if (nd >= 2) t->pitches[pen] = rp;
        t->pitches[last] = 1;
        t->size = t->number_of_elements * sizeof(T);
        check(wk_malloc(p->q(), t->size, &t->buffer));
        return t;
    }
    // Tensor.alloc, main.zig:265-277: empty + zero the whole padded buffer
    static std::unique_ptr<Tensor> alloc(Context *ctx, Pipeline *p, const std::vector<uint64_t> &shape, CreateConfig cfg = {}) {
        auto t = empty(ctx, p, shape, cfg);
        check(wk_memset_zero(p->q(), t->buffer, t->size));
        return t;
    }
    void release(Pipeline *p) {  // main.zig:253-263 (synchronises first)
        if (buffer) check(wk_free(p->q(), buffer));
        buffer = nullptr;
    }
    ~Tensor() {  // the reference requires release(pipeline); a tensor dropped without it frees its buffer only while its Context lives
        if (buffer && context_alive && *context_alive && !context->command_queues.empty()) wk_free(context->command_queues[0].handle, buffer);
    }
    uint64_t pitchSum() const {
        uint64_t s = 0;
        for (uint64_t v : pitches) s += v;
        return s;
    }
};

namespace tensor_module {
namespace helpers {
template <typename T> void eqlTensorsShape(const Tensor<T> *a, const Tensor<T> *b) {  // tensor/helpers.zig:53-57
    if (a->shape != b->shape) fail(WK_ERR_UNEQUAL_SHAPE);
}
template <typename T> void eqlTensors(const Tensor<T> *a, const Tensor<T> *b) {
    eqlTensorsShape(a, b);
    if (a->vectors_enabled != b->vectors_enabled || a->number_of_elements != b->number_of_elements) fail(WK_ERR_UNEQUAL_ATTRIBUTE);
}
}  // namespace helpers

namespace memory {
// memory.readFromBuffer (host -> tensor), read_from_buffer.zig:13-63.  `n` host elements, dense; the host buffer must
// stay alive until pipeline.waitAndCleanup() (same rule as the reference's non-blocking rect write).
template <typename T> void readFromBuffer(Pipeline *p, Tensor<T> *t, const T *host, size_t n) {
    if (n != t->number_of_elements_without_padding) fail(WK_ERR_INVALID_BUFFER);
    check(wk_h2d_rect(p->q(), t->buffer, t->row_pitch * sizeof(T), t->slice_pitch * sizeof(T), host, t->cols * sizeof(T), t->rows,
                      t->depth));
}
template <typename T> void readFromBuffer(Pipeline *p, Tensor<T> *t, const std::vector<T> &host) {
    readFromBuffer(p, t, host.data(), host.size());
}
// memory.writeToBuffer (tensor -> host), write_to_buffer.zig:13-63; valid after waitAndCleanup()
template <typename T> void writeToBuffer(Pipeline *p, const Tensor<T> *t, T *host, size_t n) {
    if (n != t->number_of_elements_without_padding) fail(WK_ERR_INVALID_BUFFER);
    check(wk_d2h_rect(p->q(), host, t->buffer, t->row_pitch * sizeof(T), t->slice_pitch * sizeof(T), t->cols * sizeof(T), t->rows,
                      t->depth));
}
template <typename T> void writeToBuffer(Pipeline *p, const Tensor<T> *t, std::vector<T> &host) {
    writeToBuffer(p, t, host.data(), host.size());
}
template <typename T> void copy(Pipeline *p, const Tensor<T> *src, Tensor<T> *dst) {  // copy.zig:85-98
    helpers::eqlTensorsShape(src, dst);
    if (src->row_pitch == dst->row_pitch && src->slice_pitch == dst->slice_pitch) check(wk_d2d(p->q(), dst->buffer, src->buffer, src->size));
    else
        check(wk_d2d_rect(p->q(), dst->buffer, dst->row_pitch * sizeof(T), dst->slice_pitch * sizeof(T), src->buffer,
                          src->row_pitch * sizeof(T), src->slice_pitch * sizeof(T), src->cols * sizeof(T), src->rows, src->depth));
}
template <typename T> size_t offsetOf(const Tensor<T> *t, const std::vector<uint64_t> &coords) {
    if (coords.size() != t->shape.size()) fail(WK_ERR_INVALID_COORDINATES);
    size_t off = 0;
    for (size_t i = 0; i < coords.size(); i++) {
        if (coords[i] >= t->shape[i]) fail(WK_ERR_INVALID_COORDINATES);
        off += coords[i] * t->pitches[i];
    }
    return off * sizeof(T);
}
template <typename T> void putValue(Pipeline *p, Tensor<T> *t, const std::vector<uint64_t> &coords, T v) {
    check(wk_put_value(p->q(), t->buffer, offsetOf(t, coords), &v, sizeof(T)));
}
template <typename T> T getValue(Pipeline *p, const Tensor<T> *t, const std::vector<uint64_t> &coords) {
    T v{};
    check(wk_get_value(p->q(), t->buffer, offsetOf(t, coords), &v, sizeof(T)));
    return v;
}
}  // namespace memory

namespace fill {
template <typename T> void constant(Pipeline *p, Tensor<T> *t, T v) {  // fill.zig:15-70 (logical region only)
    check(wk_fill(p->q(), Tensor<T>::type_index, t->depth, t->rows, t->cols, t->buffer, t->row_pitch, t->slice_pitch, &v));
}
template <typename T> void one(Pipeline *p, Tensor<T> *t) { constant<T>(p, t, (T)1); }
template <typename T> void zeroes(Pipeline *p, Tensor<T> *t) { check(wk_memset_zero(p->q(), t->buffer, t->size)); }  // fill.zig:72-95
}  // namespace fill

template <typename T> void identity(Pipeline *p, Tensor<T> *t) {  // identity.zig:16-70
    for (uint64_t s : t->shape)
        if (s != t->shape[0]) fail(WK_ERR_INVALID_VALUE);
    check(wk_identity(p->q(), Tensor<T>::type_index, t->buffer, t->number_of_elements, t->shape[0], t->pitchSum()));
}
template <typename T> void transpose(Pipeline *p, Tensor<T> *result, const Tensor<T> *t, size_t dim0, size_t dim1) {  // transpose.zig:15-113
    const size_t nd = t->shape.size();
    if (result->shape.size() != nd) fail(WK_ERR_UNEQUAL_DIMENSION);
    if (dim0 >= nd || dim1 >= nd) fail(WK_ERR_INVALID_VALUE);
    if (t->number_of_elements_without_padding != result->number_of_elements_without_padding) fail(WK_ERR_UNEQUAL_DIMENSION);
    if (result->shape[dim0] != t->shape[dim1] || result->shape[dim1] != t->shape[dim0]) fail(WK_ERR_INVALID_VALUE);
    if (dim0 == dim1) return memory::copy(p, t, result);
    if (nd == 2) {
        check(wk_transpose2d(p->q(), Tensor<T>::type_index, t->rows, t->cols, t->buffer, t->row_pitch, result->buffer, result->row_pitch));
        return;
    }
    check(wk_transpose_nd(p->q(), Tensor<T>::type_index, (uint32_t)nd, t->buffer, t->pitches.data(), result->buffer, result->pitches.data(),
                          t->row_pitch, t->slice_pitch, t->rows * t->row_pitch, t->cols, t->number_of_elements,
                          (uint32_t)std::min(dim0, dim1), (uint32_t)std::max(dim0, dim1)));
}
namespace random {
// random.uniform(T, pipeline, tensor, ?seed, ?min, ?max), uniform.zig:60-123; seed == nullopt -> time() like :82
template <typename T>
void uniform(Pipeline *p, Tensor<T> *t, std::optional<uint64_t> seed = std::nullopt, std::optional<T> min = std::nullopt,
             std::optional<T> max = std::nullopt) {
    if (min.has_value() != max.has_value()) fail(WK_ERR_INVALID_VALUE);
    if (min && !(*min < *max)) fail(WK_ERR_INVALID_VALUE);
    const uint64_t s = seed ? *seed : (uint64_t)std::time(nullptr);
    check(wk_uniform(p->q(), Tensor<T>::type_index, t->depth, t->rows, t->cols, t->buffer, t->row_pitch, t->slice_pitch, s,
                     min ? &*min : nullptr, max ? &*max : nullptr));
}
}  // namespace random
}  // namespace tensor_module

// ------------------------------------------------------------------------------------------------ blas
namespace blas {
enum class Operation : int32_t { no_transpose = 0, transpose = 1 };  // gemm.zig:25-29

template <typename T> struct PackedTensors {  // gemm.zig:60-359: API-compatible handle, owns nothing (TMA reads in place)
    uint64_t n_size, m_size, k_size;
    static std::unique_ptr<PackedTensors> init(Pipeline *, const Tensor<T> *result, uint64_t k_size, bool /*vectors_enabled*/) {
        if (result->shape.size() != 2) fail(WK_ERR_INVALID_VALUE);
        return std::unique_ptr<PackedTensors>(new PackedTensors{result->shape[0], result->shape[1], k_size});
    }
    static std::unique_ptr<PackedTensors> initWithDimensions(Pipeline *, uint64_t n, uint64_t m, uint64_t k, bool = true) {
        return std::unique_ptr<PackedTensors>(new PackedTensors{n, m, k});
    }
    void validateTensors(const Tensor<T> *a, Operation op_a, const Tensor<T> *b, Operation op_b) const {  // gemm.zig:250-270
        const bool ta = op_a == Operation::transpose, tb = op_b == Operation::transpose;
        bool ok = ta ? (a->shape[1] == n_size && a->shape[0] == k_size) : (a->shape[0] == n_size && a->shape[1] == k_size);
        ok = ok && (tb ? (b->shape[1] == k_size && b->shape[0] == m_size) : (b->shape[0] == k_size && b->shape[1] == m_size));
        if (!ok) fail(WK_ERR_INVALID_VALUE);
    }
    void deinit(Pipeline *) {}
};

namespace detail {
template <typename T> void validateTensors(const Tensor<T> *a, const Tensor<T> *b, const Tensor<T> *c, Operation op_a, Operation op_b) {
    // gemm.zig:442-485
    if (a->context != b->context || a->context != c->context) fail(WK_ERR_UNEQUAL_CONTEXT);
    if (a->shape.size() != 2 || b->shape.size() != 2 || c->shape.size() != 2) fail(WK_ERR_INVALID_VALUE);
    const uint64_t am = a->shape[0], ak = a->shape[1], bk = b->shape[0], bn = b->shape[1], cm = c->shape[0], cn = c->shape[1];
    const bool ta = op_a == Operation::transpose, tb = op_b == Operation::transpose;
    bool ok;
    if (ta) ok = tb ? (am == bn && bk == cn && ak == cm) : (am == bk && bn == cn && ak == cm);
    else ok = tb ? (ak == bn && bk == cn && am == cm) : (ak == bk && bn == cn && am == cm);
    if (!ok) fail(WK_ERR_INVALID_VALUE);
}
// the reference's kernels run over the whole padded C; on its packed path the padding becomes beta*padding or 0
template <typename T> void finishPadding(Pipeline *p, Tensor<T> *c, const std::optional<T> &beta) {
    const uint64_t M = c->shape[0], N = c->shape[1];
    struct Region { size_t off; uint64_t rows, cols; };
    std::vector<Region> regions;
    if (c->row_pitch > N) regions.push_back({N * sizeof(T), c->rows_padded, c->row_pitch - N});
    if (c->rows_padded > M) regions.push_back({M * c->row_pitch * sizeof(T), c->rows_padded - M, N});
    for (auto &r : regions) {
        void *ptr = (char *)c->buffer + r.off;
        if (!beta) {
            T zero = (T)0;
            check(wk_fill(p->q(), Tensor<T>::type_index, 1, r.rows, r.cols, ptr, c->row_pitch, c->slice_pitch, &zero));
        } else {
            T b = *beta;
            check(wk_scal(p->q(), Tensor<T>::type_index, 1, r.rows, r.cols, &b, ptr, c->row_pitch, c->slice_pitch));
        }
    }
}
}  // namespace detail

// blas.gemm(T, pipeline, alpha, a, op_a, b, op_b, beta, c, packed), gemm.zig:834-874
template <typename T>
void gemm(Pipeline *p, std::optional<T> alpha, const Tensor<T> *a, Operation op_a, const Tensor<T> *b, Operation op_b,
          std::optional<T> beta, Tensor<T> *c, PackedTensors<T> *packed = nullptr) {
    detail::validateTensors(a, b, c, op_a, op_b);
    if (packed) packed->validateTensors(a, op_a, b, op_b);
    const uint64_t M = c->shape[0], N = c->shape[1], K = a->shape[op_a == Operation::transpose ? 0 : 1];
    check(wk_gemm(p->q(), Tensor<T>::type_index, (int32_t)op_a, (int32_t)op_b, M, N, K, alpha ? &*alpha : nullptr, a->buffer, a->row_pitch,
                  b->buffer, b->row_pitch, beta ? &*beta : nullptr, c->buffer, c->row_pitch));
    detail::finishPadding(p, c, beta);
}
// blas.axpy(T, pipeline, x, alpha, y), axpy.zig:93-169
template <typename T> void axpy(Pipeline *p, const Tensor<T> *x, std::optional<T> alpha, Tensor<T> *y) {
    tensor_module::helpers::eqlTensorsShape(x, y);
    check(wk_axpy(p->q(), Tensor<T>::type_index, x->depth, x->rows, x->cols, alpha ? &*alpha : nullptr, x->buffer, x->row_pitch,
                  x->slice_pitch, y->buffer, y->row_pitch, y->slice_pitch));
}
template <typename T> void scal(Pipeline *p, T alpha, Tensor<T> *x) {  // old_src/blas.c:41-67
    check(wk_scal(p->q(), Tensor<T>::type_index, x->depth, x->rows, x->cols, &alpha, x->buffer, x->row_pitch, x->slice_pitch));
}
}  // namespace blas

// ------------------------------------------------------------------------------------------------ math
namespace math {
template <typename T> void dot(Pipeline *p, Tensor<T> *x, const Tensor<T> *y) {  // basic.zig:17-76: x *= y
    tensor_module::helpers::eqlTensorsShape(x, y);
    check(wk_hadamard(p->q(), Tensor<T>::type_index, x->depth, x->rows, x->cols, x->buffer, x->row_pitch, x->slice_pitch, y->buffer,
                      y->row_pitch, y->slice_pitch));
}
template <typename T> T sum(Pipeline *p, const Tensor<T> *x) {  // basic.zig:131-203 (padded rows included, sum.cl:33-35)
    T out{};
    if (x->shape.back() > 1) {
        check(wk_sum(p->q(), Tensor<T>::type_index, x->depth, x->rows, x->row_pitch, x->slice_pitch, x->buffer, &out));
    } else {  // basic.zig:150-152,193-202
        uint64_t row_length = 1;
        for (size_t i = 0; i + 1 < x->shape.size(); i++) row_length *= x->shape[i];
        check(wk_sum(p->q(), Tensor<T>::type_index, 1, 1, row_length, row_length, x->buffer, &out));
    }
    return out;
}
template <typename T> T mean(Pipeline *p, const Tensor<T> *x) {  // basic.zig:206-240
    return sum(p, x) / (T)x->number_of_elements_without_padding;
}
#define WEKUA_UNARY(name, op) \
    template <typename T> void name(Pipeline *p, Tensor<T> *t) { check(wk_unary(p->q(), Tensor<T>::type_index, op, t->buffer, t->number_of_elements)); }
WEKUA_UNARY(sin, WK_OP_SIN)
WEKUA_UNARY(cos, WK_OP_COS)
WEKUA_UNARY(tan, WK_OP_TAN)
WEKUA_UNARY(sinh, WK_OP_SINH)
WEKUA_UNARY(cosh, WK_OP_COSH)
WEKUA_UNARY(tanh, WK_OP_TANH)
#undef WEKUA_UNARY
}  // namespace math

// ------------------------------------------------------------------------------------------------ nn
namespace nn {
namespace activation_module {
template <typename T> struct Activation {  // activation/main.zig: vtable {run, getDerivative}
    virtual ~Activation() {}
    virtual void run(Pipeline *p, Tensor<T> *net_output) const = 0;
    virtual void derivative(Pipeline *p, const Tensor<T> *output, Tensor<T> *d) const = 0;
    virtual int kind() const { return WK_ACT_NONE; }  // WK_ACT_* of the fused epilogues (0: no fused form)
    void getDerivative(Pipeline *p, const Tensor<T> *output, Tensor<T> *d) const {
        tensor_module::helpers::eqlTensors(output, d);
        derivative(p, output, d);
    }
};
template <typename T> struct Sigmoid : Activation<T> {  // sigmoid.zig:39-137
    static std::shared_ptr<Activation<T>> init() { return std::make_shared<Sigmoid<T>>(); }
    int kind() const override { return WK_ACT_SIGMOID; }
    void run(Pipeline *p, Tensor<T> *o) const override {
        check(wk_unary(p->q(), Tensor<T>::type_index, WK_OP_SIGMOID, o->buffer, o->number_of_elements));
    }
    void derivative(Pipeline *p, const Tensor<T> *o, Tensor<T> *d) const override {
        check(wk_sigmoid_dev(p->q(), Tensor<T>::type_index, o->buffer, d->buffer, o->number_of_elements));
    }
};
template <typename T> struct Tanh : Activation<T> {  // tanh.zig:25-84
    static std::shared_ptr<Activation<T>> init() { return std::make_shared<Tanh<T>>(); }
    int kind() const override { return WK_ACT_TANH; }
    void run(Pipeline *p, Tensor<T> *o) const override {
        check(wk_unary(p->q(), Tensor<T>::type_index, WK_OP_TANH, o->buffer, o->number_of_elements));
    }
    void derivative(Pipeline *p, const Tensor<T> *o, Tensor<T> *d) const override {
        check(wk_tanh_dev(p->q(), Tensor<T>::type_index, o->buffer, d->buffer, o->number_of_elements));
    }
};
}  // namespace activation_module

namespace layer_module {
template <typename T> using TensorPtr = std::unique_ptr<Tensor<T>>;

template <typename T> struct LayerCache {  // linear.zig:57-71 (one per Linear; Sequential holds one per sub-layer)
    std::vector<TensorPtr<T>> outputs, sensitivities, acti_derivatives, gradients, bias_gradients;
    std::vector<std::unique_ptr<blas::PackedTensors<T>>> forward_packed, grad_packed, sensitivity_packed;
    std::vector<std::unique_ptr<LayerCache>> children;  // Sequential
};

template <typename T> struct Layer {  // layer/main.zig:18-56
    virtual ~Layer() {}
    virtual void deinit(Pipeline *p) = 0;
    virtual std::unique_ptr<LayerCache<T>> prepareCache(Pipeline *p, uint64_t number_of_elements) = 0;
    virtual Tensor<T> *forward(Pipeline *p, Tensor<T> *input, LayerCache<T> *cache) = 0;
    virtual void backward(Pipeline *p, LayerCache<T> *cache, Tensor<T> *input, Tensor<T> *input_sensitivity) = 0;
    virtual Tensor<T> *getCachedOutput(LayerCache<T> *cache) = 0;
    virtual Tensor<T> *getSensitivity(LayerCache<T> *cache) = 0;
    // (parameter, gradient) pairs in the order gd.zig:55-94 walks them: weights then biases
    virtual void parameters(LayerCache<T> *cache, std::vector<std::pair<Tensor<T> *, Tensor<T> *>> &out) = 0;
};

struct ExtraParams {  // linear.zig:23-26
    uint64_t deep = 1;
    bool enable_bias = true;
    std::optional<uint64_t> seed;  // the reference seeds from the wall clock (uniform.zig:82); tests pin it
    // fused = true: forward runs gemm + bias + activation as ONE launch (wk_gemm_bias_act) wherever the output tensor has
    // no padding (padded outputs keep the reference's three launches, whose whole-buffer kernels define what the padding
    // holds), and backward folds act'(output) * sensitivity into one pass (wk_act_backward)
    bool fused = true;
};

namespace linear_module {
template <typename T> struct Linear : Layer<T> {  // linear.zig:88-678
    Context *context;
    std::shared_ptr<activation_module::Activation<T>> activation;
    bool bias_enabled, fused = true;
    std::vector<TensorPtr<T>> weights, bias;

    static std::unique_ptr<Linear> init(Context *ctx, Pipeline *p, uint64_t input, uint64_t output,
                                        std::shared_ptr<activation_module::Activation<T>> acti, ExtraParams extra = {}) {
        if (input == 0 || output == 0 || extra.deep == 0) fail(WK_ERR_INVALID_VALUE);
        std::unique_ptr<Linear> l(new Linear());
        l->context = ctx;
        l->activation = acti;
        l->bias_enabled = extra.enable_bias;
        l->fused = extra.fused;
        auto limits = [](uint64_t a, uint64_t b) { return (T)std::sqrt(6.0 / (double)(a + b)); };  // linear.zig:28-45
        for (uint64_t i = 0; i < extra.deep; i++) {
            const uint64_t in = i == 0 ? input : output;
            auto w = Tensor<T>::alloc(ctx, p, {output, in});
            const T lim = limits(in, output);
            tensor_module::random::uniform<T>(p, w.get(), extra.seed ? std::optional<uint64_t>(*extra.seed + i) : std::nullopt, -lim, lim);
            l->weights.push_back(std::move(w));
            if (extra.enable_bias) l->bias.push_back(Tensor<T>::alloc(ctx, p, {output}));
        }
        return l;
    }
    void deinit(Pipeline *p) override {
        for (auto &w : weights) w->release(p);
        for (auto &b : bias) b->release(p);
    }
    std::unique_ptr<LayerCache<T>> prepareCache(Pipeline *p, uint64_t n) override {  // linear.zig:219-378
        std::unique_ptr<LayerCache<T>> c(new LayerCache<T>());
        for (auto &w : weights) {
            const uint64_t out = w->shape[0];
            c->outputs.push_back(Tensor<T>::alloc(context, p, {n, out}));
            auto s = Tensor<T>::alloc(context, p, {n, out});
            tensor_module::fill::one<T>(p, s.get());
            c->sensitivities.push_back(std::move(s));
            c->acti_derivatives.push_back(Tensor<T>::alloc(context, p, {n, out}));
            c->gradients.push_back(Tensor<T>::alloc(context, p, w->shape));
            c->bias_gradients.push_back(Tensor<T>::alloc(context, p, {out}));
        }
        for (size_t i = 0; i < weights.size(); i++) {
            c->forward_packed.push_back(blas::PackedTensors<T>::init(p, c->outputs[i].get(), weights[i]->shape[1], true));
            c->grad_packed.push_back(blas::PackedTensors<T>::init(p, c->gradients[i].get(), n, true));
            if (i > 0) c->sensitivity_packed.push_back(blas::PackedTensors<T>::init(p, c->sensitivities[i - 1].get(), weights[i]->shape[0], true));
            else c->sensitivity_packed.push_back(blas::PackedTensors<T>::initWithDimensions(p, n, weights[0]->shape[1], weights[0]->shape[0]));
        }
        return c;
    }
    Tensor<T> *forward(Pipeline *p, Tensor<T> *input, LayerCache<T> *c) override {  // linear.zig:480-525
        using blas::Operation;
        Tensor<T> *in = input;
        for (size_t i = 0; i < weights.size(); i++) {
            Tensor<T> *out = c->outputs[i].get();
            const bool dense = out->row_pitch == out->cols && out->rows_padded == out->rows;
            if (fused && dense && (!activation || activation->kind() != WK_ACT_NONE)) {
                c->forward_packed[i]->validateTensors(in, Operation::no_transpose, weights[i].get(), Operation::transpose);
                check(wk_gemm_bias_act(p->q(), Tensor<T>::type_index, 0, 1, out->rows, out->cols, in->cols, in->buffer, in->row_pitch,
                                       weights[i]->buffer, weights[i]->row_pitch, out->buffer, out->row_pitch,
                                       bias_enabled ? bias[i]->buffer : nullptr, activation ? activation->kind() : WK_ACT_NONE));
                in = out;
                continue;
            }
            blas::gemm<T>(p, std::nullopt, in, Operation::no_transpose, weights[i].get(), Operation::transpose, std::nullopt, out,
                          c->forward_packed[i].get());
            if (bias_enabled)  // addBias, linear.zig:424-478
                check(wk_bias_add(p->q(), Tensor<T>::type_index, out->buffer, bias[i]->buffer, out->row_pitch, out->number_of_elements));
            if (activation) activation->run(p, out);
            in = out;
        }
        return in;
    }
    void backward(Pipeline *p, LayerCache<T> *c, Tensor<T> *input, Tensor<T> *input_sensitivity) override {  // linear.zig:579-678
        using blas::Operation;
        size_t index = weights.size() - 1;
        Tensor<T> *sens = c->sensitivities[index].get();
        Tensor<T> *output = c->outputs[index].get();
        for (;;) {
            Tensor<T> *d = c->acti_derivatives[index].get();
            if (activation && fused && activation->kind() != WK_ACT_NONE) {
                tensor_module::helpers::eqlTensors<T>(output, d);
                check(wk_act_backward(p->q(), Tensor<T>::type_index, activation->kind(), output->buffer, d->buffer, sens->buffer,
                                      sens->number_of_elements));
            } else if (activation) {
                activation->getDerivative(p, output, d);
                math::dot<T>(p, sens, d);
            }
            Tensor<T> *prev = index >= 1 ? c->outputs[index - 1].get() : input;
            output = prev;
            blas::gemm<T>(p, std::nullopt, sens, Operation::transpose, prev, Operation::no_transpose, std::nullopt,
                          c->gradients[index].get(), c->grad_packed[index].get());
            if (bias_enabled)  // getBiasSensitivity, linear.zig:534-577
                check(wk_bias_step(p->q(), Tensor<T>::type_index, sens->buffer, c->bias_gradients[index]->buffer, sens->row_pitch,
                                   sens->shape[0], c->bias_gradients[index]->row_pitch));
            Tensor<T> *next = index >= 1 ? c->sensitivities[index - 1].get() : input_sensitivity;
            if (!next) return;
            blas::gemm<T>(p, std::nullopt, sens, Operation::no_transpose, weights[index].get(), Operation::no_transpose, std::nullopt, next,
                          c->sensitivity_packed[index].get());
            if (index == 0) break;
            index--;
            sens = next;
        }
    }
    Tensor<T> *getCachedOutput(LayerCache<T> *c) override { return c->outputs.back().get(); }
    Tensor<T> *getSensitivity(LayerCache<T> *c) override { return c->sensitivities.back().get(); }
    void parameters(LayerCache<T> *c, std::vector<std::pair<Tensor<T> *, Tensor<T> *>> &out) override {
        for (size_t i = 0; i < weights.size(); i++) out.push_back({weights[i].get(), c->gradients[i].get()});
        if (bias_enabled)
            for (size_t i = 0; i < bias.size(); i++) out.push_back({bias[i].get(), c->bias_gradients[i].get()});
    }
};
}  // namespace linear_module

namespace sequential_module {
template <typename T> struct Sequential : Layer<T> {  // sequential.zig
    std::vector<std::unique_ptr<Layer<T>>> layers;
    static std::unique_ptr<Sequential> init() { return std::unique_ptr<Sequential>(new Sequential()); }
    void append(std::unique_ptr<Layer<T>> l) { layers.push_back(std::move(l)); }
    Layer<T> *layer() { return this; }
    void deinit(Pipeline *p) override {
        for (auto &l : layers) l->deinit(p);
    }
    std::unique_ptr<LayerCache<T>> prepareCache(Pipeline *p, uint64_t n) override {
        std::unique_ptr<LayerCache<T>> c(new LayerCache<T>());
        for (auto &l : layers) c->children.push_back(l->prepareCache(p, n));
        return c;
    }
    Tensor<T> *forward(Pipeline *p, Tensor<T> *input, LayerCache<T> *c) override {  // sequential.zig:211-228
        Tensor<T> *out = input;
        for (size_t i = 0; i < layers.size(); i++) out = layers[i]->forward(p, out, c->children[i].get());
        return out;
    }
    void backward(Pipeline *p, LayerCache<T> *c, Tensor<T> *input, Tensor<T> *input_gradient) override {  // sequential.zig:242-274
        for (size_t i = layers.size(); i-- > 0;) {
            Tensor<T> *in = i == 0 ? input : layers[i - 1]->getCachedOutput(c->children[i - 1].get());
            Tensor<T> *gr = i == 0 ? input_gradient : layers[i - 1]->getSensitivity(c->children[i - 1].get());
            layers[i]->backward(p, c->children[i].get(), in, gr);
        }
    }
    Tensor<T> *getCachedOutput(LayerCache<T> *c) override { return layers.back()->getCachedOutput(c->children.back().get()); }
    Tensor<T> *getSensitivity(LayerCache<T> *c) override { return layers.back()->getSensitivity(c->children.back().get()); }
    void parameters(LayerCache<T> *c, std::vector<std::pair<Tensor<T> *, Tensor<T> *>> &out) override {
        for (size_t i = 0; i < layers.size(); i++) layers[i]->parameters(c->children[i].get(), out);
    }
};
}  // namespace sequential_module

template <typename T> struct Cache {  // cache.zig:19-52
    struct Slot {
        std::unique_ptr<LayerCache<T>> cache;
        Layer<T> *layer;
    };
    std::vector<Slot> slots;
    TensorPtr<T> error_tensor;
    static std::unique_ptr<Cache> init(Context *ctx, Pipeline *p, uint64_t number_of_elements, const std::vector<Layer<T> *> &layers) {
        std::unique_ptr<Cache> c(new Cache());
        for (auto *l : layers) c->slots.push_back({l->prepareCache(p, number_of_elements), l});
        Tensor<T> *last = c->slots.back().layer->getSensitivity(c->slots.back().cache.get());
        c->error_tensor = Tensor<T>::alloc(ctx, p, last->shape);
        return c;
    }
    LayerCache<T> *getLayerCache(size_t i) { return slots[i].cache.get(); }
    void deinit(Pipeline *p) { p->waitAndCleanup(); }  // tensors free themselves (RAII) after the stream drained
};
}  // namespace layer_module

namespace loss_module {
// loss.mse(T, calc_dev, pipeline, output, expected, cache, ?*error_result), mse.zig:63-132
template <typename T>
void mse(bool calculate_derivative, Pipeline *p, const Tensor<T> *output, const Tensor<T> *expected, layer_module::Cache<T> *cache,
         T *error_result) {
    Tensor<T> *err = cache->error_tensor.get();
    tensor_module::helpers::eqlTensors(output, expected);
    tensor_module::helpers::eqlTensors<T>(err, output);
    void *dev = nullptr;
    if (calculate_derivative) {
        auto &last = cache->slots.back();
        dev = last.layer->getSensitivity(last.cache.get())->buffer;
    }
    check(wk_mse(p->q(), Tensor<T>::type_index, output->buffer, expected->buffer, err->buffer, dev, output->number_of_elements));
    if (error_result) *error_result = math::mean<T>(p, err);
}
}  // namespace loss_module

namespace optimizer_module {
template <typename T> struct Optimizer {  // optimizers/main.zig vtable {step, zero, deinit}
    using Cache = layer_module::Cache<T>;
    using Params = std::vector<std::pair<Tensor<T> *, Tensor<T> *>>;
    virtual ~Optimizer() {}
    virtual void step(Pipeline *p, Cache *cache) = 0;
    virtual void zero(Pipeline *) {}
    virtual void deinit(Pipeline *) {}
    static Params walk(Cache *cache) {  // (parameter, gradient) pairs in the order gd.zig:55-94 walks them
        Params params;
        for (auto &s : cache->slots) s.layer->parameters(s.cache.get(), params);
        return params;
    }
    // the whole parameter list in ONE launch (wk_optimizer_step_multi) instead of the reference's kernel per tensor;
    // bit-identical element-wise results.  state(x, i) returns the i-th state tensor of parameter x, or nullptr.
    template <typename StateOf>
    static void stepMulti(Pipeline *p, Cache *cache, int kind, StateOf state, const T *lr, const T *h0 = nullptr,
                          const T *h1 = nullptr, const T *h2 = nullptr, uint64_t t = 0) {
        std::vector<wk_opt_param_t> recs;
        for (auto &pg : walk(cache)) {
            Tensor<T> *s0 = state(pg.first, 0), *s1 = state(pg.first, 1);
            recs.push_back({pg.first->buffer, pg.second->buffer, s0 ? s0->buffer : nullptr, s1 ? s1->buffer : nullptr,
                            pg.first->number_of_elements});
        }
        check(wk_optimizer_step_multi(p->q(), Tensor<T>::type_index, kind, recs.data(), (uint32_t)recs.size(), lr, h0, h1, h2, t));
    }
};

template <typename T> struct GDConfig { T lr = (T)1; };
template <typename T> struct GD : Optimizer<T> {  // gd.zig:30-94: axpy(g, -lr, w); lr == 1 takes the SUBSTRACT kernel
    T neg_lr;
    static std::unique_ptr<GD> init(GDConfig<T> cfg = {}) {
        std::unique_ptr<GD> o(new GD());
        o->neg_lr = -cfg.lr;
        return o;
    }
    void step(Pipeline *p, typename Optimizer<T>::Cache *cache) override {
        // axpy touches the LOGICAL region only: tensors whose logical region is a contiguous prefix of the buffer share
        // one launch (wk_optimizer_step_multi), tensors with padded rows keep the pitched per-tensor kernel
        std::vector<wk_opt_param_t> recs;
        auto prefix = [](const Tensor<T> *t) { return t->depth == 1 && (t->rows == 1 || t->row_pitch == t->cols); };
        for (auto &pg : Optimizer<T>::walk(cache)) {
            if (prefix(pg.first) && prefix(pg.second))
                recs.push_back({pg.first->buffer, pg.second->buffer, nullptr, nullptr, pg.first->rows * pg.first->cols});
            else
                blas::axpy<T>(p, pg.second, neg_lr, pg.first);
        }
        check(wk_optimizer_step_multi(p->q(), Tensor<T>::type_index, WK_OPT_GD, recs.data(), (uint32_t)recs.size(), &neg_lr, nullptr,
                                      nullptr, nullptr, 0));
    }
};

// one state tensor per parameter (the reference's history index never advances)
template <typename T> struct Stateful : Optimizer<T> {
    std::vector<std::pair<Tensor<T> *, layer_module::TensorPtr<T>>> state;
    Tensor<T> *stateFor(Pipeline *p, Tensor<T> *x) {
        for (auto &s : state)
            if (s.first == x) return s.second.get();
        state.push_back({x, Tensor<T>::alloc(x->context, p, x->shape)});
        return state.back().second.get();
    }
    auto firstState(Pipeline *p) {
        return [this, p](Tensor<T> *x, int i) -> Tensor<T> * { return i == 0 ? stateFor(p, x) : nullptr; };
    }
    void zero(Pipeline *p) override {
        for (auto &s : state) tensor_module::fill::zeroes<T>(p, s.second.get());
    }
    void deinit(Pipeline *p) override {
        for (auto &s : state) s.second->release(p);
        state.clear();
    }
};
template <typename T> struct GDM : Stateful<T> {  // gdm.zig + gdm.cl:3-33
    T lr, beta;
    static std::unique_ptr<GDM> init(T lr, T beta) {
        std::unique_ptr<GDM> o(new GDM());
        o->lr = lr;
        o->beta = beta;
        return o;
    }
    void step(Pipeline *p, typename Optimizer<T>::Cache *cache) override {
        Optimizer<T>::stepMulti(p, cache, WK_OPT_GDM, this->firstState(p), &lr, &beta);
    }
};
template <typename T> struct Adagrad : Stateful<T> {  // adagrad.zig + adagrad.cl:3-48
    T lr;
    static std::unique_ptr<Adagrad> init(T lr) {
        std::unique_ptr<Adagrad> o(new Adagrad());
        o->lr = lr;
        return o;
    }
    void step(Pipeline *p, typename Optimizer<T>::Cache *cache) override {
        Optimizer<T>::stepMulti(p, cache, WK_OPT_ADAGRAD, this->firstState(p), &lr);
    }
};
template <typename T> struct RMSProp : Stateful<T> {  // rmsprop.zig:111-202 + rmsprop.cl:3-57
    T lr, gamma;
    static std::unique_ptr<RMSProp> init(T lr, T gamma = (T)0.9) {
        std::unique_ptr<RMSProp> o(new RMSProp());
        o->lr = lr;
        o->gamma = gamma;
        return o;
    }
    void step(Pipeline *p, typename Optimizer<T>::Cache *cache) override {
        Optimizer<T>::stepMulti(p, cache, WK_OPT_RMSPROP, this->firstState(p), &lr, &gamma);
    }
};
// Adam: adam.zig is an empty file in the reference; textbook bias-corrected update, two state tensors
template <typename T> struct Adam : Optimizer<T> {
    T lr, beta1, beta2, eps;
    uint64_t t = 0;
    Stateful<T> m, v;
    static std::unique_ptr<Adam> init(T lr = (T)1e-3, T beta1 = (T)0.9, T beta2 = (T)0.999, T eps = (T)1e-8) {
        std::unique_ptr<Adam> o(new Adam());
        o->lr = lr; o->beta1 = beta1; o->beta2 = beta2; o->eps = eps;
        return o;
    }
    void step(Pipeline *p, typename Optimizer<T>::Cache *cache) override {
        ++t;
        Optimizer<T>::stepMulti(p, cache, WK_OPT_ADAM,
                                [this, p](Tensor<T> *x, int i) -> Tensor<T> * { return i == 0 ? m.stateFor(p, x) : v.stateFor(p, x); },
                                &lr, &beta1, &beta2, &eps, t);
    }
    void zero(Pipeline *p) override { m.zero(p); v.zero(p); t = 0; }
    void deinit(Pipeline *p) override { m.deinit(p); v.deinit(p); }
};
}  // namespace optimizer_module
}  // namespace nn
}  // namespace wekua
