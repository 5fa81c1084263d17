// benchmark/gemm.cpp -- the timing loop of the reference's benchmark/gemm.zig:150-283 on the CUDA backend: f32 (or f64)
// square GEMM, N = 4 * 2^i, ops NN / NT / TN / TT, A,B,C ~ U[0,1), fresh alpha,beta ~ U[0,1) per call, one warm-up call,
// average ms over `iters` calls including everything a call does (with PackedTensors, like the reference).
//   usage: gemm [f32|f64] [max_exp=11] [iters=10]
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>

#include "wekua.hpp"

using namespace wekua;

template <typename T> static void run(Context *ctx, Pipeline *p, int max_exp, int iters) {
    using blas::Operation;
    const Operation ops[4][2] = {{Operation::no_transpose, Operation::no_transpose}, {Operation::no_transpose, Operation::transpose},
                                 {Operation::transpose, Operation::no_transpose}, {Operation::transpose, Operation::transpose}};
    const char *names[4] = {"NN", "NT", "TN", "TT"};
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    std::printf("%8s %4s %12s %12s\n", "N", "op", "ms/call", "TFLOP/s");
    for (int e = 0; e <= max_exp; e++) {
        const uint64_t n = 4ull << e;
        auto a = Tensor<T>::alloc(ctx, p, {n, n}), b = Tensor<T>::alloc(ctx, p, {n, n}), c = Tensor<T>::alloc(ctx, p, {n, n});
        tensor_module::random::uniform<T>(p, a.get(), 42);
        tensor_module::random::uniform<T>(p, b.get(), 43);
        tensor_module::random::uniform<T>(p, c.get(), 44);
        auto packed = blas::PackedTensors<T>::init(p, c.get(), n, true);
        for (int o = 0; o < 4; o++) {
            blas::gemm<T>(p, (T)u01(rng), a.get(), ops[o][0], b.get(), ops[o][1], (T)u01(rng), c.get(), packed.get());  // warm-up
            p->waitAndCleanup();
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < iters; i++)
                blas::gemm<T>(p, (T)u01(rng), a.get(), ops[o][0], b.get(), ops[o][1], (T)u01(rng), c.get(), packed.get());
            p->waitAndCleanup();
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
            std::printf("%8llu %4s %12.4f %12.3f\n", (unsigned long long)n, names[o], ms, 2.0 * n * n * n / (ms * 1e-3) / 1e12);
        }
    }
}

int main(int argc, char **argv) {
    try {
        const bool f64 = argc > 1 && !std::strcmp(argv[1], "f64");
        const int max_exp = argc > 2 ? std::atoi(argv[2]) : 11, iters = argc > 3 ? std::atoi(argv[3]) : 10;
        auto context = core::Context::initFromDeviceType();
        auto pipeline = core::Pipeline::init(&context->command_queues[0]);
        if (f64) run<double>(context.get(), pipeline.get(), max_exp, iters);
        else run<float>(context.get(), pipeline.get(), max_exp, iters);
        return 0;
    } catch (const wekua::Error &e) {
        std::fprintf(stderr, "wekua error: %s\n", e.what());
        return 1;
    }
}
