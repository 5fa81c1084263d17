// benchmark/axpy.cpp -- the timing loop of the reference's benchmark/axpy.zig:119-203 on the CUDA backend: N = 4096 * 2^i,
// TOTAL ms for `iters` alternating axpy(x, a, y) / axpy(y, a, x) calls (alpha ~ U[-1,1)/sqrt(2) instead of the reference's
// U[-10,10) so 1000 alternations stay finite), plus the implied GB/s at 3*N*sizeof(T) bytes per call.
//   usage: axpy [f32|f64] [max_exp=16] [iters=1000]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "wekua.hpp"

using namespace wekua;

template <typename T> static void run(Context *ctx, Pipeline *p, int max_exp, int iters) {
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::printf("%12s %14s %12s\n", "N", "total ms", "GB/s");
    for (int e = 0; e <= max_exp; e++) {
        const uint64_t n = 4096ull << e;
        auto x = Tensor<T>::alloc(ctx, p, {n}), y = Tensor<T>::alloc(ctx, p, {n});
        tensor_module::random::uniform<T>(p, x.get(), 42);
        tensor_module::random::uniform<T>(p, y.get(), 43);
        blas::axpy<T>(p, x.get(), (T)0.5, y.get());  // warm-up, benchmark/axpy.zig:143-144
        p->waitAndCleanup();
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; i++) {
            const T alpha = (T)(u(rng) / std::sqrt(2.0));
            if (i % 2 == 0) blas::axpy<T>(p, x.get(), alpha, y.get());
            else blas::axpy<T>(p, y.get(), alpha, x.get());
        }
        p->waitAndCleanup();
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::printf("%12llu %14.3f %12.1f\n", (unsigned long long)n, ms, 3.0 * n * sizeof(T) * iters / (ms * 1e-3) / 1e9);
    }
}

int main(int argc, char **argv) {
    try {
        const bool f64 = argc > 1 && !std::strcmp(argv[1], "f64");
        const int max_exp = argc > 2 ? std::atoi(argv[2]) : 16, iters = argc > 3 ? std::atoi(argv[3]) : 1000;
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
