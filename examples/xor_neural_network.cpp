// examples/xor_neural_network.cpp -- the reference's examples/xor_neural_network.zig, line for line, against the C++
// host mirror (include/wekua.hpp) of the CUDA backend: 2 -> 10 -> 1 sigmoid MLP, MSE, GD lr = 1, 300 iterations.
//   usage: xor_neural_network [seed]      (the reference seeds the weights from the wall clock; a seed makes it repeatable)
#include <cstdio>
#include <cstdlib>

#include "wekua.hpp"

using namespace wekua;
namespace nn_layer = nn::layer_module;
using FloatTensor = Tensor<float>;
using FloatLinear = nn_layer::linear_module::Linear<float>;

int main(int argc, char **argv) {
    try {
        auto context = core::Context::initFromDeviceType();  // xor_neural_network.zig:21-27
        auto *command_queue = &context->command_queues[0];
        auto pipeline = core::Pipeline::init(command_queue);

        const std::vector<float> expected_outputs_buf = {0, 1, 1, 0};
        const std::vector<float> inputs_buf = {1, 1, 0, 1, 1, 0, 0, 0};

        auto inputs = FloatTensor::alloc(context.get(), pipeline.get(), {4, 2});
        auto expected_outputs = FloatTensor::alloc(context.get(), pipeline.get(), {4, 1});
        tensor_module::memory::readFromBuffer(pipeline.get(), inputs.get(), inputs_buf);
        tensor_module::memory::readFromBuffer(pipeline.get(), expected_outputs.get(), expected_outputs_buf);

        auto seq_layers = nn_layer::sequential_module::Sequential<float>::init();
        auto activation_layer = nn::activation_module::Sigmoid<float>::init();
        nn_layer::ExtraParams extra1, extra2;
        if (argc > 1) {
            extra1.seed = std::strtoull(argv[1], nullptr, 10);
            extra2.seed = *extra1.seed + 1;
        }
        seq_layers->append(FloatLinear::init(context.get(), pipeline.get(), 2, 10, activation_layer, extra1));
        seq_layers->append(FloatLinear::init(context.get(), pipeline.get(), 10, 1, activation_layer, extra2));
        auto *layers = seq_layers->layer();

        auto cache = nn_layer::Cache<float>::init(context.get(), pipeline.get(), 4, {layers});
        auto optimizer = nn::optimizer_module::GD<float>::init({.lr = 1});
        auto *layer_cache = cache->getLayerCache(0);

        for (int it = 0; it < 300; it++) {  // xor_neural_network.zig:116-128
            auto *output = layers->forward(pipeline.get(), inputs.get(), layer_cache);
            nn::loss_module::mse<float>(true, pipeline.get(), output, expected_outputs.get(), cache.get(), nullptr);
            layers->backward(pipeline.get(), layer_cache, inputs.get(), nullptr);
            optimizer->step(pipeline.get(), cache.get());
        }

        auto *output = layers->forward(pipeline.get(), inputs.get(), layer_cache);
        std::vector<float> host(4);
        tensor_module::memory::writeToBuffer(pipeline.get(), output, host);
        pipeline->waitAndCleanup();
        std::printf("output:   [%.6f, %.6f, %.6f, %.6f]\n", host[0], host[1], host[2], host[3]);
        std::printf("expected: [%.6f, %.6f, %.6f, %.6f]\n", expected_outputs_buf[0], expected_outputs_buf[1], expected_outputs_buf[2],
                    expected_outputs_buf[3]);
        std::printf("kernel launches: %llu\n", (unsigned long long)wk_launch_count());
        optimizer->deinit(pipeline.get());
        cache->deinit(pipeline.get());
        return 0;
    } catch (const wekua::Error &e) {
        std::fprintf(stderr, "wekua error: %s\n", e.what());
        return 1;
    }
}
