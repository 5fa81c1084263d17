#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
timeout 120 ./build/bin/xor_neural_network 42
timeout 300 ./build/bin/bench_gemm f32 12 10 | tail -16
timeout 300 ./build/bin/bench_axpy f32 16 1000 | tail -6
