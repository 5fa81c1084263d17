#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_kernels.py tests/test_golden.py tests/test_gpu_multi_tensor.py tests/test_gpu_complex.py tests/test_gpu_xor.py -m gpu -q -x --tb=short 2>&1 | tail -6
timeout 600 python tools/stream_sweep.py gpurun_out/sweep_stream_r01b 27 > gpurun_out/stream_sweep.log 2>&1; cat gpurun_out/sweep_stream_r01b.md
