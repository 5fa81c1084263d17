#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_tensor.py tests/test_gpu_xor.py tests/test_gpu_complex.py tests/test_golden.py tests/test_gpu_stream_kernels.py -m gpu -q -x --tb=short 2>&1 | tail -4
WK_SWEEP_ONLY=sum,dot_reduce,cos,cosh timeout 300 python tools/stream_sweep.py gpurun_out/sweep_tmp6 27 2>&1 | tail -8
timeout 200 python tools/xor_time.py 2>&1 | tail -4
