#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_splitk.py tests/test_golden.py tests/test_gpu_stream_kernels.py tests/test_gpu_parity.py tests/test_gpu_linear.py tests/test_gpu_xor.py tests/test_gpu_complex.py -m gpu -q -x --tb=short 2>&1 | tail -30 > gpurun_out/pytest_splitk.log; tail -12 gpurun_out/pytest_splitk.log
timeout 300 python tools/gemm_time.py f32 512 1024 1536 2048 4096 2>&1 | tail -6
WK_GEMM_SPLITK=1 timeout 300 python tools/gemm_time.py f32 512 1024 1536 2048 2>&1 | tail -5
timeout 200 python tools/linear_time.py 256 4096 4096 2>&1 | tail -4
WK_SWEEP_ONLY=bias_add,bias_step,transpose2d timeout 300 python tools/stream_sweep.py gpurun_out/sweep_tmp 27 2>&1 | tail -8
