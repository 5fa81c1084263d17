#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_splitk.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_linear.py -m gpu -q -x --tb=short 2>&1 | tail -6
WK_DEBUG=1 timeout 100 python tools/gemm_time.py f32 4096 2>&1 | grep -m2 "wk\]"
echo "== tail split on";  timeout 300 python tools/gemm_time.py f32 2304 3072 4096 5120 6144 8192 2>&1 | tail -6
echo "== tail split off"; WK_GEMM_TAILSPLIT=0 timeout 300 python tools/gemm_time.py f32 2304 3072 4096 5120 6144 8192 2>&1 | tail -6
WK_SWEEP_ONLY=bias_step,tan,cosh,tanh,sum timeout 300 python tools/stream_sweep.py gpurun_out/sweep_tmp 27 2>&1 | tail -10
