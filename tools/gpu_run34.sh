#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream_kernels.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x --tb=short 2>&1 | tail -3
WK_SWEEP_ONLY=cosh,sinh timeout 300 python tools/stream_sweep.py gpurun_out/sweep_tmp8 27 2>&1 | tail -4
