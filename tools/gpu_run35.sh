#!/bin/bash
# default path unchanged? + first look at the experimental pre-split variant (WK_GEMM_PRESPLIT=1)
(timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_splitk.py -m gpu -q -x -k "gemm or splitk" --tb=line 2>&1 | tail -2) &
wait
timeout 40 python tools/gemm_time.py f32 8192 16384 2>&1 | tail -2
echo "== presplit"
WK_GEMM_PRESPLIT=1 timeout 40 python tools/gemm_time.py f32 8192 16384 2>&1 | tail -2
WK_GEMM_PRESPLIT=1 timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm" --tb=line 2>&1 | tail -2
