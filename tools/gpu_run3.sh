#!/bin/bash
mkdir -p gpurun_out
WK_GEMM_CTAS=2 WK_GEMM_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 3 -c 1 -o gpurun_out/prof_gemm_f32_c2s1 \
   python bench.py --steps 2 --warmup 3 --n 8192 --no-cpu --quick --no-e2e > gpurun_out/ncu_gemm2.log 2>&1
tail -3 gpurun_out/ncu_gemm2.log
