#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 1" "1 1"; do
  set -- $cfg
  echo "=== WK_GEMM_CTAS=$1 WK_GEMM_SPLIT=$2" 
  WK_GEMM_CTAS=$1 WK_GEMM_SPLIT=$2 timeout 300 python tools/tc_check.py f32 quick > gpurun_out/tc_check_c$1_s$2.log 2>&1; echo "exit $?" >> gpurun_out/tc_check_c$1_s$2.log
  grep -v "^OK" gpurun_out/tc_check_c$1_s$2.log | tail -12
done
WK_GEMM_CTAS=2 WK_GEMM_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 3 -c 1 -o gpurun_out/prof_gemm_f32_c2s1b \
   python bench.py --steps 2 --warmup 3 --n 8192 --no-cpu --quick --no-e2e > gpurun_out/ncu_gemm2.log 2>&1
