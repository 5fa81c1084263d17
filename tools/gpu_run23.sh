#!/bin/bash
mkdir -p gpurun_out
# cooperative cluster launches (tail split-K) cannot be replayed by ncu: plain launch for the profiling pass only
WK_GEMM_SPLITK_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu exit $?"; wc -l gpurun_out/launches_r01i.csv
