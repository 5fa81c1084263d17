#!/bin/bash
mkdir -p gpurun_out
WK_SWEEP_ONLY=sum,sin,cos,tan,cosh WK_SWEEP_REPS=1 WK_SWEEP_WARM=0 timeout 600 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:'reduce_runs_kernel|UnaryF' -c 12 -o gpurun_out/prof_stream_r01d -f \
  python tools/stream_sweep.py gpurun_out/sweep_ncu_tmp 27 > gpurun_out/ncu_stream_d.log 2>&1; tail -2 gpurun_out/ncu_stream_d.log
