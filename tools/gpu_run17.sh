#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stream_kernels.py tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -3
WK_SWEEP_ONLY=sum,tanh,sigmoid timeout 300 python tools/stream_sweep.py gpurun_out/sweep_tmp3 27 2>&1 | tail -6
WK_SWEEP_ONLY=tanh,sigmoid WK_SWEEP_REPS=1 WK_SWEEP_WARM=0 timeout 600 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:'UnaryF<double' -c 4 -o gpurun_out/prof_act64_r01 -f \
  python tools/stream_sweep.py gpurun_out/sweep_ncu_tmp 27 > gpurun_out/ncu_act64.log 2>&1; tail -2 gpurun_out/ncu_act64.log
