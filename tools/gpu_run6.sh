#!/bin/bash
# streaming-kernel visit: new tests, HBM sweep of every streaming op, ncu --set full of the ones below 85 %
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream_kernels.py tests/test_gpu_parity.py -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest_stream.log; tail -4 gpurun_out/pytest_stream.log
timeout 600 python tools/stream_sweep.py gpurun_out/sweep_stream_r01 27 > gpurun_out/stream_sweep.log 2>&1; cat gpurun_out/stream_sweep.log
WK_SWEEP_ONLY=adagrad,adam,sigmoid,sin,uniform,sum WK_SWEEP_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'AdagradF|AdamF|UnaryFIfLi6|UnaryFIfLi0|UniformFIf|reduce_runs' -c 24 -o gpurun_out/prof_stream_r01 -f \
  python tools/stream_sweep.py gpurun_out/sweep_ncu_tmp 27 > gpurun_out/ncu_stream.log 2>&1; tail -3 gpurun_out/ncu_stream.log
ls -la gpurun_out/*.ncu-rep
