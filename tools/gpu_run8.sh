#!/bin/bash
# full GPU suite (golden fixtures, multi-tensor, XOR with fused GD), ncu --set full of the streaming kernels below 85 %
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
WK_SWEEP_ONLY=bias_add,transpose2d,tanh,tan,sigmoid,sin,uniform WK_SWEEP_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:'bias_add_kernel<double|transpose2d_vec_kernel<unsigned long|UnaryF<double, 5>|UnaryF<double, 2>|UnaryF<double, 6>|UnaryF<float, 6>|UnaryF<float, 0>|UniformF<float, 0>|UniformF<double, 0>' -c 36 -o gpurun_out/prof_stream_r01b -f \
  python tools/stream_sweep.py gpurun_out/sweep_ncu_tmp 27 > gpurun_out/ncu_stream.log 2>&1; tail -3 gpurun_out/ncu_stream.log
ls -la gpurun_out/*.ncu-rep
