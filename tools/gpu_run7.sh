#!/bin/bash
# multi-tensor optimizer + async reductions tests, XOR (fused GD), ncu --set full of streaming kernels below 85 %
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_tensor.py tests/test_gpu_stream_kernels.py tests/test_gpu_xor.py tests/test_gpu_linear.py -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_mt.log; tail -12 gpurun_out/pytest_mt.log
WK_SWEEP_ONLY=bias_add,transpose2d,tanh,tan,sigmoid,sin,uniform WK_SWEEP_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:'bias_add_kernel<double|transpose2d_vec_kernel<unsigned long|UnaryF<double, 5>|UnaryF<double, 2>|UnaryF<float, 6>|UnaryF<float, 0>|UniformF<float' -c 28 -o gpurun_out/prof_stream_r01b -f \
  python tools/stream_sweep.py gpurun_out/sweep_ncu_tmp 27 > gpurun_out/ncu_stream.log 2>&1; tail -3 gpurun_out/ncu_stream.log
ls -la gpurun_out/*.ncu-rep
