#!/bin/bash
# split-K (up to 8 splits) tests + timing, final streaming sweep, after-fix ncu captures, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_splitk.py tests/test_gpu_parity.py tests/test_gpu_linear.py -m gpu -q -x --tb=short 2>&1 | tail -8
SIZES="256 256 256 512 512 512 1024 1024 1024 1536 1536 1536 2048 2048 2048 256 4096 4096 64 512 4096 128 1024 8192"
timeout 300 python tools/gemm_small_time.py f32 $SIZES 2>&1 | tail -9 | tee gpurun_out/small_gemm_f32.log
timeout 600 python tools/stream_sweep.py gpurun_out/sweep_stream_r01 27 > gpurun_out/stream_sweep.log 2>&1; tail -3 gpurun_out/stream_sweep.log
WK_SWEEP_ONLY=bias_add,transpose2d,tanh,tan,sigmoid,sin,uniform,adam WK_SWEEP_REPS=1 WK_SWEEP_WARM=0 timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:'bias_add_kernel|transpose2d_vec_kernel|UnaryF|AdamF|UniformF<float, 0>|UniformF<double, 0>' -c 24 -o gpurun_out/prof_stream_r01c -f \
  python tools/stream_sweep.py gpurun_out/sweep_ncu_tmp 27 > gpurun_out/ncu_stream.log 2>&1; tail -2 gpurun_out/ncu_stream.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01g.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r01g.log | cut -c1-600
