#!/bin/bash
# re-entry visit: full GPU parity suite, smoke, TF32 tensor-pipe ceiling, config-4 sweep, default bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 120 ./build/mma_peak_tf32 > gpurun_out/mma_peak_tf32.log 2>&1
timeout 600 python tools/gemm_sweep.py gpurun_out/sweep_r01 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01e.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_r01e.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/mma_peak_tf32.log; tail -4 gpurun_out/sweep.log; tail -3 gpurun_out/bench_r01e.log
