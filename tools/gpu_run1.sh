#!/bin/bash
# first GPU visit: parity tests, smoke, FP64 ceilings, a short bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 120 ./build/mma_peak > gpurun_out/mma_peak.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench exit $?" >> gpurun_out/bench1.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/mma_peak.log; tail -3 gpurun_out/bench1.log
