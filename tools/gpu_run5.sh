#!/bin/bash
# re-entry visit 2: full GPU parity suite (incl. complex dtypes), smoke, default bench, reference arm
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01f.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_r01f.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01f.log 2>&1; echo "ref exit $?" >> gpurun_out/bench_ref_r01f.log
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; tail -3 gpurun_out/bench_r01f.log | cut -c1-1200; tail -2 gpurun_out/bench_ref_r01f.log | cut -c1-800
