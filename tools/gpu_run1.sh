#!/bin/bash
# first GPU visit: parity tests, smoke, FP64 ceilings, tensor-core check, bench, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 120 ./build/mma_peak > gpurun_out/mma_peak.log 2>&1
timeout 400 python tools/tc_check.py f32 > gpurun_out/tc_check_f32.log 2>&1; echo "tc_check exit $?" >> gpurun_out/tc_check_f32.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench exit $?" >> gpurun_out/bench1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 2 --warmup 3 --n 8192 --no-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 3 -c 1 -o gpurun_out/prof_gemm_f32_r01 \
   python bench.py --steps 2 --warmup 3 --n 8192 --no-cpu --quick --no-e2e > gpurun_out/ncu_gemm.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/mma_peak.log; tail -30 gpurun_out/tc_check_f32.log; tail -3 gpurun_out/bench1.log
