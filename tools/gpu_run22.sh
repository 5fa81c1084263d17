#!/bin/bash
# final single-GPU pass of the session: full GPU suite, streaming sweep, default bench, reference arm, ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -5
timeout 600 python tools/stream_sweep.py gpurun_out/sweep_stream_r01c 27 > gpurun_out/stream_sweep.log 2>&1; grep -E "f64 (tan|tanh|cosh|sigmoid|cos|sin) |f32 (sum|bias_step) " gpurun_out/stream_sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01i.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r01i.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01i.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref_r01i.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu exit $?"; wc -l gpurun_out/launches_r01i.csv
