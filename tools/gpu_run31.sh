#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r01k.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r01k.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'])
for k in d['layer_step_streaming']['kernels']: print(k)" || tail -20 gpurun_out/bench_r01k.log
WK_SWEEP_ONLY=sum,cos timeout 120 python tools/stream_sweep.py gpurun_out/sweep_tmp7 27 2>&1 | tail -4
