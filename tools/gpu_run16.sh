#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -6
WK_SWEEP_ONLY=sum,dot_reduce,tanh,sigmoid,cos,tan,cosh timeout 300 python tools/stream_sweep.py gpurun_out/sweep_tmp2 27 2>&1 | tail -16
for cfg in "2 16" "4 16" "4 32" "8 32"; do set -- $cfg
  echo "== e2e NJ=$1 PANELS=$2"; WK_E2E_NJ=$1 WK_E2E_PANELS=$2 timeout 300 python bench.py --quick --no-cpu --steps 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
