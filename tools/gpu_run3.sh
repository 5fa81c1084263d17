#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/f64prof.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, wekua_b200 as wk
ctx = wk.Context.init([0]); pipe = wk.Pipeline.init(ctx.command_queues[0])
n = 4096
a, b, c = (wk.Tensor.alloc(ctx, pipe, (n, n), np.float64) for _ in range(3))
wk.tensor.random.uniform(pipe, a, 42, -1, 1); wk.tensor.random.uniform(pipe, b, 43, -1, 1)
for _ in range(3):
    wk.blas.gemm(pipe, None, a, 0, b, 0, None, c)
pipe.wait_and_cleanup()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 2 -c 1 -o gpurun_out/prof_gemm_f64 python /tmp/f64prof.py > gpurun_out/ncu_f64.log 2>&1
tail -3 gpurun_out/ncu_f64.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench2.log
tail -2 gpurun_out/bench2.log
