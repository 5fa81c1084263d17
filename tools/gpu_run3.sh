#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 3 -c 1 -o gpurun_out/prof_gemm_f32_n32768 \
   python bench.py --steps 1 --warmup 3 --no-cpu --quick --no-e2e > gpurun_out/ncu_gemm3.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tf32x3 -s 3 -c 1 \
     python bench.py --steps 1 --warmup 3 --no-cpu --quick --no-e2e 2>&1 | grep -E "dram__bytes|time_duration" > gpurun_out/gemm_n32768_dram.txt
cat gpurun_out/gemm_n32768_dram.txt
cat > /tmp/axpyprof.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, wekua_b200 as wk
ctx = wk.Context.init([0]); pipe = wk.Pipeline.init(ctx.command_queues[0])
dt = np.float32 if sys.argv[1] == "f32" else np.float64
x = wk.Tensor.alloc(ctx, pipe, (1 << 28,), dt); y = wk.Tensor.alloc(ctx, pipe, (1 << 28,), dt)
wk.tensor.random.uniform(pipe, x, 42); wk.tensor.random.uniform(pipe, y, 43)
for _ in range(4):
    wk.blas.axpy(pipe, x, 0.5, y)
pipe.wait_and_cleanup()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_vec_kernel -s 3 -c 1 -o gpurun_out/prof_axpy_f32 python /tmp/axpyprof.py f32 > gpurun_out/ncu_axpy.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_vec_kernel -s 3 -c 1 -o gpurun_out/prof_axpy_f64 python /tmp/axpyprof.py f64 >> gpurun_out/ncu_axpy.log 2>&1
tail -2 gpurun_out/ncu_axpy.log
